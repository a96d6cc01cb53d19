#!/bin/bash
# Round-end validation on one B200 (through gpurun from the repo root): GPU test tier (both builds), smoke, the default
# bench line, a short run-to-run reproducibility check.  Outputs under gpurun_out/.
TAG=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_${TAG}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_${TAG}.log
tail -3 gpurun_out/gpu_tests_${TAG}.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n1.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["unet_step_ms"], d["roofline"]["frac"], d["clocks"], d["euler20_b1"]["images_per_s"])
PY
python tools/determinism_check.py euler20_b8 --reps 6 2>&1 | grep -E "^\["
