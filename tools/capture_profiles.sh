#!/bin/bash
# The ncu evidence bench.py and DESIGN.md refer to, captured on one B200 (run through gpurun from the repo root):
#   tools/capture_profiles.sh TAG     -> gpurun_out/TAG_*.{log,csv}; copy what is to be judged into profiles/
# Every ncu command runs only after the same command has exited 0 without ncu; numbers printed under ncu are not bench values.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
# 1. the bench line (driver contract) and its stderr (clock samples, side workloads)
python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err || exit 1
# 2. launch list of the bench command: per-launch durations, cold-cache and serialised -- the kernels' SHARES are what counts
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $OUT/${TAG}_launchlist_plain.log 2>&1 || exit 1
# (-k: this library's kernels only -- the first thousands of launches of the process are ATen's weight initialisation)
LIBK='regex:igemm_kernel|attention|gn_|splitk_reduce|step_|cfg_|axpby|image_to_u8|layernorm_kernel|softmax_rows|nchw_to_nhwc|nhwc_to_nchw|parity_split|silu_add|timestep_embedding|upsample2x|conv3x3_small'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$LIBK" -c 4000 --csv --log-file $OUT/${TAG}_launches_bench_euler20_b8.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > $OUT/${TAG}_launchlist_ncu.log 2>&1
# 3. DRAM traffic of the igemm launches of ONE UNet forward at batch 16 (the second executed forward: a graph replay)
python tools/unet_step.py --quick > $OUT/${TAG}_traffic_plain.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:igemm_kernel -s 222 -c 222 --csv \
    --log-file $OUT/${TAG}_igemm_dram_traffic_unet_b16.csv python tools/unet_step.py --quick > $OUT/${TAG}_traffic_ncu.log 2>&1
# 4. per-shape budgets of one UNet forward (CUDA events around every library call, eager)
python tools/prof_unet_launches.py > $OUT/${TAG}_unet_budget_b16.txt 2>&1
python tools/prof_unet_launches.py --batch 2 > $OUT/${TAG}_unet_budget_b2.txt 2>&1
python tools/prof_kernels.py --iters 10 > $OUT/${TAG}_per_shape_timings.txt 2>&1
tail -c 600 $OUT/${TAG}_bench_n1.json
