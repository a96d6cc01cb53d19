"""The bf16 build of the library (`libcremage_b200_bf16.so`, CREMAGE_B200_DTYPE=bf16) gets the same GPU test tier as
the default fp16 build: this test re-runs `pytest -m gpu` in a subprocess with the environment switched, so the
driver's single `pytest -m gpu` invocation exercises BOTH shared libraries.  Tolerances under bf16 are the fp16 ones
times the factor stated in each test file (bf16: 8 mantissa bits, unit round-off 8x fp16's)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gpu_tier_under_the_bf16_build():
    if os.environ.get("CREMAGE_B200_DTYPE", "fp16").lower() == "bf16":
        pytest.skip("already running under the bf16 build")
    env = dict(os.environ, CREMAGE_B200_DTYPE="bf16", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-q", "-p", "no:cacheprovider",
           "--deselect", "tests/test_gpu_bf16_build.py::test_gpu_tier_under_the_bf16_build", "-s"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
    tail = r.stdout[-6000:]
    keep = [ln for ln in r.stdout.splitlines() if ln.startswith("[parity]")]
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "gpu_tests_bf16.log"), "w") as f:
            f.write(r.stdout)
    except OSError:
        pass
    print("\n".join(keep[-80:]))
    print(tail[-1500:])
    assert r.returncode == 0, tail
