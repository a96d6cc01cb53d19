"""Build libcremage_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the authoring container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
DTYPES = ("fp16", "bf16")


def lib_path(dtype: str = "fp16") -> Path:
    assert dtype in DTYPES, dtype
    return PKG_DIR / f"libcremage_b200_{dtype}.so"


LIB_PATH = lib_path("fp16")
STAMP = PKG_DIR / ".libcremage_b200.stamp"

SOURCES = ["runtime.cu", "igemm.cu", "attention.cu", "attention64.cu", "plan.cu", "norm.cu", "elementwise.cu", "sampler.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcremage_b200.so")


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "cremage_b200.h"]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False):
    """Compile every CUDA source for sm_100a into one shared library per activation dtype (fp16 and bf16 builds of
    the same sources). Returns the list of library paths."""
    want = _source_hash()
    paths = [lib_path(d) for d in DTYPES]
    if not force and all(p.exists() for p in paths) and STAMP.exists() and STAMP.read_text().strip() == want:
        return paths
    procs = []
    for d in DTYPES:
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC)]
        if d == "fp16":
            cmd += ["-DCB_FP16"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += [str(CSRC / s) for s in SOURCES] + ["-o", str(lib_path(d))]
        procs.append((d, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for d, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out)
            failed = True
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libcremage_b200")
    STAMP.write_text(want)
    return paths


C_TEST_SRC = REPO_ROOT / "tests" / "c" / "cabi_plan_test.c"


def build_c_test(dtype: str = "fp16") -> Path:
    """gcc -std=c99: the C-only consumer of the C ABI (tests/c/cabi_plan_test.c) linked against the in-tree library and
    libcudart -- proves the boundary is usable without Python, torch or C++.  Returns the binary path."""
    out = C_TEST_SRC.with_name("cabi_plan_test" + ("" if dtype == "fp16" else "_" + dtype))
    lib = lib_path(dtype)
    if out.exists() and out.stat().st_mtime >= max(C_TEST_SRC.stat().st_mtime, lib.stat().st_mtime):
        return out
    cuda = Path(_nvcc()).resolve().parent.parent
    cmd = ["gcc", "-std=c99", "-Wall", "-I", str(REPO_ROOT / "include"), "-I", str(cuda / "include"), str(C_TEST_SRC), "-o", str(out),
           "-L", str(PKG_DIR), f"-l:{lib.name}", "-L", str(cuda / "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath,$ORIGIN/../../cremage_b200", f"-Wl,-rpath,{cuda / 'lib64'}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed building the C ABI test:\n" + r.stdout)
    return out


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
