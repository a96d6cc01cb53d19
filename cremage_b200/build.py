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
LIB_PATH = PKG_DIR / "libcremage_b200.so"
STAMP = PKG_DIR / ".libcremage_b200.stamp"

SOURCES = ["runtime.cu", "igemm.cu", "attention.cu", "norm.cu", "elementwise.cu", "sampler.cu"]

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


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into one shared library. Returns the library path."""
    want = _source_hash()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == want:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES] + ["-o", str(LIB_PATH)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libcremage_b200.so")
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    STAMP.write_text(want)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
