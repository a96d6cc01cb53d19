"""The C-ABI library builds for sm_100a, loads on a GPU-less host and exports every symbol include/cremage_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cremage_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_library_exports_every_declared_symbol(dtype):
    from cremage_b200 import build
    build.build()
    lib = ctypes.CDLL(str(build.lib_path(dtype)))
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cremage_b200.h but not exported by {build.lib_path(dtype).name}"
    lib.cb_act_dtype.restype = ctypes.c_int
    assert lib.cb_act_dtype() == {"fp16": 1, "bf16": 2}[dtype]


def test_ctypes_signatures_cover_the_header():
    from cremage_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.cb_version() >= 100 and lib.cb_launch_count() == 0


def test_bad_arguments_return_errors_not_crashes():
    """Argument validation happens on the host before any launch, so it is testable without a GPU."""
    from cremage_b200 import _lib
    lib = _lib.load()
    rc = lib.cb_layernorm(None, 4, 64, 1e-5, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.cb_last_error()
    d = _lib.IGemmDesc()
    assert lib.cb_igemm(ctypes.byref(d), None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(_lib.CremageB200Error):
        _lib.check(-2, "x")
    assert lib.cb_groupnorm_workspace_bytes(320, 2, 4096, 32) > 0 and lib.cb_groupnorm_workspace_bytes(7, 2, 4096, 32) == 0


def test_c99_consumer_of_the_abi_builds():
    """tests/c/cabi_plan_test.c -- plain C, no torch / Python / C++ -- compiles and links against the in-tree library."""
    from cremage_b200 import build
    for dt in build.DTYPES:
        exe = build.build_c_test(dt)
        assert exe.exists() and exe.stat().st_size > 0


def test_plan_runs_without_a_gpu_and_matches_the_measured_policy():
    """cb_igemm_plan is host code: the tiling policy of round 1 (ops.py) now lives in csrc/plan.cu."""
    import ctypes as C
    from cremage_b200 import _lib
    lib = _lib.load()

    def plan(n, h, w, c0, cout, taps, **kw):
        d, p = _lib.IGemmDesc(), _lib.IGemmPlan()
        d.n, d.h, d.w, d.a_n, d.a_h, d.a_w, d.c0, d.cout, d.taps = n, h, w, n, h, w, c0, cout, taps
        for k, v in kw.items():
            setattr(d, k, v)
        assert lib.cb_igemm_plan(C.byref(d), C.byref(p)) == 0
        return p
    p = plan(16, 64, 64, 320, 320, 9)                       # SD1.5 top-level conv3x3, UNet batch 16: MMA bound -> CTA pairs
    assert (p.tw, p.th, p.tn) == (64, 2, 1) and p.cta_pair == 1 and p.ksplit == 1 and p.gn_fusable == 1
    p = plan(2, 8, 8, 1280, 1280, 9)                        # deepest level at batch 1: one M tile -> split K by tap groups
    assert p.m_tiles == 1 and p.cta_pair == 0 and p.ksplit in (3, 9) and p.workspace_bytes == p.ksplit * 128 * 1280 * 4
    p = plan(1, 1, 65536, 320, 320, 1)                      # K = 320 linear: epilogue bound -> single CTAs
    assert p.cta_pair == 0 and p.ksplit == 1 and (p.tw, p.th, p.tn) == (128, 1, 1)
    p = plan(1, 1, 65536, 320, 320, 1, cta_pair=1, bn=64)   # pinned fields are kept
    assert p.cta_pair == 1 and p.bn == 64
    # LayerNorm statistics ride on the staged epilogue: AUTO stages small-K launches only, so a K = 5120 producer of the
    # residual stream (ff.net.2) carries them only when the caller pins the staged form (ops.igemm does, for ln_stats)
    p = plan(1, 1, 4096, 5120, 1280, 1)
    assert p.ln_out_slots == 0
    p = plan(1, 1, 4096, 5120, 1280, 1, epilogue=2)
    assert p.ln_out_slots > 0 and p.ksplit == 1
    p = plan(1, 1, 4096, 1280, 1280, 1)
    assert p.ln_out_slots > 0


@pytest.mark.gpu
def test_c99_consumer_of_the_abi_runs():
    """The C program plans (cb_igemm_plan), packs (cb_pack_weight) and runs (cb_igemm_auto, incl. the split-K workspace
    path) a linear layer and a conv3x3 and checks them against C triple loops -- no Python between it and the library."""
    import subprocess
    from cremage_b200 import _lib, build
    exe = build.build_c_test(_lib.DEFAULT_DTYPE)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    assert "ksplit 3" in r.stdout or "ksplit 9" in r.stdout
