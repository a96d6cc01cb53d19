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
