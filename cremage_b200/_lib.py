"""ctypes binding of libcremage_b200.so (the C ABI declared in include/cremage_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if it cannot be built or loaded the
import of any compute entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

from . import build as _build

_LIBS = {}
# The library is built once per 16-bit storage type (fp16: the reference's own GPU precision, the default; bf16: fp32's
# exponent range).  DTYPE is the build the calling code is currently routed to: the process default comes from
# CREMAGE_B200_DTYPE, `ops.precision("bf16")` switches it for a region (the SDXL first stage, which the reference runs
# outside fp16 autocast because its activations overflow fp16: sgm/models/diffusion.py:119-137).
DTYPE = os.environ.get("CREMAGE_B200_DTYPE", "fp16").lower()
if DTYPE not in _build.DTYPES:
    raise ValueError(f"CREMAGE_B200_DTYPE must be one of {_build.DTYPES}, got {DTYPE!r}")
DEFAULT_DTYPE = DTYPE


class IGemmDesc(C.Structure):
    """Mirror of `cb_igemm_desc` (include/cremage_b200.h)."""

    _fields_ = [
        ("a0", C.c_void_p), ("c0", C.c_int64), ("a0_ld", C.c_int64),
        ("a1", C.c_void_p), ("c1", C.c_int64), ("a1_ld", C.c_int64),
        ("a_n", C.c_int64), ("a_h", C.c_int64), ("a_w", C.c_int64),
        ("n", C.c_int64), ("h", C.c_int64), ("w", C.c_int64),
        ("tw", C.c_int), ("th", C.c_int), ("tn", C.c_int),
        ("taps", C.c_int),
        ("tap_dw", C.c_int * 9), ("tap_dh", C.c_int * 9), ("tap_dn", C.c_int * 9),
        ("wgt", C.c_void_p), ("wgt_rows", C.c_int64),
        ("cout", C.c_int64),
        ("mode", C.c_int), ("act", C.c_int),
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p), ("rowbias_ld", C.c_int64),
        ("residual", C.c_void_p), ("res_ld", C.c_int64),
        ("out", C.c_void_p), ("out_ld", C.c_int64), ("out_f32", C.c_int),
        ("out_scale", C.c_float),
        ("heads_d", C.c_int), ("heads_dpad", C.c_int), ("heads_h", C.c_int), ("heads_tokens", C.c_int),
        ("heads_which_stride", C.c_int64),
        ("bn", C.c_int), ("stages", C.c_int), ("epilogue", C.c_int), ("cta_pair", C.c_int), ("nsub", C.c_int), ("ksplit", C.c_int),
        ("out_w_stride", C.c_int64), ("out_h_stride", C.c_int64), ("out_n_stride", C.c_int64),
        ("gn_partials", C.c_void_p),
        ("gn_rows_per_image", C.c_int64), ("gn_row_offset", C.c_int64),
        ("ln_partials_out", C.c_void_p), ("ln_partials_in", C.c_void_p), ("ln_in_slots", C.c_int),
        ("ln_dim", C.c_int64), ("ln_eps", C.c_float),
    ]


class IGemmPlan(C.Structure):
    """Mirror of `cb_igemm_plan_t` (include/cremage_b200.h)."""

    _fields_ = [("tw", C.c_int), ("th", C.c_int), ("tn", C.c_int),
                ("bn", C.c_int), ("cta_pair", C.c_int), ("nsub", C.c_int), ("ksplit", C.c_int),
                ("m_tiles", C.c_int64), ("workspace_bytes", C.c_int64),
                ("gn_fusable", C.c_int), ("gn_rows_per_image", C.c_int64),
                ("ln_out_slots", C.c_int), ("ln_foldable", C.c_int)]


_i64, _int, _f32, _vp = C.c_int64, C.c_int, C.c_float, C.c_void_p

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "cb_last_error": [],
    "cb_version": [],
    "cb_act_dtype": [],
    "cb_launch_count": [],
    "cb_igemm": [C.POINTER(IGemmDesc), _vp],
    "cb_igemm_plan": [C.POINTER(IGemmDesc), C.POINTER(IGemmPlan)],
    "cb_igemm_auto": [C.POINTER(IGemmDesc), _vp, _i64, _vp],
    "cb_pack_weight": [_vp, _i64, _i64, _i64, _int, _vp, _vp],
    "cb_attention": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _f32, _vp],
    "cb_softmax_rows": [_vp, _int, _i64, _vp, _i64, _i64, _i64, _f32, _vp],
    "cb_pointwise_nchw_to_nhwc": [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _f32, _vp, _vp],
    "cb_groupnorm_workspace_bytes": [_i64, _i64, _i64, _int],
    "cb_groupnorm_nhwc": [_vp, _i64, _vp, _i64, _i64, _i64, _int, _f32, _vp, _vp, _int, _vp, _vp, _vp],
    "cb_gn_partial_blocks": [_i64, _i64, _int, _int],
    "cb_groupnorm_from_partials": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _int, _f32, _vp, _vp, _int, _vp, _vp, _vp],
    "cb_layernorm": [_vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp],
    "cb_nchw_to_nhwc": [_vp, _int, _i64, _i64, _i64, _i64, _f32, _vp, _vp],
    "cb_add_nchw_to_nhwc": [_vp, _vp, _int, _i64, _i64, _i64, _vp, _vp],
    "cb_nhwc_to_nchw_f32": [_vp, _int, _i64, _i64, _i64, _i64, _vp, _vp],
    "cb_upsample2x_nhwc": [_vp, _i64, _i64, _i64, _i64, _vp, _vp],
    "cb_parity_split_nhwc": [_vp, _i64, _i64, _i64, _i64, _vp, _vp],
    "cb_timestep_embedding": [_vp, _i64, _int, _vp, _vp, _vp],
    "cb_conv3x3_small_cin": [_vp, _i64, _i64, _i64, _int, _i64, _vp, _vp, _i64, _vp, _vp],
    "cb_silu_add": [_vp, _vp, _i64, _vp, _vp],
    "cb_splitk_reduce": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp],
    "cb_diag_gaussian": [_vp, _vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _vp],
    "cb_cfg_scale_input": [_vp, _i64, _i64, _f32, _vp, _vp],
    "cb_axpby_f32": [_vp, _f32, _vp, _f32, _i64, _vp, _vp],
    "cb_cfg_mix_f32": [_vp, _vp, _f32, _i64, _vp, _vp],
    "cb_blend_mask_f32": [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp],
    "cb_step_euler_ancestral": [_vp, _vp, _vp, _int, _vp, _i64, _f32, _f32, _f32, _f32, _vp, _vp, _vp],
    "cb_step_dpmpp_2m": [_vp, _vp, _vp, _int, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp],
    "cb_step_ddim": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp],
    "cb_bilinear_upsample_f32": [_vp, _i64, _i64, _i64, _int, _vp, _vp],
    "cb_image_to_u8": [_vp, _i64, _i64, _i64, _vp, _vp],
}
_RESTYPES = {"cb_last_error": C.c_char_p, "cb_launch_count": C.c_int64, "cb_groupnorm_workspace_bytes": C.c_int64,
             "cb_gn_partial_blocks": C.c_int64}


def lib_path(dtype: str = None) -> Path:
    dtype = DTYPE if dtype is None else dtype
    override = os.environ.get("CREMAGE_B200_LIB")   # A/B comparison of library builds (tools/); default dtype only
    return Path(override) if (override and dtype == DEFAULT_DTYPE) else _build.lib_path(dtype)


def load(dtype: str = None) -> C.CDLL:
    """Load (building first if needed) the CUDA library of `dtype` (default: the current routing, `DTYPE`).
    Raises if that is impossible."""
    dtype = DTYPE if dtype is None else dtype
    lib = _LIBS.get(dtype)
    if lib is not None:
        return lib
    path = lib_path(dtype)
    if not path.exists():
        _build.build()
    lib = C.CDLL(str(path))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.cb_act_dtype() != {"fp16": 1, "bf16": 2}[dtype]:
        raise RuntimeError(f"{path} was not built for {dtype}")
    _LIBS[dtype] = lib
    return lib


class CremageB200Error(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise CremageB200Error(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    """Kernels launched directly (outside graph replays) by every loaded build of the library."""
    load()
    return sum(int(lib.cb_launch_count()) for lib in _LIBS.values())
