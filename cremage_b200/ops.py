"""Torch-tensor front end of the C ABI (include/cremage_b200.h): argument marshalling only, no arithmetic.

Every function here enqueues hand-written sm_100a kernels on the current CUDA stream through ctypes. Activations are
NHWC 16-bit (`ops.ACT`: fp16 by default, bf16 under CREMAGE_B200_DTYPE=bf16 or `ops.precision("bf16")`). Weight repacking helpers (pure layout work done once at load time) live here too.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import IGemmDesc, check

EPI_LINEAR, EPI_GEGLU, EPI_HEADS = 0, 1, 2
EPILOGUE_AUTO, EPILOGUE_DIRECT, EPILOGUE_STAGED = 0, 1, 2
ACT_NONE, ACT_SILU = 0, 1

# 16-bit activation / weight dtype the calls are currently routed to (fp16 unless CREMAGE_B200_DTYPE=bf16); read it as
# `ops.ACT` at call time -- `precision()` switches it together with the library build for a region.
_TORCH_DTYPE = {"fp16": torch.float16, "bf16": torch.bfloat16}
ACT = _TORCH_DTYPE[_lib.DTYPE]


class precision:
    """`with ops.precision("bf16"):` routes every call in the region to the bf16 build of the library (and makes
    `ops.ACT` torch.bfloat16).  Weight packs, workspaces and captured graphs are keyed by the dtype, so a module may be
    used under both.  Not re-entrant across threads (one Python thread drives the GPU, as in the reference)."""

    def __init__(self, dtype: str):
        if dtype not in _TORCH_DTYPE:
            raise ValueError(f"precision must be one of {tuple(_TORCH_DTYPE)}, got {dtype!r}")
        self.dtype = dtype

    def __enter__(self):
        global ACT
        self._prev = _lib.DTYPE
        _lib.DTYPE = self.dtype
        ACT = _TORCH_DTYPE[self.dtype]
        _lib.load(self.dtype)
        return self

    def __exit__(self, *exc):
        global ACT
        _lib.DTYPE = self._prev
        ACT = _TORCH_DTYPE[self._prev]
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cremage_b200 has no CPU path: tensors must live on a CUDA device")


# ------------------------------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py's roofline section): CUDA events on the launching stream around every call
# ------------------------------------------------------------------------------------------------------------------
class LaunchProfile:
    """Collects (name, algorithmic flops, algorithmic bytes, start event, end event) for every C-ABI call."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _PROF
        self._prev, _PROF = _PROF, self
        return self

    def __exit__(self, *exc):
        global _PROF
        _PROF = self._prev
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, flops, nbytes, e0, e1, _tag in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out

    def by_shape(self):
        """Per (kernel, shape tag): launches, total ms, algorithmic flops / bytes -- the per-shape time budget."""
        torch.cuda.synchronize()
        out = {}
        for name, flops, nbytes, e0, e1, tag in self.records:
            d = out.setdefault((name, tag), {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


_PROF: Optional[LaunchProfile] = None


def _launch(name: str, fn, flops: float = 0.0, nbytes: float = 0.0, tag: str = "") -> None:
    if _PROF is None:
        check(fn(), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(), name)
    e1.record()
    _PROF.records.append((name, flops, nbytes, e0, e1, tag))


def ceil64(c: int) -> int:
    return (c + 63) // 64 * 64


# ------------------------------------------------------------------------------------------------------------------
# weight repacking (host side, once per load)
# ------------------------------------------------------------------------------------------------------------------
def pack_weight(w: torch.Tensor, splits: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """OIHW conv weight or [O, I] linear weight -> bf16 [O][tap][pad64(c0) | pad64(c1)] (K-major, zero padded)."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    o, i, kh, kw = w.shape
    c0, c1 = (i, 0) if splits is None else splits
    assert c0 + c1 == i, (c0, c1, i)
    wt = w.detach().permute(0, 2, 3, 1).reshape(o, kh * kw, i).float()  # [O][tap][I]
    parts = []
    for lo, n in ((0, c0), (c0, c1)):
        if n == 0:
            continue
        blk = torch.zeros(o, kh * kw, ceil64(n), dtype=torch.float32, device=w.device)
        blk[:, :, :n] = wt[:, :, lo:lo + n]
        parts.append(blk)
    return torch.cat(parts, dim=2).reshape(o, -1).to(ACT).contiguous()


def pack_geglu(w: torch.Tensor, b: torch.Tensor, bn: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Interleave GEGLU projection rows per N tile: tile t holds x rows [t*bn/2, (t+1)*bn/2) then the matching gates.

    Reference: GEGLU.forward chunks proj(x) into (x, gate) halves (ldm/modules/attention.py:59-60,94-96).
    """
    two_inner, _ = w.shape
    inner = two_inner // 2
    half = bn // 2
    assert inner % half == 0, (inner, bn)
    idx = []
    for t in range(inner // half):
        idx.extend(range(t * half, (t + 1) * half))
        idx.extend(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx, device=w.device)
    return w.detach()[idx].contiguous(), b.detach()[idx].float().contiguous()


# ------------------------------------------------------------------------------------------------------------------
# tiling: decided INSIDE the library (csrc/plan.cu: cb_igemm_plan) -- nothing here chooses a tile
# ------------------------------------------------------------------------------------------------------------------
import os as _os
GEGLU_BN = int(_os.environ.get("CB_GEGLU_BN", "256"))   # N tile of the fused GEGLU projection: fixes the weight-row order
# read by the library itself (plan.cu); mirrored here only so tools / tests can show and pin them
GN_FUSE = int(_os.environ.get("CB_GN_FUSE", "1"))
GN_FUSE_MIN_K_CHUNKS = int(_os.environ.get("CB_GN_FUSE_MIN_K_CHUNKS", "1"))
GN_FUSE_MIN_BYTES = int(_os.environ.get("CB_GN_FUSE_MIN_BYTES", "0"))


def choose_tile(n: int, h: int, w: int) -> Tuple[int, int, int]:
    """The library's 128-row tile {tw, th, tn} over an (n, h, w) pixel grid (cb_igemm_plan)."""
    d, plan = IGemmDesc(), _lib.IGemmPlan()
    d.n, d.h, d.w, d.c0, d.cout, d.taps = n, h, w, 64, 64, 1
    check(_lib.load().cb_igemm_plan(C.byref(d), C.byref(plan)), "cb_igemm_plan")
    return plan.tw, plan.th, plan.tn


TAPS_1X1 = ([0], [0], [0])
TAPS_3X3 = ([kw - 1 for kh in range(3) for kw in range(3)], [kh - 1 for kh in range(3) for kw in range(3)], [0] * 9)


def taps_3x3_stride2(n: int):
    """stride 2, pad 1 over the parity-split input [2*ph+pw][n][h/2][w/2][c]: input row 2*oy+kh-1."""
    par = {0: (1, -1), 1: (0, 0), 2: (1, 0)}  # k -> (parity, shift)
    dw, dh, dn = [], [], []
    for kh in range(3):
        for kw in range(3):
            ph, sh = par[kh]
            pw, sw = par[kw]
            dw.append(sw)
            dh.append(sh)
            dn.append((2 * ph + pw) * n)
    return dw, dh, dn


def taps_3x3_stride2_asym(n: int):
    """stride 2 after F.pad(x, (0, 1, 0, 1)) with padding 0 (the VAE encoder's Downsample, ldm/modules/
    diffusionmodules/model.py:79-84) over the parity-split input: input row 2*oy+kh, the extra zero row / column at the
    bottom / right is the TMA out-of-bounds fill of the parity plane."""
    par = {0: (0, 0), 1: (1, 0), 2: (0, 1)}  # k -> (parity, shift)
    dw, dh, dn = [], [], []
    for kh in range(3):
        for kw in range(3):
            ph, sh = par[kh]
            pw, sw = par[kw]
            dw.append(sw)
            dh.append(sh)
            dn.append((2 * ph + pw) * n)
    return dw, dh, dn


# ------------------------------------------------------------------------------------------------------------------
# implicit GEMM
# ------------------------------------------------------------------------------------------------------------------
def igemm(a0: torch.Tensor, wgt: torch.Tensor, cout: int, *, a1: Optional[torch.Tensor] = None,
          out_grid: Optional[Tuple[int, int, int]] = None, taps=TAPS_1X1, bias: Optional[torch.Tensor] = None,
          rowbias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: int = ACT_NONE,
          mode: int = EPI_LINEAR, out: Optional[torch.Tensor] = None, out_f32: bool = False, out_ld: Optional[int] = None,
          out_scale: float = 1.0, heads: Optional[Tuple[int, int, int, int, int]] = None, bn: Optional[int] = None,
          stages: int = 0, epilogue: int = 0, pair: Optional[bool] = None, nsub: int = 0,
          ksplit: Optional[int] = None, gn_stats: bool = False,
          out_pixel_strides: Optional[Tuple[int, int, int]] = None,
          gn_table: Optional[Tuple[torch.Tensor, int]] = None, algo_flops: Optional[float] = None,
          ln_stats: bool = False, ln_in: Optional["LnFold"] = None) -> torch.Tensor:
    """D = A (*) W with fused epilogue. a0/a1: NHWC bf16 [N,H,W,C] (or [M,K]); wgt: packed by pack_weight.

    out_grid: (n, h, w) of the output pixel grid if it differs from a0's (stride-2 parity input).
    heads: (d, dpad, n_heads, tokens_per_batch, which_stride) for EPI_HEADS (then `out` must be given).
    out_pixel_strides: (w, h, n) element strides of the output pixel grid inside a larger NHWC tensor (`out` = a strided
    view's first element; staged epilogue, no residual) -- used by `conv3x3_up2x`.
    gn_table: (shared partial table [n, rows, 2, cout/2], first row of this launch) -- see conv3x3_up2x.
    gn_stats: the output feeds a GroupNorm -- when the launch qualifies (see `_gn_fusable`) its epilogue also writes
    per-block partial statistics, attached to the returned tensor as `_gn_part` for `groupnorm` to pick up.
    ln_stats: the output is a residual stream that a LayerNorm will read -- when the launch qualifies (cb_igemm_plan:
    ln_out_slots > 0) its epilogue also writes per-row partial sums, attached to the returned tensor as `_ln_part`.
    ln_in: `LnFold` -- a0 is the UN-normalised row matrix and `wgt` / `bias` were folded with the LayerNorm's gamma /
    beta (`fold_layernorm`); the epilogue applies mean / rstd from the producer's `_ln_part`.  The caller checks
    `ln_foldable(...)` first; a launch that cannot fold raises.
    """
    _need_cuda(a0, a1, wgt, bias, rowbias, residual, out)
    assert a0.dtype == ACT and wgt.dtype == ACT and a0.is_contiguous() and wgt.is_contiguous()
    if a0.dim() == 2:
        a_n, a_h, a_w, c0 = 1, 1, a0.shape[0], a0.shape[1]
    else:
        a_n, a_h, a_w, c0 = a0.shape
    c1 = 0
    if a1 is not None:
        assert a1.dtype == ACT and a1.is_contiguous() and tuple(a1.shape[:-1]) == tuple(a0.shape[:-1])
        c1 = a1.shape[-1]
    n, h, w = out_grid if out_grid is not None else (a_n, a_h, a_w)
    rows = n * h * w
    ncols = 2 * cout if mode == EPI_GEGLU else cout
    dw, dh, dn = taps
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(-1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if residual is not None:
        assert residual.dtype == ACT and residual.is_contiguous()
    want_f32 = out_f32 if out is None else (out.dtype == torch.float32)

    # the PROBLEM; the tiling fields stay 0 (= the library chooses) unless the caller pins them (tests, tools)
    d = IGemmDesc()
    d.a0, d.c0, d.a0_ld = _p(a0), c0, 0
    d.a1, d.c1, d.a1_ld = _p(a1), c1, 0
    d.a_n, d.a_h, d.a_w = a_n, a_h, a_w
    d.n, d.h, d.w = n, h, w
    d.taps = len(dw)
    for i in range(len(dw)):
        d.tap_dw[i], d.tap_dh[i], d.tap_dn[i] = dw[i], dh[i], dn[i]
    d.wgt, d.wgt_rows = _p(wgt), wgt.shape[0]
    d.cout = cout
    d.mode, d.act = mode, act
    d.out_f32 = int(want_f32)
    d.out_scale = out_scale
    d.out_ld = (out_ld if out_ld is not None else (cout if out is None else out.shape[-1])) if mode != EPI_HEADS else 0
    if heads is not None:
        d.heads_d, d.heads_dpad, d.heads_h, d.heads_tokens, d.heads_which_stride = heads
    d.bn, d.stages, d.epilogue, d.nsub = (bn or 0), stages, epilogue, nsub
    d.cta_pair = 0 if pair is None else (1 if pair else -1)
    d.ksplit = 0 if ksplit is None else ksplit
    if out_pixel_strides is not None:
        assert residual is None and mode == EPI_LINEAR and not want_f32
        d.out_w_stride, d.out_h_stride, d.out_n_stride = (int(v) for v in out_pixel_strides)
        d.epilogue = 2   # CB_EPILOGUE_STAGED: the strides live in the TMA-store tensor map
    plan = _lib.IGemmPlan()
    check(_lib.load().cb_igemm_plan(C.byref(d), C.byref(plan)), "cb_igemm_plan")
    if ln_stats and LN_FUSE and LN_STAGE_BIG_K and plan.ln_out_slots == 0 and d.epilogue == 0 and plan.ksplit == 1:
        # a large-K producer of the residual stream (ff.net.2 of a 1280-wide block: K = 5120) would take the direct
        # epilogue, which carries no row statistics, and the next block's norm1 would fall back to a stand-alone
        # LayerNorm launch: pin the staged epilogue when that makes the launch a statistics producer
        d.epilogue = 2
        plan2 = _lib.IGemmPlan()
        check(_lib.load().cb_igemm_plan(C.byref(d), C.byref(plan2)), "cb_igemm_plan")
        if plan2.ln_out_slots > 0 and plan2.ksplit == 1:
            plan = plan2
        else:
            d.epilogue = 0
    d.tw, d.th, d.tn = plan.tw, plan.th, plan.tn
    bn, pair, ksplit = plan.bn, bool(plan.cta_pair), plan.ksplit
    d.bn, d.cta_pair, d.nsub, d.ksplit = plan.bn, plan.cta_pair, plan.nsub, plan.ksplit

    gn_part = None
    if gn_stats and plan.gn_fusable and out is None and out_ld is None:
        gn_part = torch.empty((n, plan.gn_rows_per_image, 2, cout // 2), dtype=torch.float32, device=a0.device)
    ln_part = None
    if ln_stats and plan.ln_out_slots > 0 and LN_FUSE:
        ln_part = torch.empty((rows, plan.ln_out_slots, 2), dtype=torch.float32, device=a0.device)
        d.ln_partials_out = _p(ln_part)
    if ln_in is not None:
        if not plan.ln_foldable:
            raise ValueError("igemm: this launch cannot fold a LayerNorm (needs the staged plain / GEGLU epilogue)")
        assert ln_in.part.shape[0] == rows and bias is not None
        d.ln_partials_in, d.ln_in_slots, d.ln_dim, d.ln_eps = _p(ln_in.part), ln_in.part.shape[1], ln_in.dim, ln_in.eps
    if out is None:
        ld = out_ld if out_ld is not None else cout
        out = torch.empty((rows, ld), dtype=torch.float32 if out_f32 else ACT, device=a0.device)
    ld = out_ld if out_ld is not None else (out.shape[-1] if mode != EPI_HEADS else 0)

    final = None
    if ksplit > 1:   # raw fp32 partials now, epilogue in cb_splitk_reduce (what cb_igemm_auto does for a non-Python host)
        final = (out, bias, rowbias, residual, ld)
        out = torch.empty((plan.workspace_bytes // 4,), dtype=torch.float32, device=a0.device)
        bias = rowbias = residual = None
        ld = cout
    d.bias = _p(bias)
    d.rowbias, d.rowbias_ld = _p(rowbias), (rowbias.stride(0) if rowbias is not None else 0)
    d.residual, d.res_ld = _p(residual), (residual.shape[-1] if residual is not None else 0)
    d.out, d.out_ld, d.out_f32 = _p(out), ld, int(out.dtype == torch.float32)
    d.gn_partials = _p(gn_part)
    if gn_table is not None:
        assert gn_part is None and ksplit == 1 and (plan.tw * plan.th) % 32 == 0
        d.gn_partials, d.gn_rows_per_image, d.gn_row_offset = _p(gn_table[0]), gn_table[0].shape[1], gn_table[1]
    # algorithmic work of the reference op: 2 * rows * (taps * cin) * cout (GEGLU projects to 2 * cout columns)
    _launch("cb_igemm", lambda: _lib.load().cb_igemm(C.byref(d), _stream()),
            flops=2.0 * rows * len(dw) * (c0 + c1) * ncols if algo_flops is None else algo_flops,
            tag=f"M={rows} K={len(dw) * (c0 + c1)} N={ncols} taps={len(dw)} bn={bn}{'x2' if pair else ''}{f'/k{ksplit}' if ksplit > 1 else ''} epi={mode}"
                f"{'+res' if residual is not None else ''}{'+rowb' if rowbias is not None else ''}"
                f"{'+act' if act else ''}{'+f32' if out_f32 else ''}" if _PROF is not None else "")
    if final is not None:
        part = out
        out, bias, rowbias, residual, ld = final
        _launch("cb_splitk_reduce", lambda: _lib.load().cb_splitk_reduce(
            _p(part), ksplit, rows, cout, cout, _p(bias), _p(rowbias), rowbias.stride(0) if rowbias is not None else 0,
            h * w, _p(residual), residual.shape[-1] if residual is not None else 0, _p(out), ld, _stream()),
            nbytes=4.0 * ksplit * rows * cout + 2.0 * rows * cout)
    if gn_part is not None:
        out._gn_part = gn_part
    if ln_part is not None:
        out._ln_part = ln_part
    return out


# ------------------------------------------------------------------------------------------------------------------
# fused LayerNorm: statistics from the producer's epilogue, gamma / beta folded into the consumer's weights
# ------------------------------------------------------------------------------------------------------------------
LN_FUSE = int(_os.environ.get("CB_LN_FUSE", "1"))   # 0: stand-alone cb_layernorm launches (A/B)
LN_STAGE_BIG_K = int(_os.environ.get("CB_LN_STAGE_BIG_K", "1"))   # 0: large-K residual-stream producers keep the direct epilogue


class LnFold:
    """What a consumer GEMM needs to apply LayerNorm(x) on the fly: the producer's per-row partial sums of x, the row
    width and eps (the affine part and the centring live in the folded weights, `fold_layernorm`)."""

    def __init__(self, part: torch.Tensor, dim: int, eps: float):
        self.part, self.dim, self.eps = part, int(dim), float(eps)


def fold_layernorm(w: torch.Tensor, b: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """[O, I] linear weight (fp32) + LayerNorm affine -> (W'', b'):  W' = W diag(gamma),  W'' = W' with every ROW CENTRED
    (W''[n, :] -= mean(W'[n, :])),  b' = b + W beta.  Then  LN(x) W^T + b = rstd * (x W''^T) + b'  because
    (x - mean(x) 1) W'^T = x W'^T - mean(x) rowsum(W') = x W''^T   (ldm/modules/attention.py:900-912).  What remains of
    the mean after rounding W'' to 16 bits is mean(x) * (sum of a row's rounding errors): below the rounding of a
    stand-alone LayerNorm's 16-bit output (checked in tests/test_gpu_igemm.py)."""
    w = w.detach().float()
    wf = w * gamma.detach().float()[None, :]
    wf = wf - wf.mean(dim=1, keepdim=True)
    bf = w @ beta.detach().float()
    if b is not None:
        bf = bf + b.detach().float()
    return wf.contiguous(), bf.contiguous()


# nearest-2x upsample + conv3x3 (pad 1) folded: output pixel (2y+a, 2x+b) only sees the 2x2 low-resolution pixels
# y + DY[a], x + DX[b], and every original tap (ky, kx) that lands on the same low-resolution pixel adds its weight:
# four 2x2 convs over the LOW-resolution tensor, 16 instead of 36 multiply-adds per input pixel and channel pair, and
# the 4x larger upsampled tensor never exists (reference: F.interpolate(scale_factor=2, mode="nearest") + conv,
# openaimodel.py:113-123, model.py:60-64).
_UP2X_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}   # parity -> ((low-res offset, original taps), ...)


def pack_weight_up2x(w: torch.Tensor):
    """OIHW 3x3 weight -> [(a, b, packed 2x2 weight, taps)] for the four output parity classes."""
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    w = w.detach().float()
    out = []
    for a in (0, 1):
        for b in (0, 1):
            k = torch.zeros(w.shape[0], w.shape[1], 2, 2, dtype=torch.float32, device=w.device)
            dws, dhs = [], []
            for iy, (dy, kys) in enumerate(_UP2X_ROWS[a]):
                for ix, (dx, kxs) in enumerate(_UP2X_ROWS[b]):
                    k[:, :, iy, ix] = sum(w[:, :, ky, kx] for ky in kys for kx in kxs)
                    dws.append(dx)
                    dhs.append(dy)
            out.append((a, b, pack_weight(k), (dws, dhs, [0, 0, 0, 0])))
    return out


def conv3x3_up2x(x: torch.Tensor, packed, cout: int, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """conv3x3(nearest_upsample_2x(x)) for NHWC 16-bit x [N,H,W,C] -> [N,2H,2W,cout] by four 2x2 convs over x, each
    writing its parity class of the output through a strided TMA-store tensor map."""
    n, h, w, _ = x.shape
    out = torch.empty((n, 2 * h, 2 * w, cout), dtype=ACT, device=x.device)
    # fused GroupNorm statistics: the four launches fill disjoint row ranges of ONE partial table of the output
    tw, th, _ = choose_tile(n, h, w)
    bpi = (-(-w // tw)) * (-(-h // th))
    table = None
    if GN_FUSE and cout % 8 == 0 and (tw * th) % 32 == 0 and 2 * out.numel() >= GN_FUSE_MIN_BYTES:
        table = torch.empty((n, 4 * bpi, 2, cout // 2), dtype=torch.float32, device=x.device)
    for i, (a, b, wp, taps) in enumerate(packed):
        view = out[:, a::2, b::2, :]
        igemm(x, wp, cout, taps=taps, bias=bias, out=view, out_ld=cout,
              out_pixel_strides=(view.stride(2), view.stride(1), view.stride(0)),
              gn_table=None if table is None else (table, i * bpi),
              # profile bookkeeping: the algorithmic work is the REFERENCE op's (conv3x3 over the 2H x 2W grid,
              # SURVEY 8d); each of the four launches is credited a quarter of it although it executes 4/9 of that
              algo_flops=2.0 * (n * h * w) * 9 * x.shape[-1] * cout)
    if table is not None:
        out._gn_part = table
    return out


def nhwc(t: torch.Tensor, n: int, h: int, w: int, c: int) -> torch.Tensor:
    """[rows, c] -> [n, h, w, c] view that keeps the fused GroupNorm partials of the producing launch attached."""
    v = t.view(n, h, w, c)
    part = getattr(t, "_gn_part", None)
    if part is not None:
        v._gn_part = part
    return v


# ------------------------------------------------------------------------------------------------------------------
# attention / norms
# ------------------------------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, batch: int, heads: int, nq: int, nk: int, d: int,
              scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: 16-bit [batch*nq, heads*d] (a column slice of a wider projection output is fine: row stride = q.stride(0));
    k, v: [batch*nk, heads*d] likewise -> out 16-bit [batch*nq, heads*d]."""
    _need_cuda(q, k, v)
    for t, n in ((q, nq), (k, nk), (v, nk)):
        assert t.dtype == ACT and t.dim() == 2 and t.stride(1) == 1 and t.shape == (batch * n, heads * d), (t.shape, t.stride())
    if out is None:
        out = torch.empty((batch * nq, heads * d), dtype=ACT, device=q.device)
    _launch("cb_attention", lambda: _lib.load().cb_attention(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                                                             batch, heads, nq, nk, d, scale, _stream()),
            flops=4.0 * batch * heads * nq * nk * d, tag=f"bh={batch * heads} nq={nq} nk={nk} d={d}")
    return out


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(scale * s) over the last dim of a 2-D fp32 or bf16 tensor -> bf16 `out` (in place for bf16 if None)."""
    _need_cuda(s, out)
    assert s.dim() == 2 and s.stride(1) == 1 and s.dtype in (ACT, torch.float32)
    if out is None:
        out = s if s.dtype == ACT else torch.empty(s.shape, dtype=ACT, device=s.device)
    assert out.dtype == ACT and out.shape == s.shape and out.stride(1) == 1
    _launch("cb_softmax_rows", lambda: _lib.load().cb_softmax_rows(_p(s), int(s.dtype == torch.float32), s.stride(0), _p(out), out.stride(0),
                                      s.shape[0], s.shape[1], scale, _stream()))
    return out


def softmax_rows_(s: torch.Tensor, scale: float) -> torch.Tensor:
    return softmax_rows(s, scale, None)


def pointwise_nchw_to_nhwc(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], c_pad: int,
                           scale: float = 1.0) -> torch.Tensor:
    """1x1 channel mix of an NCHW fp32 tensor fused with the NHWC bf16 conversion (w fp32 [cout, c])."""
    _need_cuda(x, w, b)
    x = x.contiguous().float()
    n, c, h, wd = x.shape
    out = torch.empty((n, h, wd, c_pad), dtype=ACT, device=x.device)
    _launch("cb_pointwise_nchw_to_nhwc", lambda: _lib.load().cb_pointwise_nchw_to_nhwc(_p(x), n, c, h * wd, _p(w), _p(b), w.shape[0], c_pad, scale, _p(out),
                                                _stream()))
    return out


def groupnorm(x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
              x1: Optional[torch.Tensor] = None, groups: int = 32) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC bf16 [N,H,W,C0] (+ [N,H,W,C1] concatenated on channels) -> [N,H,W,C0+C1]."""
    _need_cuda(x0, x1, gamma, beta)
    assert x0.dtype == ACT and x0.is_contiguous() and gamma.dtype == torch.float32
    n = x0.shape[0]
    hw = x0.numel() // (n * x0.shape[-1])
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    out = torch.empty((*x0.shape[:-1], c0 + c1), dtype=ACT, device=x0.device)
    part0 = getattr(x0, "_gn_part", None)
    part1 = getattr(x1, "_gn_part", None) if x1 is not None else None
    if (part0 is not None and (x1 is None or part1 is not None) and ((c0 + c1) // groups) % 2 == 0 and n <= 65535
            and c0 + c1 <= 2560 and groups <= 64):
        # statistics came with the producers' epilogues: fold + one streaming pass; the workspace holds the partial
        # tables reduced to <= 32 rows per image (only written when a producer has more than 32 tiles per image)
        need = max(int(part0.shape[1]), 0 if part1 is None else int(part1.shape[1])) > 32
        stats = torch.empty((n * 32 * (c0 + c1) if need else 1,), dtype=torch.float32, device=x0.device)
        _launch("cb_groupnorm_from_partials", lambda: _lib.load().cb_groupnorm_from_partials(
            _p(x0), c0, _p(part0), part0.shape[1], _p(x1), c1, _p(part1), 0 if part1 is None else part1.shape[1], n, hw,
            groups, eps, _p(gamma), _p(beta), int(silu), _p(out), _p(stats), _stream()),
            nbytes=4.0 * n * hw * (c0 + c1), tag=f"n={n} hw={hw} c={c0}+{c1}")
        return out
    ws = int(_lib.load().cb_groupnorm_workspace_bytes(c0 + c1, n, hw, groups))
    if ws <= 0:
        raise ValueError(f"groupnorm: unsupported shape n={n} hw={hw} c={c0 + c1} groups={groups}")
    stats = torch.empty((ws // 4,), dtype=torch.float32, device=x0.device)
    _launch("cb_groupnorm_nhwc", lambda: _lib.load().cb_groupnorm_nhwc(_p(x0), c0, _p(x1), c1, n, hw, groups, eps, _p(gamma), _p(beta), int(silu),
                                        _p(out), _p(stats), _stream()),
            nbytes=4.0 * n * hw * (c0 + c1),  # algorithmic: one bf16 read + one bf16 write per element
            tag=f"n={n} hw={hw} c={c0}+{c1}")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _need_cuda(x, gamma, beta)
    assert x.dtype == ACT and x.is_contiguous()
    c = x.shape[-1]
    rows = x.numel() // c
    out = torch.empty_like(x)
    _launch("cb_layernorm", lambda: _lib.load().cb_layernorm(_p(x), rows, c, eps, _p(gamma), _p(beta), _p(out), _stream()),
            nbytes=4.0 * rows * c, tag=f"rows={rows} c={c}")
    return out


# ------------------------------------------------------------------------------------------------------------------
# layout / elementwise
# ------------------------------------------------------------------------------------------------------------------
_SRC_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def nchw_to_nhwc(x: torch.Tensor, c_pad: Optional[int] = None, scale: float = 1.0) -> torch.Tensor:
    _need_cuda(x)
    x = x.contiguous()
    n, c, h, w = x.shape
    c_pad = c if c_pad is None else c_pad
    out = torch.empty((n, h, w, c_pad), dtype=ACT, device=x.device)
    _launch("cb_nchw_to_nhwc", lambda: _lib.load().cb_nchw_to_nhwc(_p(x), _SRC_DTYPE[x.dtype], n, c, h * w, c_pad, scale, _p(out), _stream()))
    return out


def add_nchw(base: torch.Tensor, ctrl: torch.Tensor) -> torch.Tensor:
    """base NHWC 16-bit [N,H,W,C] + ctrl NCHW [N,C,H,W] (fp32 / fp16 / bf16) -> new NHWC tensor (ControlNet residuals,
    cldm/cldm.py:59-66)."""
    _need_cuda(base, ctrl)
    assert base.dtype == ACT and base.is_contiguous() and ctrl.dtype in _SRC_DTYPE
    n, h, w, c = base.shape
    if tuple(ctrl.shape) != (n, c, h, w):
        raise ValueError(f"control residual {tuple(ctrl.shape)} does not match the feature map {(n, c, h, w)}")
    ctrl = ctrl.contiguous()
    out = torch.empty_like(base)
    _launch("cb_add_nchw_to_nhwc", lambda: _lib.load().cb_add_nchw_to_nhwc(_p(base), _p(ctrl), _SRC_DTYPE[ctrl.dtype], n, c, h * w,
                                                                            _p(out), _stream()))
    return out


def nhwc_to_nchw_f32(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    """x: [N,H,W,C_ld] bf16 or fp32 -> fp32 [N,c,H,W]."""
    _need_cuda(x)
    assert x.is_contiguous()
    n, h, w, c_ld = x.shape
    c = c_ld if c is None else c
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _launch("cb_nhwc_to_nchw_f32", lambda: _lib.load().cb_nhwc_to_nchw_f32(_p(x), int(x.dtype == torch.float32), n, c, h * w, c_ld, _p(out), _stream()))
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=ACT, device=x.device)
    _launch("cb_upsample2x_nhwc", lambda: _lib.load().cb_upsample2x_nhwc(_p(x), n, h, w, c, _p(out), _stream()))
    return out


def parity_split(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((4, n, h // 2, w // 2, c), dtype=ACT, device=x.device)
    _launch("cb_parity_split_nhwc", lambda: _lib.load().cb_parity_split_nhwc(_p(x), n, h, w, c, _p(out), _stream()))
    return out


def timestep_embedding(t: torch.Tensor, dim: int, freqs: torch.Tensor) -> torch.Tensor:
    _need_cuda(t, freqs)
    assert t.dtype == torch.float32 and freqs.dtype == torch.float32 and freqs.numel() == dim // 2
    out = torch.empty((t.shape[0], dim), dtype=ACT, device=t.device)
    _launch("cb_timestep_embedding", lambda: _lib.load().cb_timestep_embedding(_p(t), t.shape[0], dim, _p(freqs), _p(out), _stream()))
    return out


def conv3x3_small_cin(x: torch.Tensor, cin: int, wgt: torch.Tensor, bias: Optional[torch.Tensor], cout: int) -> torch.Tensor:
    """x: NHWC bf16 [N,H,W,cin_ld]; wgt fp32 [3,3,cin,cout]."""
    _need_cuda(x, wgt, bias)
    n, h, w, cin_ld = x.shape
    out = torch.empty((n, h, w, cout), dtype=ACT, device=x.device)
    _launch("cb_conv3x3_small_cin", lambda: _lib.load().cb_conv3x3_small_cin(_p(x), n, h, w, cin, cin_ld, _p(wgt), _p(bias), cout, _p(out), _stream()))
    return out


def diag_gaussian(moments: torch.Tensor, noise: Optional[torch.Tensor] = None, scale: float = 1.0,
                  want_mean_std: bool = False):
    """moments fp32 NCHW [n, 2c, h, w] -> sample = scale * (mean + std * noise) (noise None: the mode) [, mean, std]."""
    _need_cuda(moments, noise)
    assert moments.dtype == torch.float32 and moments.is_contiguous() and moments.shape[1] % 2 == 0
    n, c2, h, w = moments.shape
    c = c2 // 2
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (n, c, h, w)
    sample = torch.empty((n, c, h, w), dtype=torch.float32, device=moments.device)
    mean = torch.empty_like(sample) if want_mean_std else None
    std = torch.empty_like(sample) if want_mean_std else None
    _launch("cb_diag_gaussian", lambda: _lib.load().cb_diag_gaussian(_p(moments), _p(noise), n, c, h * w, scale, _p(mean), _p(std),
                                                                     _p(sample), _stream()))
    return (sample, mean, std) if want_mean_std else sample


def silu_add(x: torch.Tensor, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, add)
    out = torch.empty_like(x)
    _launch("cb_silu_add", lambda: _lib.load().cb_silu_add(_p(x), _p(add), x.numel(), _p(out), _stream()))
    return out


def bilinear_upsample(x: torch.Tensor, factor: int) -> torch.Tensor:
    """F.interpolate(x, scale_factor=factor, mode='bilinear', align_corners=False) for fp32 NCHW latents."""
    _need_cuda(x)
    x = x.float().contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, c, h * factor, w * factor), dtype=torch.float32, device=x.device)
    _launch("cb_bilinear_upsample_f32", lambda: _lib.load().cb_bilinear_upsample_f32(_p(x), n * c, h, w, factor, _p(out), _stream()))
    return out


def image_to_u8(x: torch.Tensor) -> torch.Tensor:
    """x: fp32 NHWC [N,H,W,C_ld>=3] in [-1,1] -> uint8 [N,H,W,3]."""
    _need_cuda(x)
    n, h, w, c_ld = x.shape
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=x.device)
    _launch("cb_image_to_u8", lambda: _lib.load().cb_image_to_u8(_p(x), n, h * w, c_ld, _p(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------------------------
# sampler steps (fp32 NCHW latents)
# ------------------------------------------------------------------------------------------------------------------
def cfg_scale_input(x: torch.Tensor, c_in: float) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    b = x.shape[0]
    out = torch.empty((2 * b, *x.shape[1:]), dtype=torch.float32, device=x.device)
    _launch("cb_cfg_scale_input", lambda: _lib.load().cb_cfg_scale_input(_p(x), x.numel() // b, b, c_in, _p(out), _stream()))
    return out


def axpby(x: torch.Tensor, a: float, y: Optional[torch.Tensor] = None, b: float = 0.0) -> torch.Tensor:
    """a*x + b*y over fp32 tensors."""
    _need_cuda(x, y)
    assert x.dtype == torch.float32 and x.is_contiguous() and (y is None or (y.dtype == torch.float32 and y.is_contiguous()))
    out = torch.empty_like(x)
    _launch("cb_axpby_f32", lambda: _lib.load().cb_axpby_f32(_p(x), a, _p(y), b, x.numel(), _p(out), _stream()))
    return out


def cfg_mix(uncond: torch.Tensor, cond: torch.Tensor, scale: float) -> torch.Tensor:
    _need_cuda(uncond, cond)
    assert uncond.dtype == torch.float32 and cond.dtype == torch.float32 and uncond.is_contiguous() and cond.is_contiguous()
    out = torch.empty_like(uncond)
    _launch("cb_cfg_mix_f32", lambda: _lib.load().cb_cfg_mix_f32(_p(uncond), _p(cond), scale, uncond.numel(), _p(out), _stream()))
    return out


def blend_mask(keep: torch.Tensor, fresh: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """keep * mask + (1 - mask) * fresh over fp32 NCHW latents; mask [n, 1 or c, h, w] (ddim.py:171-174)."""
    _need_cuda(keep, fresh, mask)
    n, c, h, w = keep.shape
    keep, fresh = keep.float().contiguous(), fresh.float().contiguous()
    mask = mask.to(device=keep.device, dtype=torch.float32)
    if mask.dim() == 3:
        mask = mask[:, None]
    mask = mask.expand(n, mask.shape[1], h, w).contiguous()
    if mask.shape[1] not in (1, c):
        raise ValueError(f"mask channels {mask.shape[1]} must be 1 or {c}")
    out = torch.empty_like(keep)
    _launch("cb_blend_mask_f32", lambda: _lib.load().cb_blend_mask_f32(_p(keep), _p(fresh), _p(mask), n, c, h * w, mask.shape[1],
                                                                         _p(out), _stream()))
    return out


def _halves(eps2: torch.Tensor, b: int):
    assert eps2.dtype == torch.float32 and eps2.is_contiguous() and eps2.shape[0] == 2 * b
    return eps2[:b], eps2[b:]


def step_euler_ancestral(x, eps2, noise, cfg_scale, sigma, sigma_down, sigma_up, want_denoised=False, denoised=None):
    """Fused CFG + Euler-ancestral update. Pass `denoised` instead of `eps2` when the model call is opaque."""
    _need_cuda(x, eps2, noise, denoised)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if denoised is not None:
        eu, ec, isd = denoised.contiguous(), None, 1
    else:
        eu, ec = _halves(eps2, x.shape[0])
        isd = 0
    x_out = torch.empty_like(x)
    den = torch.empty_like(x) if want_denoised else None
    _launch("cb_step_euler_ancestral", lambda: _lib.load().cb_step_euler_ancestral(_p(x), _p(eu), _p(ec), isd, _p(noise), x.numel(), cfg_scale, sigma,
                                              sigma_down, sigma_up, _p(x_out), _p(den), _stream()))
    return x_out, den


def step_dpmpp_2m(x, eps2, old_denoised, cfg_scale, sigma, ratio, em1, c_new, c_old, denoised=None, want_denoised=True):
    _need_cuda(x, eps2, old_denoised, denoised)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if denoised is not None:
        eu, ec, isd = denoised.contiguous(), None, 1
    else:
        eu, ec = _halves(eps2, x.shape[0])
        isd = 0
    x_out = torch.empty_like(x)
    den = torch.empty_like(x) if want_denoised else None
    _launch("cb_step_dpmpp_2m", lambda: _lib.load().cb_step_dpmpp_2m(_p(x), _p(eu), _p(ec), isd, _p(old_denoised), x.numel(), cfg_scale, sigma, ratio,
                                       em1, c_new, c_old, _p(x_out), _p(den), _stream()))
    return x_out, den


def step_ddim(x, eps2, noise, cfg_scale, sqrt_at, sqrt_one_minus_at, sqrt_aprev, dir_coef, sigma_t, want_x0=True,
              eps=None):
    """Fused CFG + DDIM update; `eps` (no guidance) may be given instead of the doubled `eps2`."""
    _need_cuda(x, eps2, noise, eps)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if eps is not None:
        eu = ec = eps.contiguous()
        cfg_scale = 0.0
    else:
        eu, ec = _halves(eps2, x.shape[0])
    x_out = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    _launch("cb_step_ddim", lambda: _lib.load().cb_step_ddim(_p(x), _p(eu), _p(ec), _p(noise), x.numel(), cfg_scale, sqrt_at, sqrt_one_minus_at,
                                   sqrt_aprev, dir_coef, sigma_t, _p(x_out), _p(x0), _stream()))
    return x_out, x0
