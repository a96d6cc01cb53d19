"""Shared host-side machinery of the drop-in modules: lazy weight packing, zero-padded attention workspaces and
CUDA-graph capture of a module call.  No arithmetic happens here."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

# Bumped whenever ANY PackedModule's parameters may have been replaced (`.to()/.half()/.cuda()/.cpu()` on the module or on
# a sub-tree, `load_state_dict`).  A root module that owns CUDA graphs over its whole tree compares it on every call: a
# sub-tree `_apply` (UNetModel.convert_to_fp16 -> blocks.half(), openaimodel.py:758-764) never reaches the root's own
# `_apply`, and neither `.half()` nor `p.data = ...` bumps `Parameter._version`.
PACK_EPOCH = 0


def _bump_epoch():
    global PACK_EPOCH
    PACK_EPOCH += 1


def param_fingerprint(params):
    """(storage address, dtype, version) of every parameter: changes on .to()/.half()/`p.data = ...`/in-place updates
    through autograd-visible ops.  (An edit through `p.data.<op>_()` is invisible to torch itself: call
    `invalidate_packed()` after such an edit.)"""
    return tuple((p.data_ptr(), p.dtype, p._version) for p in params)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: cremage_b200 runs hand-written sm_100a CUDA only and has no CPU fallback; "
            f"got a tensor on '{t.device}'. Move the module and its inputs to a CUDA device.")


class PackedModule(nn.Module):
    """nn.Module whose parameters (reference names, any float dtype) are repacked lazily into the bf16 / fp32 device
    layouts the kernels read.  The packed cache is dropped whenever the parameters may have changed: `.to()/.half()/
    .cuda()/.cpu()` (the reference's low_vram_shift moves the UNet between devices, ldm/models/diffusion/ddpm.py:1460-1498),
    `load_state_dict`, or an in-place update (tracked through the tensors' version counters)."""

    def __init__(self):
        super().__init__()
        self._cb_packed = None
        self._cb_key = None

    def _apply(self, fn, *args, **kwargs):
        self._cb_packed = None
        _bump_epoch()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._cb_packed = None
        _bump_epoch()
        return super()._load_from_state_dict(*args, **kwargs)

    def _pack(self, device: torch.device) -> dict:  # pragma: no cover - overridden
        raise NotImplementedError

    def _own_params(self):
        return list(self.parameters(recurse=True))

    def packed(self, device: torch.device) -> dict:
        key = (str(device), ops.ACT) + param_fingerprint(self._own_params())
        if self._cb_packed is None or self._cb_key != key:
            # one pack per 16-bit type: a module used under both builds of the library (ops.precision) keeps both
            cache = self.__dict__.setdefault("_cb_pack_cache", {})
            if self._cb_packed is None:
                cache.clear()
            hit = cache.get(key)
            if hit is None:
                with torch.no_grad():
                    hit = self._pack(device)
                for k in [k for k in cache if k[:2] == key[:2]]:
                    del cache[k]            # same device and dtype, older parameters
                cache[key] = hit
            self._cb_packed = hit
            self._cb_key = key
        return self._cb_packed

    def invalidate_packed(self):
        _bump_epoch()
        for m in self.modules():
            if isinstance(m, PackedModule):
                m._cb_packed = None


def f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def packw(w: torch.Tensor, device, splits=None) -> torch.Tensor:
    return ops.pack_weight(w.detach().to(device=device, dtype=torch.float32), splits)


# ----------------------------------------------------------------------------------------------------------------------
# persistent zero-initialised workspaces (per-head padded Q/K/V: the pad columns are never written and must stay zero)
# ----------------------------------------------------------------------------------------------------------------------
_WORKSPACES: Dict[Tuple, torch.Tensor] = {}


def zero_workspace(tag: str, shape: Tuple[int, ...], device: torch.device, init=None) -> torch.Tensor:
    """Persistent zero-filled buffer; `init(buf)` runs once at creation (e.g. the attention row-sum column)."""
    key = (tag, tuple(shape), str(device), ops.ACT)
    buf = _WORKSPACES.get(key)
    if buf is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("workspace allocation during CUDA-graph capture; run one eager call first")
        buf = torch.zeros(shape, dtype=ops.ACT, device=device)
        if init is not None:
            init(buf)
        _WORKSPACES[key] = buf
    return buf


def qkv_workspace(tag: str, which: int, v_index: int, bh: int, tokens: int, d: int, dpad: int,
                  device: torch.device) -> torch.Tensor:
    """[which][bh][tokens][dpad] per-head padded projection target. Pad columns stay zero except column d of the V
    plane, which holds 1.0 when dpad > d: cb_attention's tensor core then accumulates the softmax row sum for free."""
    def init(buf):
        if dpad > d:
            buf[v_index, :, :, d] = 1.0
    return zero_workspace(f"{tag}_d{d}", (which, bh, tokens, dpad), device, init)


def clear_workspaces():
    _WORKSPACES.clear()


def head_pad(d: int) -> int:
    """Head dim padded to whole 128-byte swizzle rows (64 bf16)."""
    return (d + 63) // 64 * 64


# ----------------------------------------------------------------------------------------------------------------------
# CUDA graph capture of a (tensor...) -> tensor function with static shapes
# ----------------------------------------------------------------------------------------------------------------------
GRAPH_REPLAYED_LAUNCHES = 0  # kernels (of this library) executed through CUDA-graph replays in this process


def total_launches() -> int:
    """Kernels launched by libcremage_b200 in this process: direct launches + those replayed inside CUDA graphs."""
    from . import _lib
    return _lib.launch_count() + GRAPH_REPLAYED_LAUNCHES


class GraphedCall:
    """Captures `fn(*tensors)` once per input signature and replays it. Inputs are copied into static buffers, the
    output is a static buffer (cloned on return unless `clone_output=False`)."""

    def __init__(self, fn: Callable, warmup: int = 2, clone_output: bool = True, keepalive: Optional[Callable] = None):
        self.fn = fn
        self.warmup = warmup
        self.clone_output = clone_output
        self.keepalive = keepalive      # () -> objects the captured kernels read (packed weights): held per graph
        self._graphs: Dict[Tuple, Tuple] = {}

    @staticmethod
    def _sig(tensors):
        return (ops.ACT,) + tuple((tuple(t.shape), t.dtype, str(t.device)) for t in tensors)

    def __call__(self, *tensors: torch.Tensor) -> torch.Tensor:
        sig = self._sig(tensors)
        entry = self._graphs.get(sig)
        if entry is None:
            static_in = [t.clone() for t in tensors]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):  # eager warm-up: packs weights, allocates workspaces
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import _lib
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            # No garbage collection while the stream is capturing: a collection that happens to destroy an unrelated
            # dead model's CUDA graphs / pooled memory calls cudaFree-class APIs, which invalidate a global-mode capture
            # (seen as error 901 in long test runs).  torch.cuda.graph() collects once on entry; keep it off until exit.
            import gc
            gc_was_enabled = gc.isenabled()
            gc.collect()
            gc.disable()
            try:
                with torch.cuda.graph(graph):
                    static_out = self.fn(*static_in)
            finally:
                if gc_was_enabled:
                    gc.enable()
            # the graph's kernels hold raw pointers: keep what they read alive for as long as the graph can be replayed
            held = self.keepalive() if self.keepalive is not None else None
            entry = (graph, static_in, static_out, _lib.launch_count() - before, held)
            self._graphs[sig] = entry
        graph, static_in, static_out, n_launches, _held = entry
        for dst, src in zip(static_in, tensors):
            dst.copy_(src)
        graph.replay()
        global GRAPH_REPLAYED_LAUNCHES
        GRAPH_REPLAYED_LAUNCHES += n_launches
        return static_out.clone() if self.clone_output else static_out

    def reset(self):
        self._graphs.clear()
