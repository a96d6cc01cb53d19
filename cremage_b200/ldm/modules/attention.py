"""Drop-in mirror of the reference's `ldm.modules.attention` for the denoising path, running on the sm_100a kernels.

Same class names, constructor arguments, parameter (state-dict) names and forward signatures as
`modules/ldm/modules/attention.py` in HowToSD/cremage:
  CrossAttention (:265; CrossAttentionOriginal :537 and MemoryEfficientCrossAttention :696 are aliases -- all three
  compute softmax(q k^T / sqrt(d)) v), GEGLU (:57) / GEGLU_with_lora (:66), FeedForward (:119),
  BasicTransformerBlock (:862), SpatialTransformer (:915).
LoRA side branches (attention.py:79-96,148-168,306-376,966-1055) keep the reference's parameter names
(`q_lora_downs.{i}.weight`, `q_lora_ups.{i}.weight`, `q_lora_alphas.{i}`, ...) and are MERGED into the packed base
weight when the module packs:  W_eff = W + sum_i lora_weights[i] * (alpha_i / rank_i) * up_i @ down_i  -- the same
linear map (every branch is `up(down(x))` added to the base projection of the same input), at zero run-time cost.
IP-Adapter tokens (attention.py:338-341,354-358,448-521; CrossAttentionOriginal :623-627,660-681): with
`ipa_num_tokens = T > 0` the cross-attention (attn2 only, :895-901) owns `to_k_ipa` / `to_v_ipa`, the last T context
tokens go through them, a second attention with the same queries runs over those T tokens, and
`out + ipa_scale * out_ipa` enters to_out -- here as a second to_out GEMM with the weight pre-scaled by ipa_scale and
the first one's result as its residual (to_out is linear), so no extra elementwise pass exists.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ... import ops
from ...engine import PackedModule, f32, packw, require_cuda

GEGLU_BN = ops.GEGLU_BN  # N tile of the fused GEGLU projection (x / gate rows interleaved per 64 output columns)


def _check_extras(lora_ranks, ipa_num_tokens):
    if ipa_num_tokens is not None and int(ipa_num_tokens) < 0:
        raise ValueError("ipa_num_tokens must be >= 0")


def zero_init_module(m):
    with torch.no_grad():
        for p_ in m.parameters():
            p_.zero_()
    return m


class LoraBranches:
    """Mixin: reference-named LoRA parameter containers + the load-time merge."""

    def _init_lora(self, lora_ranks, lora_weights):
        self.lora_ranks = list(lora_ranks) if lora_ranks else []
        self.lora_weights = list(lora_weights) if lora_weights else [1.0] * len(self.lora_ranks)
        if len(self.lora_weights) != len(self.lora_ranks):
            raise ValueError("lora_weights and lora_ranks differ in length")

    def _add_lora(self, prefix: str, dim_in: int, dim_out: int, conv: bool = False):
        downs, ups, alphas = nn.ModuleList(), nn.ModuleList(), nn.ParameterList()
        for rank in self.lora_ranks:
            if conv:
                downs.append(zero_init_module(nn.Conv2d(dim_in, rank, kernel_size=1, stride=1, padding=0, bias=False)))
                ups.append(zero_init_module(nn.Conv2d(rank, dim_out, kernel_size=1, stride=1, padding=0, bias=False)))
            else:
                downs.append(zero_init_module(nn.Linear(dim_in, rank, bias=False)))
                ups.append(zero_init_module(nn.Linear(rank, dim_out, bias=False)))
            alphas.append(nn.Parameter(torch.tensor(float(rank))))
        setattr(self, prefix + "_lora_downs", downs)
        setattr(self, prefix + "_lora_ups", ups)
        setattr(self, prefix + "_lora_alphas", alphas)

    def _merged(self, base: torch.Tensor, prefix: str, device) -> torch.Tensor:
        """fp32 [out, in] base weight + the LoRA deltas of branch `prefix`."""
        w = base.detach().to(device=device, dtype=torch.float32).reshape(base.shape[0], -1).clone()
        downs, ups = getattr(self, prefix + "_lora_downs"), getattr(self, prefix + "_lora_ups")
        alphas = getattr(self, prefix + "_lora_alphas")
        for i, rank in enumerate(self.lora_ranks):
            up = ups[i].weight.detach().to(device=device, dtype=torch.float32).reshape(ups[i].weight.shape[0], -1)
            down = downs[i].weight.detach().to(device=device, dtype=torch.float32).reshape(rank, -1)
            scale = float(self.lora_weights[i]) * (alphas[i].detach().to(device=device, dtype=torch.float32) / float(rank))
            w += scale * (up @ down)
        return w


def Normalize(in_channels):
    """attention.py:189 -- GroupNorm(32, eps=1e-6, affine)."""
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class GEGLU(nn.Module, LoraBranches):
    """Parameter container for `ff.net.0.proj` (attention.py:57-62,66-96); computed fused inside FeedForward."""

    def __init__(self, dim_in, dim_out, lora_ranks: List[int] = None, lora_weights: List[float] = None):
        super().__init__()
        self._init_lora(lora_ranks, lora_weights)
        self.proj = nn.Linear(dim_in, dim_out * 2)
        self._add_lora("proj", dim_in, dim_out * 2)


GEGLU_with_lora = GEGLU


class FeedForward(PackedModule, LoraBranches):
    """attention.py:119-168: Linear(d, 8d) -> x * gelu(gate) -> Linear(4d, d). Only the gated (glu=True) form is on
    the path (BasicTransformerBlock passes gated_ff=True)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0., lora_ranks: List[int] = None,
                 lora_weights: List[float] = None):
        super().__init__()
        self._init_lora(lora_ranks, lora_weights)
        if not glu:
            raise NotImplementedError("cremage_b200: FeedForward without GEGLU is not on the SD path")
        inner_dim = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        self.dim, self.inner_dim, self.dim_out = dim, inner_dim, dim_out
        self.net = nn.ModuleList([GEGLU(dim, inner_dim, lora_ranks=lora_ranks, lora_weights=lora_weights),
                                  nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])
        self._add_lora("net_2", inner_dim, dim_out)

    def _pack(self, device):
        wq, bq = ops.pack_geglu(self.net[0]._merged(self.net[0].proj.weight, "proj", device),
                                self.net[0].proj.bias.to(device).float(), GEGLU_BN)
        return {"w1": ops.pack_weight(wq), "b1": bq.contiguous(),
                "w2": ops.pack_weight(self._merged(self.net[2].weight, "net_2", device)),
                "b2": f32(self.net[2].bias, device)}

    def _run(self, x2d: torch.Tensor, residual: Optional[torch.Tensor] = None, ln=None) -> torch.Tensor:
        """ln = (LnFold, folded packed w1, folded b1): x2d is the UN-normalised row matrix and the LayerNorm in front of
        the feed-forward is applied inside the GEGLU projection's epilogue (BasicTransformerBlock._run)."""
        p = self.packed(x2d.device)
        if ln is not None:
            h = ops.igemm(x2d, ln[1], self.inner_dim, bias=ln[2], mode=ops.EPI_GEGLU, bn=GEGLU_BN, ln_in=ln[0])
        else:
            h = ops.igemm(x2d, p["w1"], self.inner_dim, bias=p["b1"], mode=ops.EPI_GEGLU, bn=GEGLU_BN)
        return ops.igemm(h, p["w2"], self.dim_out, bias=p["b2"], residual=residual, ln_stats=residual is not None)

    def forward(self, x):
        require_cuda(x, "FeedForward.forward")
        shp = x.shape
        y = self._run(x.reshape(-1, shp[-1]).to(ops.ACT).contiguous())
        return y.view(*shp[:-1], self.dim_out).to(x.dtype)


class CrossAttention(PackedModule, LoraBranches):
    """attention.py:265-534 / :537-693 / :696-861. q/k/v projections write the per-head padded layout straight from
    the GEMM epilogue; the core is one fused flash-style kernel; to_out fuses bias (+ the block residual)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0., lora_ranks: List[int] = None,
                 lora_weights: List[float] = None, ipa_scale=1.0, ipa_num_tokens=0):
        super().__init__()
        _check_extras(lora_ranks, ipa_num_tokens)
        self._init_lora(lora_ranks, lora_weights)
        inner_dim = dim_head * heads
        self.is_self = context_dim is None
        context_dim = query_dim if context_dim is None else context_dim
        if dim_head % 8 != 0 or dim_head > 192:
            raise ValueError(f"cremage_b200: head dim {dim_head} unsupported (multiple of 8, <= 192)")
        self.query_dim, self.context_dim, self.inner_dim = query_dim, context_dim, inner_dim
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self._add_lora("q", query_dim, inner_dim)
        self._add_lora("k", context_dim, inner_dim)
        self._add_lora("v", context_dim, inner_dim)
        self._add_lora("out", inner_dim, query_dim)
        self.ipa_scale = ipa_scale
        self.ipa_num_tokens = int(ipa_num_tokens or 0)
        if self.ipa_num_tokens > 0:   # IP-Adapter FaceID support, attention.py:338-341
            self.to_k_ipa = nn.Linear(context_dim, inner_dim, bias=False)
            self.to_v_ipa = nn.Linear(context_dim, inner_dim, bias=False)

    def _pack(self, device):
        wq, wk, wv = (self._merged(getattr(self, "to_" + n).weight, n, device) for n in ("q", "k", "v"))
        wo = self._merged(self.to_out[0].weight, "out", device)
        p = {"wq": ops.pack_weight(wq), "wkv": ops.pack_weight(torch.cat([wk, wv], 0)),
             "wo": ops.pack_weight(wo), "bo": f32(self.to_out[0].bias, device)}
        if self.context_dim == self.query_dim:
            p["wqkv"] = ops.pack_weight(torch.cat([wq, wk, wv], 0))
        if self.ipa_num_tokens > 0:
            wk2 = self.to_k_ipa.weight.detach().to(device=device, dtype=torch.float32)
            wv2 = self.to_v_ipa.weight.detach().to(device=device, dtype=torch.float32)
            p["wkv_ipa"] = ops.pack_weight(torch.cat([wk2, wv2], 0))
            p["wo_ipa"] = ops.pack_weight(wo * float(self.ipa_scale))     # to_out(out + s * out_ipa) = to_out(out) + (s W) out_ipa
        return p

    def _run(self, x2d: torch.Tensor, batch: int, nq: int, ctx2d: Optional[torch.Tensor], nk: int,
             residual: Optional[torch.Tensor] = None, ln=None) -> torch.Tensor:
        """x2d: bf16 [batch*nq, query_dim]; ctx2d: bf16 [batch*nk, context_dim] or None for self-attention.
        ln = (LnFold, folded packed projection weight, folded bias): x2d is the UN-normalised row matrix and the
        LayerNorm in front of this attention is applied inside the q (self-attention: q | k | v) projection's epilogue.
        The output (to_out + residual) carries per-row partial sums for the next LayerNorm when the launch can."""
        p = self.packed(x2d.device)
        h, d, inner = self.heads, self.dim_head, self.inner_dim
        stats = residual is not None
        if ctx2d is None:
            if "wqkv" not in p:
                raise ValueError("self-attention requested on a CrossAttention built with a different context_dim")
            nk = nq
            if ln is not None:
                qkv = ops.igemm(x2d, ln[1], 3 * inner, bias=ln[2], ln_in=ln[0])
            else:
                qkv = ops.igemm(x2d, p["wqkv"], 3 * inner)         # [M, q | k | v]: read in place by the attention kernel
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
        else:
            pre = getattr(ctx2d, "_cb_kv", None)      # K/V of this context cached by the UNet across sampler steps
            kvs = pre.get(id(self)) if pre is not None else None
            if kvs is None:
                kvs = self.compute_kv(ctx2d, batch, nk)
            q = ops.igemm(x2d, ln[1], inner, bias=ln[2], ln_in=ln[0]) if ln is not None else ops.igemm(x2d, p["wq"], inner)
            if self.ipa_num_tokens > 0:
                t_ipa = self.ipa_num_tokens
                kv, kv2 = kvs
                a = ops.attention(q, kv[:, :inner], kv[:, inner:], batch, h, nq, nk - t_ipa, d, self.scale)
                a2 = ops.attention(q, kv2[:, :inner], kv2[:, inner:], batch, h, nq, t_ipa, d, self.scale)
                y = ops.igemm(a, p["wo"], self.query_dim, bias=p["bo"], residual=residual)
                return ops.igemm(a2, p["wo_ipa"], self.query_dim, residual=y, ln_stats=stats)
            kv = kvs[0]
            k, v = kv[:, :inner], kv[:, inner:]
        a = ops.attention(q, k, v, batch, h, nq, nk, d, self.scale)
        return ops.igemm(a, p["wo"], self.query_dim, bias=p["bo"], residual=residual, ln_stats=stats)

    def compute_kv(self, ctx2d: torch.Tensor, batch: int, nk: int, out=None):
        """K | V projections of the context: ([batch*nk, 2*inner],) -- with IP-Adapter tokens ([batch*(nk-T), 2*inner],
        [batch*T, 2*inner]) from to_k/to_v and to_k_ipa/to_v_ipa.  They depend on the context and the weights only, so the
        UNet computes them once per context and reuses them for every sampler step (`out`: buffers to overwrite)."""
        p = self.packed(ctx2d.device)
        inner = self.inner_dim
        if self.ipa_num_tokens == 0:
            return (ops.igemm(ctx2d, p["wkv"], 2 * inner, out=None if out is None else out[0]),)
        # the last T context tokens are the IP-Adapter's (attention.py:354-358); the split copies are made once per
        # context tensor and shared by every cross-attention of the forward
        t_ipa = self.ipa_num_tokens
        if nk <= t_ipa:
            raise ValueError(f"context has {nk} tokens, not more than ipa_num_tokens = {t_ipa}")
        split = getattr(ctx2d, "_cb_ipa_split", None)
        if split is None or split[0] != t_ipa:
            c3 = ctx2d.view(batch, nk, -1)
            split = (t_ipa, c3[:, :nk - t_ipa].reshape(batch * (nk - t_ipa), -1).contiguous(),
                     c3[:, nk - t_ipa:].reshape(batch * t_ipa, -1).contiguous())
            ctx2d._cb_ipa_split = split
        _, ctx_text, ctx_ipa = split
        return (ops.igemm(ctx_text, p["wkv"], 2 * inner, out=None if out is None else out[0]),
                ops.igemm(ctx_ipa, p["wkv_ipa"], 2 * inner, out=None if out is None else out[1]))

    def forward(self, x, context=None, mask=None):
        require_cuda(x, "CrossAttention.forward")
        if mask is not None:
            raise NotImplementedError("cremage_b200: attention masks are not used on the SD path")
        b, n, _ = x.shape
        x2d = x.reshape(b * n, -1).to(ops.ACT).contiguous()
        ctx2d, nk = None, n
        if context is not None:
            nk = context.shape[1]
            ctx2d = context.reshape(b * nk, -1).to(ops.ACT).contiguous()
        elif not self.is_self and self.context_dim != self.query_dim:
            raise ValueError("context is required for this CrossAttention")
        return self._run(x2d, b, n, ctx2d, nk).view(b, n, self.query_dim).to(x.dtype)


CrossAttentionOriginal = CrossAttention
MemoryEfficientCrossAttention = CrossAttention


class BasicTransformerBlock(PackedModule):
    """attention.py:862-912: x = attn1(LN1 x) + x; x = attn2(LN2 x, ctx) + x; x = ff(LN3 x) + x (LayerNorm eps 1e-5)."""

    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False, lora_ranks: List[int] = None, lora_weights: List[float] = None, ipa_scale=1.0,
                 ipa_num_tokens=0):
        super().__init__()
        _check_extras(lora_ranks, ipa_num_tokens)
        self.disable_self_attn = disable_self_attn
        lk = dict(lora_ranks=lora_ranks, lora_weights=lora_weights)
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                                    context_dim=context_dim if self.disable_self_attn else None, **lk)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff, **lk)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout, ipa_scale=ipa_scale, ipa_num_tokens=ipa_num_tokens, **lk)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint
        self.dim = dim

    def _own_params(self):
        # the folded packs below depend on the projections' weights (and their LoRA branches) too
        return list(self.parameters())

    def _pack(self, device):
        """LayerNorm parameters, plus -- for the fused form -- the three consumer projections with gamma folded into the
        weights (rows centred) and beta into the bias (ops.fold_layernorm): attn1 q|k|v, attn2 q, the GEGLU projection."""
        p = {f"n{i}{k[0]}": f32(getattr(getattr(self, f"norm{i}"), k), device)
             for i in (1, 2, 3) for k in ("weight", "bias")}
        a1, a2, ff = self.attn1, self.attn2, self.ff
        if not self.disable_self_attn and a1.context_dim == a1.query_dim:
            w = torch.cat([a1._merged(getattr(a1, "to_" + n).weight, n, device) for n in ("q", "k", "v")], 0)
            wf, bf = ops.fold_layernorm(w, None, p["n1w"], p["n1b"])
            p["ln1"] = (ops.pack_weight(wf), bf)
        wf, bf = ops.fold_layernorm(a2._merged(a2.to_q.weight, "q", device), None, p["n2w"], p["n2b"])
        p["ln2"] = (ops.pack_weight(wf), bf)
        wf, bf = ops.fold_layernorm(ff.net[0]._merged(ff.net[0].proj.weight, "proj", device),
                                    ff.net[0].proj.bias.to(device), p["n3w"], p["n3b"])
        wq, bq = ops.pack_geglu(wf, bf, GEGLU_BN)          # x / gate rows interleaved per N tile, like FeedForward._pack
        p["ln3"] = (ops.pack_weight(wq), bq.contiguous())
        return p

    def _ln(self, x2d, p, i):
        """LnFold for norm{i} when the producer of x2d delivered the row statistics, else None."""
        part = getattr(x2d, "_ln_part", None)
        key = f"ln{i}"
        if part is None or key not in p or not ops.LN_FUSE:
            return None
        w, b = p[key]
        return (ops.LnFold(part, self.dim, getattr(self, f"norm{i}").eps), w, b)

    def _run(self, x2d: torch.Tensor, batch: int, n: int, ctx2d: Optional[torch.Tensor], nk: int) -> torch.Tensor:
        """Each LayerNorm is either folded into the GEMM that consumes it (the producer of x2d wrote per-row partial sums:
        no LayerNorm launch, the token matrix is read once less and written once less) or a stand-alone launch."""
        p = self.packed(x2d.device)
        ln = self._ln(x2d, p, 1)
        h = x2d if ln is not None else ops.layernorm(x2d, p["n1w"], p["n1b"], self.norm1.eps)
        if self.disable_self_attn:
            x2d = self.attn1._run(h, batch, n, ctx2d, nk, residual=x2d)
        else:
            x2d = self.attn1._run(h, batch, n, None, n, residual=x2d, ln=ln)
        ln = self._ln(x2d, p, 2)
        h = x2d if ln is not None else ops.layernorm(x2d, p["n2w"], p["n2b"], self.norm2.eps)
        x2d = self.attn2._run(h, batch, n, ctx2d, nk, residual=x2d, ln=ln)
        ln = self._ln(x2d, p, 3)
        h = x2d if ln is not None else ops.layernorm(x2d, p["n3w"], p["n3b"], self.norm3.eps)
        return self.ff._run(h, residual=x2d, ln=ln)

    def forward(self, x, context=None):
        require_cuda(x, "BasicTransformerBlock.forward")
        b, n, c = x.shape
        x2d = x.reshape(b * n, c).to(ops.ACT).contiguous()
        ctx2d, nk = None, n
        if context is not None:
            nk = context.shape[1]
            ctx2d = context.reshape(b * nk, -1).to(ops.ACT).contiguous()
        return self._run(x2d, b, n, ctx2d, nk).view(b, n, c).to(x.dtype)


class SpatialTransformer(PackedModule, LoraBranches):
    _HONOR_USE_LINEAR = False  # the ldm class ignores `use_linear`; the sgm mirror (SDXL) honours it

    """attention.py:915-1057. GroupNorm(eps 1e-6) -> 1x1 proj_in -> transformer blocks on [b, hw, c] -> 1x1 proj_out
    -> + x_in.  In NHWC the two rearranges are free and the 1x1 convs are plain GEMMs. `use_linear` is accepted and
    ignored exactly like the reference (docstring at attention.py:930-945)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=True, lora_ranks: List[int] = None, lora_weights: List[float] = None,
                 ipa_scale=1.0, ipa_num_tokens=0):
        super().__init__()
        _check_extras(lora_ranks, ipa_num_tokens)
        self._init_lora(lora_ranks, lora_weights)
        if context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = [context_dim]
        if context_dim is None:
            context_dim = [None] * depth
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.inner_dim = inner_dim
        self.norm = Normalize(in_channels)
        self.use_linear = bool(use_linear) and self._HONOR_USE_LINEAR
        if self.use_linear:  # sgm/modules/attention.py:999,1056: nn.Linear on the [b, hw, c] view -- the same GEMM
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                  disable_self_attn=disable_self_attn, checkpoint=use_checkpoint,
                                  lora_ranks=self.lora_ranks, lora_weights=self.lora_weights,
                                  ipa_scale=ipa_scale, ipa_num_tokens=ipa_num_tokens)
            for d in range(depth)])
        if self.use_linear:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
        with torch.no_grad():  # zero_module, attention.py:1002
            self.proj_out.weight.zero_()
            self.proj_out.bias.zero_()
        self._add_lora("proj_in", in_channels, inner_dim, conv=not self.use_linear)
        self._add_lora("proj_out", inner_dim, in_channels, conv=not self.use_linear)

    def _own_params(self):
        ps = [self.norm.weight, self.norm.bias, self.proj_in.weight, self.proj_in.bias, self.proj_out.weight,
              self.proj_out.bias]
        for pre in ("proj_in", "proj_out"):
            for kind in ("downs", "ups", "alphas"):
                ps += list(getattr(self, f"{pre}_lora_{kind}").parameters())
        return ps

    def _pack(self, device):
        return {"ng": f32(self.norm.weight, device), "nb": f32(self.norm.bias, device),
                "wi": ops.pack_weight(self._merged(self.proj_in.weight, "proj_in", device)),
                "bi": f32(self.proj_in.bias, device),
                "wo": ops.pack_weight(self._merged(self.proj_out.weight, "proj_out", device)),
                "bo": f32(self.proj_out.bias, device)}

    def _run(self, x: torch.Tensor, ctx2d: Optional[torch.Tensor], nk: int) -> torch.Tensor:
        """x: NHWC bf16 [b, h, w, c] -> same shape."""
        p = self.packed(x.device)
        b, hh, ww, c = x.shape
        n = hh * ww
        g = ops.groupnorm(x, p["ng"], p["nb"], self.norm.eps, silu=False)
        h2d = ops.igemm(g.view(b * n, c), p["wi"], self.inner_dim, bias=p["bi"], ln_stats=True)
        for blk in self.transformer_blocks:
            h2d = blk._run(h2d, b, n, ctx2d, nk)
        # rows as a (b, 1, n) pixel grid so that the fused GroupNorm statistics of the output are per image
        out = ops.igemm(h2d.view(b, 1, n, h2d.shape[-1]), p["wo"], c, bias=p["bo"], residual=x.view(b * n, c), gn_stats=True)
        return ops.nhwc(out, b, hh, ww, c)

    def forward(self, x, context=None):
        require_cuda(x, "SpatialTransformer.forward")
        ctx2d, nk = None, 0
        if context is not None:
            nk = context.shape[1]
            ctx2d = context.reshape(-1, context.shape[-1]).to(ops.ACT).contiguous()
        y = self._run(ops.nchw_to_nhwc(x), ctx2d, nk)
        return ops.nhwc_to_nchw_f32(y).to(x.dtype)
