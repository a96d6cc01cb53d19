"""Drop-in mirror of the reference's `ldm.modules.diffusionmodules.openaimodel` (HowToSD/cremage
modules/ldm/modules/diffusionmodules/openaimodel.py) running on hand-written sm_100a kernels.

`UNetModel` keeps the reference constructor (:447-481), attribute names (`time_embed`, `input_blocks`, `middle_block`,
`output_blocks`, `out` -- ControlNet walks them, cldm/cldm.py:44-70), state-dict keys (pinned by
test/ldm/ldm_instantiation_test.py:23-26) and `forward(x, timesteps, context, y)` (:780).  Swapping the YAML `target:`
string is the whole integration (ldm/util.py:81-96).

Internally activations are NHWC bf16, every contraction is the tcgen05 implicit GEMM, the skip concat
(`th.cat([h, hs.pop()], 1)`, :808) is never materialised (GroupNorm and the GEMM read two sources), the 22 per-block
timestep projections (:222-228) run as ONE GEMM, and the per-image timestep bias / residual adds live in GEMM epilogues.
There is no CPU fallback: tensors on a non-CUDA device raise.
"""
from __future__ import annotations

import os
from abc import abstractmethod
from typing import Sequence, List, Optional

import torch
import torch.nn as nn

from .... import ops
from ....engine import GraphedCall, PackedModule, f32, packw, require_cuda
from ..attention import SpatialTransformer
from .util import conv_nd, linear, normalization, timestep_freqs, zero_module


def _ceil8(c: int) -> int:
    return (c + 7) // 8 * 8


class TimestepBlock(nn.Module):
    """openaimodel.py:60-69."""

    @abstractmethod
    def forward(self, x, emb):
        """Apply the module to `x` given `emb` timestep embeddings."""


class Upsample(PackedModule):
    """openaimodel.py:95-123: nearest 2x then conv3x3."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or padding != 1:
            raise ValueError("cremage_b200: Upsample supports dims=2, padding=1")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)

    def _pack(self, device):
        if not self.use_conv:
            return {}
        w = self.conv.weight.detach().to(device=device, dtype=torch.float32)
        return {"w4": ops.pack_weight_up2x(w), "b": f32(self.conv.bias, device)}

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        p = self.packed(x.device)
        if not self.use_conv:
            return ops.upsample2x(x)
        # nearest 2x + conv3x3 folded into four 2x2 convs over the low-resolution tensor (ops.conv3x3_up2x)
        return ops.conv3x3_up2x(x, p["w4"], self.out_channels, p["b"])

    def forward(self, x):
        require_cuda(x, "Upsample.forward")
        assert x.shape[1] == self.channels
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


class Downsample(PackedModule):
    """openaimodel.py:138-164: conv3x3 stride 2 pad 1 (use_conv) on parity planes."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or padding != 1 or not use_conv:
            raise ValueError("cremage_b200: Downsample supports the learned stride-2 conv (dims=2, padding=1) only")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=padding)

    def _pack(self, device):
        return {"w": packw(self.op.weight, device), "b": f32(self.op.bias, device)}

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        p = self.packed(x.device)
        n, h, w, c = x.shape
        if h % 2 or w % 2:
            raise ValueError("cremage_b200: Downsample needs even spatial extents")
        xs = ops.parity_split(x)
        out = ops.igemm(xs.view(4 * n, h // 2, w // 2, c), p["w"], self.out_channels, out_grid=(n, h // 2, w // 2),
                        taps=ops.taps_3x3_stride2(n), bias=p["b"], gn_stats=True)
        return ops.nhwc(out, n, h // 2, w // 2, self.out_channels)

    def forward(self, x):
        require_cuda(x, "Downsample.forward")
        assert x.shape[1] == self.channels
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


class ResBlock(TimestepBlock, PackedModule):
    """openaimodel.py:167-279: conv3x3(SiLU(GN32(x))) + emb -> conv3x3(SiLU(GN32(.))) + skip(x)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        PackedModule.__init__(self)
        if use_scale_shift_norm or up or down or use_conv:
            raise NotImplementedError("cremage_b200: ResBlock variants use_scale_shift_norm / up / down / use_conv are "
                                      "not used by the SD1.5 / SDXL configs and are not implemented")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.updown = False
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)
        self._emb_offset = None  # column of this block inside the UNet-wide fused timestep projection

    def _pack(self, device):
        return self._pack_split(device, None)

    def _pack_split(self, device, split):
        # conv1 reads the normalised concat, which GroupNorm writes as ONE tensor -> single-source K layout;
        # only the 1x1 skip conv reads the two raw sources
        p = {"g1": f32(self.in_layers[0].weight, device), "b1": f32(self.in_layers[0].bias, device),
             "w1": packw(self.in_layers[2].weight, device), "c1": f32(self.in_layers[2].bias, device),
             "g2": f32(self.out_layers[0].weight, device), "b2": f32(self.out_layers[0].bias, device),
             "w2": packw(self.out_layers[3].weight, device), "c2": f32(self.out_layers[3].bias, device),
             "we": packw(self.emb_layers[1].weight, device), "be": f32(self.emb_layers[1].bias, device),
             "split": split}
        if not isinstance(self.skip_connection, nn.Identity):
            p["ws"] = packw(self.skip_connection.weight, device, split)
            p["cs"] = f32(self.skip_connection.bias, device)
        return p

    def packed_for(self, device, split):
        """Weights packed for a (c0, c1) two-source input (the K layout pads each source to whole 64-channel chunks)."""
        p = self.packed(device)
        if p["split"] != split:
            with torch.no_grad():
                self._cb_packed = self._pack_split(device, split)
            p = self._cb_packed
        return p

    def _run(self, x: torch.Tensor, skip: Optional[torch.Tensor], emb_bias: torch.Tensor) -> torch.Tensor:
        """x (+ skip, concatenated on channels): NHWC bf16; emb_bias: fp32 [n, out_channels] (may be a strided view)."""
        split = None if skip is None else (x.shape[-1], skip.shape[-1])
        p = self.packed_for(x.device, split)
        n, hh, ww, _ = x.shape
        co = self.out_channels
        g = ops.groupnorm(x, p["g1"], p["b1"], self.in_layers[0].eps, silu=True, x1=skip)
        h = ops.nhwc(ops.igemm(g, p["w1"], co, taps=ops.TAPS_3X3, bias=p["c1"], rowbias=emb_bias, gn_stats=True), n, hh, ww, co)
        g2 = ops.groupnorm(h, p["g2"], p["b2"], self.out_layers[0].eps, silu=True)
        if "ws" in p:
            xs = ops.igemm(x, p["ws"], co, a1=skip, bias=p["cs"])
        else:
            if skip is not None:
                raise ValueError("identity skip connection with a concatenated input")
            xs = x.view(-1, co)
        out = ops.igemm(g2, p["w2"], co, taps=ops.TAPS_3X3, bias=p["c2"], residual=xs, gn_stats=True)
        return ops.nhwc(out, n, hh, ww, co)

    def _emb_out(self, emb: torch.Tensor) -> torch.Tensor:
        p = self.packed(emb.device)
        semb = ops.silu_add(emb.to(ops.ACT).contiguous())
        return ops.igemm(semb, p["we"], self.out_channels, bias=p["be"], out_f32=True)

    def forward(self, x, emb):
        require_cuda(x, "ResBlock.forward")
        y = self._run(ops.nchw_to_nhwc(x), None, self._emb_out(emb))
        return ops.nhwc_to_nchw_f32(y).to(x.dtype)


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """openaimodel.py:74-92: routes (x, emb) / (x, context) / (x) by layer type."""

    def _run(self, h, skip, emb_all, ctx2d, nk):
        for layer in self:
            if isinstance(layer, ResBlock):
                if callable(emb_all):
                    eb = emb_all(layer)
                else:
                    off = layer._emb_offset
                    eb = emb_all[:, off:off + layer.out_channels]
                h = layer._run(h, skip, eb)
                skip = None
            elif isinstance(layer, SpatialTransformer):
                h = layer._run(h, ctx2d, nk)
            else:
                h = layer._run(h)
        if skip is not None:
            raise ValueError("skip connection was not consumed (block does not start with a ResBlock)")
        return h

    def forward(self, x, emb, context=None):
        require_cuda(x, "TimestepEmbedSequential.forward")
        if len(self) == 1 and isinstance(self[0], nn.Conv2d):
            raise NotImplementedError("call UNetModel.forward; the input conv has no standalone CUDA wrapper")
        ctx2d, nk = None, 0
        if context is not None:
            nk = context.shape[1]
            ctx2d = context.reshape(-1, context.shape[-1]).to(ops.ACT).contiguous()
        y = self._run(ops.nchw_to_nhwc(x), None, lambda layer: layer._emb_out(emb), ctx2d, nk)
        return ops.nhwc_to_nchw_f32(y).to(x.dtype)


class UNetModel(PackedModule):
    """openaimodel.py:417-816 -- the full UNet with attention and timestep embedding.

    Also the implementation behind the sgm (SDXL) mirror `cremage_b200.sgm...openaimodel.UNetModel`, which adds
    per-level `transformer_depth`, honoured `use_linear_in_transformer` and the `num_classes="sequential"` vector
    conditioning `y` (sgm/modules/diffusionmodules/openaimodel.py:506-534,617-625,828-874)."""

    _ST_CLS = SpatialTransformer

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 disable_self_attentions=None, num_attention_blocks=None, disable_middle_self_attn=False,
                 use_linear_in_transformer=False, lora_ranks: List[int] = None, lora_weights: List[float] = None,
                 ipa_scale=1.0, ipa_num_tokens=0, adm_in_channels=None):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "context_dim is required with use_spatial_transformer"
        if context_dim is not None:
            assert use_spatial_transformer, "context_dim requires use_spatial_transformer"
            if not isinstance(context_dim, int):
                context_dim = list(context_dim)
        if not use_spatial_transformer:
            raise NotImplementedError("cremage_b200: the legacy AttentionBlock UNet (use_spatial_transformer=False) is "
                                      "not on the SD path and is not implemented")
        if dims != 2 or use_scale_shift_norm or resblock_updown or n_embed is not None or not conv_resample:
            raise NotImplementedError("cremage_b200: dims!=2 / use_scale_shift_norm / resblock_updown / n_embed / "
                                      "conv_resample=False are not implemented")
        if num_classes is not None and num_classes != "sequential":
            raise NotImplementedError("cremage_b200: only num_classes=None or 'sequential' (SDXL vector conditioning) "
                                      "is implemented")
        if num_classes == "sequential":
            assert adm_in_channels is not None, "num_classes='sequential' needs adm_in_channels"
        if isinstance(transformer_depth, int):
            transformer_depth = len(channel_mult) * [transformer_depth]
        transformer_depth = list(transformer_depth)
        assert len(transformer_depth) == len(channel_mult), "transformer_depth must be an int or one entry per level"
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        if isinstance(num_res_blocks, int):
            self.num_res_blocks = len(channel_mult) * [num_res_blocks]
        else:
            if len(num_res_blocks) != len(channel_mult):
                raise ValueError("provide num_res_blocks either as an int (globally constant) or "
                                 "as a list/tuple (per-level) with the same length as channel_mult")
            self.num_res_blocks = num_res_blocks
        if disable_self_attentions is not None:
            assert len(disable_self_attentions) == len(channel_mult)
        if num_attention_blocks is not None:
            assert len(num_attention_blocks) == len(self.num_res_blocks)
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False
        self.context_dim = context_dim

        def make_st(ch, heads, dim_head, depth, disabled_sa=False):
            return self._ST_CLS(ch, heads, dim_head, depth=depth, context_dim=context_dim,
                                disable_self_attn=disabled_sa, use_linear=use_linear_in_transformer,
                                use_checkpoint=use_checkpoint, lora_ranks=lora_ranks, lora_weights=lora_weights,
                                ipa_scale=ipa_scale, ipa_num_tokens=ipa_num_tokens)

        def head_cfg(ch, heads):
            if num_head_channels == -1:
                dim_head = ch // heads
            else:
                heads = ch // num_head_channels
                dim_head = num_head_channels
            if legacy:
                dim_head = ch // heads  # use_spatial_transformer is always True here
            return heads, dim_head

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), nn.SiLU(),
                                        linear(time_embed_dim, time_embed_dim))
        if num_classes == "sequential":  # sgm openaimodel.py:617-625
            self.adm_in_channels = adm_in_channels
            self.label_emb = nn.Sequential(nn.Sequential(linear(adm_in_channels, time_embed_dim), nn.SiLU(),
                                                         linear(time_embed_dim, time_embed_dim)))
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        self._feature_size = model_channels
        input_block_chans = [model_channels]
        ch = model_channels
        ds = 1
        for level, mult in enumerate(channel_mult):
            for nr in range(self.num_res_blocks[level]):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                   use_checkpoint=use_checkpoint)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    heads, dim_head = head_cfg(ch, num_heads)
                    disabled_sa = disable_self_attentions[level] if disable_self_attentions is not None else False
                    if num_attention_blocks is None or nr < num_attention_blocks[level]:
                        layers.append(make_st(ch, heads, dim_head, transformer_depth[level], disabled_sa))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                out_ch = ch
                self.input_blocks.append(
                    TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=out_ch)))
                ch = out_ch
                input_block_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        heads, dim_head = head_cfg(ch, num_heads)
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint),
            make_st(ch, heads, dim_head, transformer_depth[-1], disable_middle_self_attn),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint))
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(self.num_res_blocks[level] + 1):
                ich = input_block_chans.pop()
                layers = [ResBlock(ch + ich, time_embed_dim, dropout, out_channels=model_channels * mult, dims=dims,
                                   use_checkpoint=use_checkpoint)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    heads, dim_head = head_cfg(ch, num_heads_upsample)
                    disabled_sa = disable_self_attentions[level] if disable_self_attentions is not None else False
                    if num_attention_blocks is None or i < num_attention_blocks[level]:
                        layers.append(make_st(ch, heads, dim_head, transformer_depth[level], disabled_sa))
                if level and i == self.num_res_blocks[level]:
                    out_ch = ch
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=out_ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))

        # fused timestep projection: every ResBlock gets a column range of one [sum(out_channels), 4*mc] GEMM
        off = 0
        self._res_blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        for rb in self._res_blocks:
            rb._emb_offset = off
            off += rb.out_channels
        self._emb_total = off
        # CUDA-graph replay of the whole forward (one capture per input signature); CREMAGE_B200_GRAPH=0 disables
        self.use_cuda_graph = os.environ.get("CREMAGE_B200_GRAPH", "1") != "0"
        self._graphed = None
        # cross-attention K/V of the current context (they depend on the context and the weights only): computed once
        # per context tensor, reused by every sampler step; the replayed graph then holds no K/V projection launches
        self.use_kv_cache = os.environ.get("CREMAGE_B200_KV_CACHE", "1") != "0"
        self._kv = None
        self._graphed_kv = None

    # -- packing -------------------------------------------------------------------------------------------------
    def _own_params(self):
        ps = list(self.time_embed.parameters()) + list(self.input_blocks[0].parameters()) + list(self.out.parameters())
        if self.num_classes is not None:
            ps += list(self.label_emb.parameters())
        for rb in self._res_blocks:
            ps += list(rb.emb_layers.parameters())
        return ps

    def _pack(self, device):
        cin, cin_pad = self.in_channels, _ceil8(self.in_channels)
        w_in = self.input_blocks[0][0].weight.detach().to(device=device, dtype=torch.float32)  # [mc, cin, 3, 3]
        embw = torch.cat([rb.emb_layers[1].weight.detach().to(device=device, dtype=torch.float32)
                          for rb in self._res_blocks], 0)
        embb = torch.cat([rb.emb_layers[1].bias.detach().to(device=device, dtype=torch.float32)
                          for rb in self._res_blocks], 0)
        extra = {}
        if self.num_classes is not None:
            le = self.label_emb[0]
            extra = {"le0w": packw(le[0].weight, device), "le0b": f32(le[0].bias, device),
                     "le2w": packw(le[2].weight, device), "le2b": f32(le[2].bias, device)}
        return {
            **extra,
            "freqs": timestep_freqs(self.model_channels).to(device),
            "te0w": packw(self.time_embed[0].weight, device), "te0b": f32(self.time_embed[0].bias, device),
            "te2w": packw(self.time_embed[2].weight, device), "te2b": f32(self.time_embed[2].bias, device),
            "embw": ops.pack_weight(embw), "embb": embb.contiguous(),
            "inw": ops.pack_weight(w_in), "inb": f32(self.input_blocks[0][0].bias, device),
            "cin": cin, "cin_pad": cin_pad,
            "og": f32(self.out[0].weight, device), "ob": f32(self.out[0].bias, device),
            "ow": packw(self.out[2].weight, device), "oc": f32(self.out[2].bias, device),
        }

    # -- forward -------------------------------------------------------------------------------------------------
    def _forward_impl(self, x: torch.Tensor, t: torch.Tensor, context: Optional[torch.Tensor], y: torch.Tensor = None,
                      control: Optional[Sequence[torch.Tensor]] = None, kv: Optional[dict] = None) -> torch.Tensor:
        """control: ControlNet residuals in the order ControlledUnetModel.forward pops them (cldm/cldm.py:59-66):
        control[0] is added to the middle block's output, control[1 + i] to the skip of output block i (NCHW)."""
        dev = x.device
        p = self.packed(dev)
        n, c, hh, ww = x.shape
        mc, ted = self.model_channels, self.model_channels * 4
        # timestep embedding -> time_embed MLP (SiLU fused) -> all per-block projections in one GEMM
        temb = ops.timestep_embedding(t, mc, p["freqs"])
        e = ops.igemm(temb, p["te0w"], ted, bias=p["te0b"], act=ops.ACT_SILU)
        if self.num_classes is None:
            semb = ops.igemm(e, p["te2w"], ted, bias=p["te2b"], act=ops.ACT_SILU)  # silu(emb): every consumer applies SiLU first
        else:  # emb = time_embed(t_emb) + label_emb(y)   (sgm openaimodel.py:851-859), then the shared SiLU
            emb_t = ops.igemm(e, p["te2w"], ted, bias=p["te2b"])
            l0 = ops.igemm(y.to(ops.ACT).contiguous(), p["le0w"], ted, bias=p["le0b"], act=ops.ACT_SILU)
            emb = ops.igemm(l0, p["le2w"], ted, bias=p["le2b"], residual=emb_t)
            semb = ops.silu_add(emb)
        emb_all = ops.igemm(semb, p["embw"], self._emb_total, bias=p["embb"], out_f32=True)
        if kv is not None:      # context already projected (see _kv_for): ctx2d carries the cached K/V of every attn2
            nk, ctx2d = kv["nk"], kv["ctx2d"]
        else:
            nk = context.shape[1]
            ctx2d = context.reshape(n * nk, context.shape[-1]).to(ops.ACT).contiguous()

        h = ops.nchw_to_nhwc(x, c_pad=p["cin_pad"])
        # conv_in through the tensor-core path: the 64-channel TMA box reads channels >= cin_pad as out-of-bounds zeros
        h = ops.nhwc(ops.igemm(h, p["inw"], mc, taps=ops.TAPS_3X3, bias=p["inb"], gn_stats=True), *h.shape[:3], mc)
        hs = [h]
        for module in list(self.input_blocks)[1:]:
            h = module._run(h, None, emb_all, ctx2d, nk)
            hs.append(h)
        h = self.middle_block._run(h, None, emb_all, ctx2d, nk)
        control = list(control) if control else []
        if control:
            h = ops.add_nchw(h, control[0])
        for i, module in enumerate(self.output_blocks):
            skip = hs.pop()
            if 1 + i < len(control):
                skip = ops.add_nchw(skip, control[1 + i])
            h = module._run(h, skip, emb_all, ctx2d, nk)
        g = ops.groupnorm(h, p["og"], p["ob"], self.out[0].eps, silu=True)
        o = ops.igemm(g, p["ow"], self.out_channels, taps=ops.TAPS_3X3, bias=p["oc"], out_f32=True)
        return ops.nhwc_to_nchw_f32(o.view(n, hh, ww, self.out_channels))

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """Apply the model to an input batch (openaimodel.py:780-816).
        x: [N, C, H, W]; timesteps: [N] (float, possibly fractional, or int64); context: [N, S, context_dim]."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        require_cuda(x, "UNetModel.forward")
        if timesteps is None or context is None:
            raise ValueError("UNetModel.forward needs timesteps and context")
        if x.shape[0] != timesteps.shape[0] or x.shape[0] != context.shape[0]:
            raise ValueError("batch sizes of x, timesteps and context differ")
        if y is not None:
            assert y.shape[0] == x.shape[0] and y.shape[1] == self.adm_in_channels, "bad vector conditioning shape"
        if x.shape[2] % (2 ** (len(self.channel_mult) - 1)) or x.shape[3] % (2 ** (len(self.channel_mult) - 1)):
            raise ValueError("latent height/width must be divisible by the total downsampling factor")
        t = timesteps.to(device=x.device, dtype=torch.float32).contiguous()
        ctx = context.to(device=x.device)
        xin = x.contiguous()
        extra = () if y is None else (y.to(device=x.device).contiguous(),)
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            self.packed(x.device)  # refresh packs (and drop stale graphs / cached K/V) if parameters changed
            if self.use_kv_cache and self._kv_modules() is not None:
                self._kv_for(ctx)
                if self._graphed_kv is None:
                    self._graphed_kv = GraphedCall(self._forward_kv, keepalive=self._graph_keepalive)
                out = self._graphed_kv(xin, t, *extra)
            else:
                if self._graphed is None:
                    self._graphed = GraphedCall(self._forward_impl, keepalive=self._graph_keepalive)
                out = self._graphed(xin, t, ctx, *extra)
        else:
            out = self._forward_impl(xin, t, ctx, *extra)
        return out.to(x.dtype)

    # -- cross-attention K/V cache ----------------------------------------------------------------------------------
    def _forward_kv(self, x, t, *extra):
        return self._forward_impl(x, t, None, *extra, kv=self._kv)

    def _kv_modules(self):
        """The cross-attention modules whose K/V depend on the context only, in forward order (None: a configuration
        where attn1 also reads the context -- disable_self_attn -- keeps the uncached path)."""
        mods = self.__dict__.get("_cb_kv_mods")
        if mods is None:
            from ..attention import BasicTransformerBlock
            blocks = [m for m in self.modules() if isinstance(m, BasicTransformerBlock)]
            mods = False if any(b.disable_self_attn for b in blocks) or not blocks else [b.attn2 for b in blocks]
            self.__dict__["_cb_kv_mods"] = mods
        return mods or None

    def _kv_for(self, ctx: torch.Tensor) -> dict:
        """K/V of `ctx` for every attn2.  Valid while the caller passes the SAME tensor object, unmodified (identity +
        torch's in-place version counter; a strong reference keeps the address from being reused): the samplers hand
        the same CFG-doubled context to every step.  A new / modified context is re-projected eagerly into the same
        buffers (their addresses are baked into the captured graph); a new shape gets new buffers and a new graph."""
        kv = self._kv
        shape = tuple(ctx.shape)
        if (kv is not None and kv["ctx"] is ctx and kv["ver"] == ctx._version and kv["shape"] == shape
                and kv["dtype"] == ctx.dtype):
            return kv
        if (kv is not None and kv["shape"] == shape and kv["dtype"] == ctx.dtype and kv["ctx"].device == ctx.device
                and kv["ver"] == kv["ctx"]._version            # the cached tensor itself is still what was projected
                and not torch.cuda.is_current_stream_capturing() and torch.equal(kv["ctx"], ctx)):
            # a NEW tensor with the SAME content: the reference's own wrapper builds a fresh torch.cat([uc, c]) on every
            # sampler step (ldm_wrapper_for_k_diffusion.py:67-92).  One compare kernel + a host sync (~30 us) keeps such a
            # caller on the cached path instead of re-projecting the context through 16 (SDXL: 70) GEMMs per step.
            kv["ctx"], kv["ver"] = ctx, ctx._version
            return kv
        n, nk, cdim = shape
        mods = self._kv_modules()
        fresh = ctx.reshape(n * nk, cdim).to(ops.ACT).contiguous()
        if kv is not None and kv["shape"] == shape and kv["ctx2d"].device == fresh.device:
            kv["ctx2d"].copy_(fresh)
            kv["ctx2d"].__dict__.pop("_cb_ipa_split", None)
            for m, bufs in zip(mods, kv["bufs"]):
                m.compute_kv(kv["ctx2d"], n, nk, out=bufs)
        else:
            ctx2d = fresh.clone() if fresh.data_ptr() == ctx.data_ptr() else fresh   # never alias the caller's tensor
            bufs = [m.compute_kv(ctx2d, n, nk) for m in mods]
            ctx2d._cb_kv = {id(m): b for m, b in zip(mods, bufs)}
            kv = {"ctx2d": ctx2d, "bufs": bufs, "nk": nk, "shape": shape}
            self._kv = kv
            if self._graphed_kv is not None:
                self._graphed_kv.reset()
        kv["ctx"], kv["ver"], kv["dtype"] = ctx, ctx._version, ctx.dtype
        return kv

    def packed(self, device):
        before = self._cb_packed
        p = super().packed(device)
        # captured graphs hold the packed weight buffers of EVERY sub-module: any in-place parameter update anywhere in
        # the tree (optimiser step, LoRA alpha edit, weight patching) must drop them, not only this module's own packs
        from .... import engine
        epoch = engine.PACK_EPOCH
        params = self.__dict__.get("_cb_all_params")
        stale = False
        if params is None or epoch != self.__dict__.get("_cb_epoch"):
            # some PackedModule somewhere was moved / cast / reloaded since the last call (possibly a sub-tree of this
            # model: convert_to_fp16 -> blocks.half() never passes through this module's own _apply): re-list the
            # parameters and compare where they live
            params = list(self.parameters())
            self.__dict__["_cb_all_params"] = params
            self.__dict__["_cb_epoch"] = epoch
            where = tuple((q.data_ptr(), q.dtype) for q in params)
            stale = where != self.__dict__.get("_cb_all_where")
            self.__dict__["_cb_all_where"] = where
            self.__dict__.pop("_cb_kv_mods", None)
        versions = sum(q._version for q in params)
        stale = stale or versions != self.__dict__.get("_cb_all_versions")
        self.__dict__["_cb_all_versions"] = versions
        if p is not before or stale:
            self._reset_graphs()  # parameters changed: captured graphs hold stale weight buffers
        return p

    def _graph_keepalive(self):
        """Everything the captured kernels read besides the static inputs: the packed weights of every sub-module and
        the cached K/V buffers."""
        return [m._cb_packed for m in self.modules() if isinstance(m, PackedModule)] + [self.__dict__.get("_kv")]

    def _reset_graphs(self):
        if self._graphed is not None:
            self._graphed.reset()
        if self.__dict__.get("_graphed_kv") is not None:
            self._graphed_kv.reset()
        self.__dict__["_kv"] = None   # the cached K/V were projected with the old weights

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_cb_all_params", None)   # .to() / .half() may replace Parameter objects
        self.__dict__.pop("_cb_kv_mods", None)
        return super()._apply(fn, *args, **kwargs)

    def invalidate_packed(self):
        super().invalidate_packed()
        self._reset_graphs()

    def convert_to_fp16(self):
        """openaimodel.py:758-764. Parameter storage dtype only; the kernels always compute 16-bit x 16-bit -> fp32 (ops.ACT)."""
        for blocks in (self.input_blocks, self.middle_block, self.output_blocks):
            blocks.half()

    def convert_to_fp32(self):
        for blocks in (self.input_blocks, self.middle_block, self.output_blocks):
            blocks.float()
