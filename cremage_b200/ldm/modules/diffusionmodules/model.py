"""Drop-in mirror of the reference's VAE decoder and encoder (`ldm.modules.diffusionmodules.model`, HowToSD/cremage
modules/ldm/modules/diffusionmodules/model.py): Decoder (:469), Encoder (:375), ResnetBlock (:89), AttnBlock (:157),
Upsample (:49), Downsample (:66), Normalize (:45, GroupNorm eps 1e-6) and nonlinearity (:40, x*sigmoid(x)).  Same constructor arguments, parameter names
and forward signatures; the arithmetic runs on the sm_100a kernels (NHWC bf16, fp32 accumulate / statistics).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .... import ops
from ....engine import PackedModule, f32, packw, require_cuda


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(PackedModule):
    """model.py:49-64: nearest 2x then conv3x3."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        self.in_channels = in_channels
        if self.with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def _pack(self, device):
        if not self.with_conv:
            return {}
        w = self.conv.weight.detach().to(device=device, dtype=torch.float32)
        return {"w4": ops.pack_weight_up2x(w), "b": f32(self.conv.bias, device)}

    def _run(self, x):
        p = self.packed(x.device)
        if not self.with_conv:
            return ops.upsample2x(x)
        # nearest 2x + conv3x3 folded into four 2x2 convs over the low-resolution tensor (ops.conv3x3_up2x)
        return ops.conv3x3_up2x(x, p["w4"], x.shape[-1], p["b"])

    def forward(self, x):
        require_cuda(x, "Upsample.forward")
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


class Downsample(PackedModule):
    """model.py:66-86: F.pad(x, (0, 1, 0, 1)) then conv3x3 stride 2 padding 0 (asymmetric: the extra zero row / column
    sits at the bottom / right).  Runs as the stride-2 implicit GEMM over the parity-split input with the asymmetric
    tap table; the padding is TMA's out-of-bounds zero fill."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        self.in_channels = in_channels
        if not with_conv:
            raise NotImplementedError("cremage_b200: Downsample without conv (avg_pool2d) is not used by the SD VAE")
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def _pack(self, device):
        return {"w": packw(self.conv.weight, device), "b": f32(self.conv.bias, device)}

    def _run(self, x):
        p = self.packed(x.device)
        n, h, w, c = x.shape
        if h % 2 or w % 2:
            raise ValueError("cremage_b200: Downsample needs even spatial extents")
        xs = ops.parity_split(x)
        out = ops.igemm(xs.view(4 * n, h // 2, w // 2, c), p["w"], c, out_grid=(n, h // 2, w // 2),
                        taps=ops.taps_3x3_stride2_asym(n), bias=p["b"], gn_stats=True)
        return ops.nhwc(out, n, h // 2, w // 2, c)

    def forward(self, x):
        require_cuda(x, "Downsample.forward")
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


class ResnetBlock(PackedModule):
    """model.py:89-148 with temb_channels = 0 (the VAE passes temb=None)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if conv_shortcut:
            raise NotImplementedError("cremage_b200: ResnetBlock conv_shortcut is not used by the SD VAE")
        if temb_channels > 0:
            raise NotImplementedError("cremage_b200: VAE ResnetBlock is built with temb_channels=0")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def _pack(self, device):
        p = {"g1": f32(self.norm1.weight, device), "b1": f32(self.norm1.bias, device),
             "w1": packw(self.conv1.weight, device), "c1": f32(self.conv1.bias, device),
             "g2": f32(self.norm2.weight, device), "b2": f32(self.norm2.bias, device),
             "w2": packw(self.conv2.weight, device), "c2": f32(self.conv2.bias, device)}
        if self.in_channels != self.out_channels:
            p["ws"] = packw(self.nin_shortcut.weight, device)
            p["cs"] = f32(self.nin_shortcut.bias, device)
        return p

    def _run(self, x):
        p = self.packed(x.device)
        n, hh, ww, _ = x.shape
        co = self.out_channels
        g = ops.groupnorm(x, p["g1"], p["b1"], self.norm1.eps, silu=True)
        h = ops.nhwc(ops.igemm(g, p["w1"], co, taps=ops.TAPS_3X3, bias=p["c1"], gn_stats=True), n, hh, ww, co)
        g2 = ops.groupnorm(h, p["g2"], p["b2"], self.norm2.eps, silu=True)
        xs = ops.igemm(x, p["ws"], co, bias=p["cs"]) if "ws" in p else x.view(-1, co)
        return ops.nhwc(ops.igemm(g2, p["w2"], co, taps=ops.TAPS_3X3, bias=p["c2"], residual=xs, gn_stats=True), n, hh, ww, co)

    def forward(self, x, temb=None):
        require_cuda(x, "ResnetBlock.forward")
        if temb is not None:
            raise NotImplementedError("cremage_b200: VAE ResnetBlock takes temb=None")
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


class AttnBlock(PackedModule):
    """model.py:157-209: single-head attention over all pixels, head dim = channels (512 in SD): too wide for the
    fused kernel's TMEM budget (S 128 + O 512 columns > 512), so it runs as tcgen05 GEMMs around a row-softmax kernel,
    one image and one chunk of query rows at a time (the chunk's scores and probabilities stay in L2):
      q, k = 1x1(h);  S = q k^T (fp32);  P = softmax(S / sqrt(c)) (16-bit);  v^T = Wv h^T;  O = P v + bv;  x + 1x1(O)
    (the value bias is added after the PV product -- exact, softmax rows sum to one)."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def _pack(self, device):
        c = self.in_channels
        if c % 64:
            raise ValueError("cremage_b200: AttnBlock channels must be a multiple of 64")
        return {"g": f32(self.norm.weight, device), "b": f32(self.norm.bias, device),
                "wq": packw(self.q.weight, device), "bq": f32(self.q.bias, device),
                "wk": packw(self.k.weight, device), "bk": f32(self.k.bias, device),
                "wv_rows": self.v.weight.detach().to(device=device, dtype=torch.float32).reshape(c, c).to(ops.ACT).contiguous(),
                "bv": f32(self.v.bias, device),
                "wo": packw(self.proj_out.weight, device), "bo": f32(self.proj_out.bias, device)}

    def _run(self, x):
        p = self.packed(x.device)
        n, hh, ww, c = x.shape
        npx = hh * ww
        if npx % 8:
            raise ValueError("cremage_b200: AttnBlock needs h*w to be a multiple of 8")
        hn = ops.groupnorm(x, p["g"], p["b"], self.norm.eps, silu=False).view(n, npx, c)
        q = ops.igemm(hn.view(n * npx, c), p["wq"], c, bias=p["bq"]).view(n, npx, c)
        k = ops.igemm(hn.view(n * npx, c), p["wk"], c, bias=p["bk"]).view(n, npx, c)
        o = torch.empty((n, npx, c), dtype=ops.ACT, device=x.device)
        # The N x N score matrix never exists: query rows go through S -> softmax -> P V in chunks whose fp32 scores
        # (<= 32 MiB) and 16-bit probabilities stay in the 126 MB L2 between the three launches.  At 128 x 128 latents
        # (1024^2 decode: N = 16 384) the full matrix would be 1 GiB of fp32 + 0.5 GiB of P per image.
        rows_c = npx
        while rows_c > 128 and rows_c * npx * 4 > (32 << 20) and rows_c % 2 == 0:
            rows_c //= 2
        s = torch.empty((rows_c, npx), dtype=torch.float32, device=x.device)
        pm = torch.empty((rows_c, npx), dtype=ops.ACT, device=x.device)
        vt = torch.empty((c, npx), dtype=ops.ACT, device=x.device)
        scale = float(int(c) ** (-0.5))
        for i in range(n):
            ops.igemm(p["wv_rows"], hn[i], npx, out=vt)              # v^T [c, npx] = Wv h^T
            for r0 in range(0, npx, rows_c):
                r1 = min(r0 + rows_c, npx)
                sc, pc = s[:r1 - r0], pm[:r1 - r0]
                ops.igemm(q[i][r0:r1], k[i], npx, out=sc)            # S = q k^T, the "weights" operand is k itself
                ops.softmax_rows(sc, scale, out=pc)
                ops.igemm(pc, vt, c, bias=p["bv"], out=o[i][r0:r1])  # O = P v + bv
        out = ops.igemm(o.view(n, 1, npx, c), p["wo"], c, bias=p["bo"], residual=x.view(n * npx, c), gn_stats=True)
        return ops.nhwc(out, n, hh, ww, c)

    def forward(self, x):
        require_cuda(x, "AttnBlock.forward")
        return ops.nhwc_to_nchw_f32(self._run(ops.nchw_to_nhwc(x))).to(x.dtype)


def make_attn(in_channels, attn_type="vanilla"):
    """model.py:212-220."""
    if attn_type == "vanilla":
        return AttnBlock(in_channels)
    if attn_type == "none":
        return nn.Identity(in_channels)
    raise NotImplementedError(f"cremage_b200: attn_type {attn_type} is not implemented")


class Decoder(PackedModule):
    """model.py:469-575."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or give_pre_end or tanh_out:
            raise NotImplementedError("cremage_b200: use_linear_attn / give_pre_end / tanh_out are not implemented")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.out_ch = out_ch
        self.z_channels = z_channels
        self.give_pre_end = give_pre_end
        self.tanh_out = tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for i_block in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def _own_params(self):
        return list(self.conv_in.parameters()) + list(self.norm_out.parameters()) + list(self.conv_out.parameters())

    def _pack(self, device):
        w_in = self.conv_in.weight.detach().to(device=device, dtype=torch.float32)
        return {"inw": ops.pack_weight(w_in), "inb": f32(self.conv_in.bias, device),
                "og": f32(self.norm_out.weight, device), "ob": f32(self.norm_out.bias, device),
                "ow": packw(self.conv_out.weight, device), "oc": f32(self.conv_out.bias, device)}

    def _run(self, z_nhwc: torch.Tensor) -> torch.Tensor:
        """z_nhwc: bf16 [n, h, w, >= z_channels] -> fp32 NHWC [n, 8h.., 8w.., 4] (first out_ch channels valid)."""
        p = self.packed(z_nhwc.device)
        block_in = self.conv_in.out_channels
        # (the 64-channel TMA box reads channels beyond the tensor's ceil8(z_channels) as out-of-bounds zeros)
        h = ops.nhwc(ops.igemm(z_nhwc, p["inw"], block_in, taps=ops.TAPS_3X3, bias=p["inb"], gn_stats=True),
                     *z_nhwc.shape[:3], block_in)
        h = self.mid.block_1._run(h)
        if isinstance(self.mid.attn_1, AttnBlock):
            h = self.mid.attn_1._run(h)
        h = self.mid.block_2._run(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block]._run(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block]._run(h)
            if i_level != 0:
                h = self.up[i_level].upsample._run(h)
        g = ops.groupnorm(h, p["og"], p["ob"], self.norm_out.eps, silu=True)
        n, hh, ww, _ = g.shape
        ld = 4 if self.out_ch <= 4 else (self.out_ch + 3) // 4 * 4
        o = ops.igemm(g, p["ow"], self.out_ch, taps=ops.TAPS_3X3, bias=p["oc"], out_f32=True, out_ld=ld)
        return o.view(n, hh, ww, ld)

    def forward(self, z):
        require_cuda(z, "Decoder.forward")
        self.last_z_shape = z.shape
        o = self._run(ops.nchw_to_nhwc(z.float(), c_pad=8))
        return ops.nhwc_to_nchw_f32(o, self.out_ch).to(z.dtype)


class Encoder(PackedModule):
    """model.py:375-466: conv_in, per level `num_res_blocks` ResnetBlocks (+ Downsample), mid (Res, Attn, Res),
    GroupNorm + swish + conv_out to 2 * z_channels moments.  `_run` returns the moments as fp32 NHWC."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if use_linear_attn:
            raise NotImplementedError("cremage_b200: use_linear_attn is not implemented")
        if in_channels > 8:
            raise NotImplementedError("cremage_b200: Encoder input channels above 8 are not implemented")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for i_block in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.out_channels = 2 * z_channels if double_z else z_channels
        self.conv_out = nn.Conv2d(block_in, self.out_channels, kernel_size=3, stride=1, padding=1)

    def _own_params(self):
        return list(self.conv_in.parameters()) + list(self.norm_out.parameters()) + list(self.conv_out.parameters())

    def _pack(self, device):
        w_in = self.conv_in.weight.detach().to(device=device, dtype=torch.float32)
        return {"inw": ops.pack_weight(w_in), "inb": f32(self.conv_in.bias, device),
                "og": f32(self.norm_out.weight, device), "ob": f32(self.norm_out.bias, device),
                "ow": packw(self.conv_out.weight, device), "oc": f32(self.conv_out.bias, device)}

    def _trunk(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        """image NHWC 16-bit [n, H, W, >= in_channels] -> normalised + swish features before conv_out."""
        p = self.packed(x_nhwc.device)
        h = ops.nhwc(ops.igemm(x_nhwc, p["inw"], self.ch, taps=ops.TAPS_3X3, bias=p["inb"], gn_stats=True),
                     *x_nhwc.shape[:3], self.ch)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block]._run(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block]._run(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample._run(h)
        h = self.mid.block_1._run(h)
        if isinstance(self.mid.attn_1, AttnBlock):
            h = self.mid.attn_1._run(h)
        h = self.mid.block_2._run(h)
        return ops.groupnorm(h, p["og"], p["ob"], self.norm_out.eps, silu=True)

    def _run(self, x_nhwc: torch.Tensor, out_w=None, out_b=None, out_c=None) -> torch.Tensor:
        """-> fp32 NHWC [n, h, w, ld] moments; (out_w, out_b, out_c) override conv_out (AutoencoderKL folds quant_conv
        into it: two linear maps with nothing in between)."""
        p = self.packed(x_nhwc.device)
        g = self._trunk(x_nhwc)
        n, hh, ww, _ = g.shape
        co = self.out_channels if out_c is None else out_c
        ld = (co + 3) // 4 * 4
        o = ops.igemm(g, p["ow"] if out_w is None else out_w, co, taps=ops.TAPS_3X3,
                      bias=p["oc"] if out_b is None else out_b, out_f32=True, out_ld=ld)
        return o.view(n, hh, ww, ld)

    def forward(self, x):
        require_cuda(x, "Encoder.forward")
        o = self._run(ops.nchw_to_nhwc(x.float(), c_pad=8))
        return ops.nhwc_to_nchw_f32(o, self.out_channels).to(x.dtype)
