"""FLUX.1 autoencoder -- encoder side -- on the sm_100a kernels.

Drop-in for /root/reference/Continuous/src/flux/modules/autoencoder.py: ``AutoEncoderParams`` (:8-18),
``AutoEncoder`` with ``.encoder`` / ``.reg`` / ``.encode(x)`` (:277-305) and the encoder's state_dict keys
(``encoder.conv_in.*``, ``encoder.down.{l}.block.{j}.{norm1,conv1,norm2,conv2,nin_shortcut}.*``,
``encoder.down.{l}.downsample.conv.*``, ``encoder.mid.{block_1,attn_1,block_2}.*``, ``encoder.norm_out.*``,
``encoder.conv_out.*``), so FLUX's ``ae.safetensors`` loads with ``strict=False`` exactly as in util.py:227-246.

Only the encoder is on GenHancer's training path (every train script calls ``vae.encode`` under no_grad,
train_SigLIP_stage1.py:242-243); the decoder is used by the stale reconstruction demo alone
and its tensors are ignored when a checkpoint is loaded.

Data layout: activations are NHWC bf16 (channels innermost = the K axis of the implicit GEMM); every 3x3 conv
is one tcgen05 implicit-GEMM launch whose A tiles are shifted 4-D TMA boxes (zero fill = padding, incl. the
Downsample's right/bottom pad); GroupNorm(32)+swish is a two-pass HBM-bound kernel with fp32/fp64 statistics;
the conv_out epilogue writes fp32 moments which the sampling kernel turns into the 2x2-patchified latent.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from ... import kernels as K
from ...kernels import ACT_NONE, BF16, F32


@dataclass
class AutoEncoderParams:  # autoencoder.py:8-18
    resolution: int
    in_channels: int
    ch: int
    out_ch: int
    ch_mult: list[int]
    num_res_blocks: int
    z_channels: int
    scale_factor: float
    shift_factor: float


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=ch, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):  # parameter tree of autoencoder.py:57-82
    def __init__(self, in_channels: int, out_channels: int | None = None):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _gn(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _gn(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):  # parameter tree of autoencoder.py:25-55
    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _gn(in_channels)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(in_channels, in_channels, 1) for _ in range(4))


class Downsample(nn.Module):  # autoencoder.py:85-95
    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class _ConvStack(nn.Module):
    """What Encoder and Decoder share on the B200 path: cached bf16 conv operands and the ResnetBlock / AttnBlock
    schedules over the implicit-GEMM conv, GroupNorm and GEMM kernels (NHWC bf16 activations)."""

    # ---- cached kernel operands: conv weights as bf16 [Cout, (kh,kw,ci)], fp32 biases / norm affines ----------
    def _prepared(self) -> dict:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._cache is not None and self._cache_key == key:
            return self._cache
        W: dict = {}
        for name, m in self.named_modules():
            if isinstance(m, nn.Conv2d):
                w = m.weight.detach()
                o, i, kh, kw = w.shape
                if name == "conv_in" and i == 3:  # k = (kh*3+kw)*3 + c, padded 27 -> 32 (gh_im2col3x3_c3 layout)
                    wk = torch.zeros(o, 32, dtype=BF16, device=w.device)
                    wk[:, :27] = w.permute(0, 2, 3, 1).reshape(o, 27).to(BF16)
                elif name == "conv_in":  # decoder: z channels, F.unfold column order (ci, kh, kw)
                    wk = w.reshape(o, i * kh * kw).to(BF16).contiguous()
                elif o % 8 != 0:  # conv_out of the decoder (3 channels): zero rows up to a multiple of 8
                    o8 = (o + 7) // 8 * 8
                    wk = torch.zeros(o8, kh * kw * i, dtype=BF16, device=w.device)
                    wk[:o] = w.permute(0, 2, 3, 1).reshape(o, kh * kw * i).to(BF16)
                    bb = torch.zeros(o8, dtype=torch.float32, device=w.device)
                    bb[:o] = m.bias.detach().float()
                    W[name] = (wk, bb)
                    continue
                else:
                    wk = w.permute(0, 2, 3, 1).reshape(o, kh * kw * i).to(BF16).contiguous()
                W[name] = (wk, m.bias.detach().float().contiguous())
            elif isinstance(m, nn.GroupNorm):
                W[name] = (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
        a = self.mid.attn_1  # q,k,v 1x1 convs as ONE GEMM with N = 3C
        W["mid.attn_1.qkv"] = (torch.cat([W["mid.attn_1.q"][0], W["mid.attn_1.k"][0], W["mid.attn_1.v"][0]], 0).contiguous(),
                               torch.cat([W["mid.attn_1.q"][1], W["mid.attn_1.k"][1], W["mid.attn_1.v"][1]], 0).contiguous())
        del a
        self._cache, self._cache_key = W, key
        return W

    def _res(self, W: dict, name: str, blk: ResnetBlock, h: Tensor) -> Tensor:
        t = K.groupnorm_swish_nhwc(h, *W[f"{name}.norm1"])
        t = K.conv2d_nhwc(t, W[f"{name}.conv1"][0], 3, 3, 1, 1, bias=W[f"{name}.conv1"][1])
        t = K.groupnorm_swish_nhwc(t, *W[f"{name}.norm2"])
        if blk.in_channels != blk.out_channels:
            B, H, Wd, Ci = h.shape
            sc = K.gemm(h.view(-1, Ci), W[f"{name}.nin_shortcut"][0], bias=W[f"{name}.nin_shortcut"][1])
            h = sc.view(B, H, Wd, blk.out_channels)
        return K.conv2d_nhwc(t, W[f"{name}.conv2"][0], 3, 3, 1, 1, bias=W[f"{name}.conv2"][1], residual=h)

    def _attn(self, W: dict, h: Tensor) -> Tensor:
        """AttnBlock (autoencoder.py:37-55): single head over H*W tokens, d = C, scale C^-0.5, residual."""
        B, H, Wd, C = h.shape
        L = H * Wd
        y = K.groupnorm_swish_nhwc(h, *W["mid.attn_1.norm"], swish=False)
        qkv = K.gemm(y.view(-1, C), W["mid.attn_1.qkv"][0], bias=W["mid.attn_1.qkv"][1]).view(B, L, 3 * C)
        # all B per-image products in ONE batched launch each (flat [B*L, .] operands, gh_gemm_bf16 batch mode):
        # 32 x (Q K^T, softmax, P V) was 96 launch-bound kernels
        ldp = (L + 7) // 8 * 8
        flat = qkv.view(B * L, 3 * C)
        s = K.gemm(flat[:, :C], flat[:, C:2 * C], out_dtype=F32, batch=B)            # [B*L, L] fp32 scores
        p = K.softmax_rows(s, L, C ** -0.5, ldp)                                     # bf16, pad columns zero
        o = K.gemm(p[:, :L], flat[:, 2 * C:], b_mn=True, batch=B)                    # P @ V (V is [K=L, N=C] per image)
        out = K.gemm(o.view(-1, C), W["mid.attn_1.proj_out"][0], bias=W["mid.attn_1.proj_out"][1],
                     residual=h.view(-1, C))
        return out.view(B, H, Wd, C)



class Upsample(nn.Module):  # autoencoder.py:98-106
    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Encoder(_ConvStack):
    """Encoder.forward of autoencoder.py:159-180; returns the [B, 2z, H/8, W/8] moments (fp32, NCHW view)."""

    def __init__(self, resolution: int, in_channels: int, ch: int, ch_mult: list[int], num_res_blocks: int,
                 z_channels: int):
        super().__init__()
        if in_channels != 3:
            raise NotImplementedError("the sm_100a conv_in gather is written for 3-channel images")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.z_channels = resolution, in_channels, z_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(self.num_resolutions):
            block = nn.ModuleList()
            block_in = ch * in_ch_mult[lvl]
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            down = nn.Module()
            down.block = block
            down.attn = nn.ModuleList()
            if lvl != self.num_resolutions - 1:
                down.downsample = Downsample(block_in)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels, 3, 1, 1)
        self._cache = None
        self._cache_key = None

    def moments_nhwc(self, img: Tensor, mean: float = 0.0, std: float = 1.0) -> Tensor:
        """img fp32 NCHW; (img - mean) / std is folded into the conv_in gather. -> fp32 [B, H/8, W/8, 2z]."""
        u8 = K.is_u8_image(img)                 # the decoded uint8 HWC batch: u8 / 255 happens inside the conv_in gather
        if not u8 and (img.dim() != 4 or img.shape[1] != 3):
            raise ValueError(f"AutoEncoder expects [B,3,H,W] images, got {tuple(img.shape)}")
        W = self._prepared()
        B, H, Wd = K.image_bhw(img)
        a = K.im2col3x3_c3(img.contiguous() if u8 else img.float().contiguous(), mean, std)
        h = K.gemm(a, W["conv_in"][0], bias=W["conv_in"][1]).view(B, H, Wd, self.ch)
        for lvl in range(self.num_resolutions):
            for j, blk in enumerate(self.down[lvl].block):
                h = self._res(W, f"down.{lvl}.block.{j}", blk, h)
            if lvl != self.num_resolutions - 1:
                n = f"down.{lvl}.downsample.conv"
                Hh, Ww = h.shape[1], h.shape[2]
                h = K.conv2d_nhwc(h, W[n][0], 3, 3, stride=2, pad=0, Ho=(Hh - 2) // 2 + 1, Wo=(Ww - 2) // 2 + 1,
                                  bias=W[n][1])
        h = self._res(W, "mid.block_1", self.mid.block_1, h)
        h = self._attn(W, h)
        h = self._res(W, "mid.block_2", self.mid.block_2, h)
        h = K.groupnorm_swish_nhwc(h, *W["norm_out"])
        return K.conv2d_nhwc(h, W["conv_out"][0], 3, 3, 1, 1, bias=W["conv_out"][1], act=ACT_NONE, out_dtype=F32)

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        return self.moments_nhwc(x).permute(0, 3, 1, 2)


class Decoder(_ConvStack):
    """Decoder.forward of autoencoder.py:236-259 on the sm_100a kernels: z [B, z, h, w] -> image [B, out_ch, 8h, 8w]
    (fp32, NCHW).  Not on the training path (the reference only uses it for reconstruction demos); SURVEY.md 8f-1."""

    def __init__(self, ch: int, out_ch: int, ch_mult: list[int], num_res_blocks: int, in_channels: int, resolution: int,
                 z_channels: int):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.out_ch = resolution, in_channels, out_ch
        self.ffactor = 2 ** (self.num_resolutions - 1)
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            up = nn.Module()
            up.block = block
            up.attn = nn.ModuleList()
            if lvl != 0:
                up.upsample = Upsample(block_in)
            self.up.insert(0, up)  # prepend to get consistent order
        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)
        self._cache = None
        self._cache_key = None

    @torch.no_grad()
    def forward(self, z: Tensor) -> Tensor:
        if z.dim() != 4 or z.shape[1] != self.conv_in.in_channels:
            raise ValueError(f"Decoder expects [B,{self.conv_in.in_channels},h,w] latents, got {tuple(z.shape)}")
        W = self._prepared()
        B, zc, h0, w0 = z.shape
        # conv_in has z_channels (16) inputs: below the 64-channel granularity of the implicit-GEMM conv, so its 3x3
        # patches are unfolded (16*9 = 144 columns of a [B h w, 144] bf16 matrix: 40 KB per image) and fed to the GEMM
        a = torch.nn.functional.unfold(z.float(), 3, padding=1).transpose(1, 2).reshape(B * h0 * w0, zc * 9).to(BF16).contiguous()
        h = K.gemm(a, W["conv_in"][0], bias=W["conv_in"][1]).view(B, h0, w0, -1)
        h = self._res(W, "mid.block_1", self.mid.block_1, h)
        h = self._attn(W, h)
        h = self._res(W, "mid.block_2", self.mid.block_2, h)
        for lvl in reversed(range(self.num_resolutions)):
            for j, blk in enumerate(self.up[lvl].block):
                h = self._res(W, f"up.{lvl}.block.{j}", blk, h)
            if lvl != 0:
                n = f"up.{lvl}.upsample.conv"
                h = K.conv2d_nhwc(K.upsample2x_nhwc(h), W[n][0], 3, 3, 1, 1, bias=W[n][1])
        h = K.groupnorm_swish_nhwc(h, *W["norm_out"])
        y = K.conv2d_nhwc(h, W["conv_out"][0], 3, 3, 1, 1, bias=W["conv_out"][1], act=ACT_NONE, out_dtype=F32)
        return y[..., :self.out_ch].permute(0, 3, 1, 2).contiguous()


class DiagonalGaussian(nn.Module):  # autoencoder.py:262-274
    def __init__(self, sample: bool = True, chunk_dim: int = 1):
        super().__init__()
        self.sample, self.chunk_dim = sample, chunk_dim

    def forward(self, z: Tensor) -> Tensor:
        mean, logvar = torch.chunk(z, 2, dim=self.chunk_dim)
        if not self.sample:
            return mean
        raise RuntimeError("sampling is fused into AutoEncoder.encode / encode_patchified on the B200 path")


class AutoEncoder(nn.Module):
    def __init__(self, params: AutoEncoderParams):
        super().__init__()
        self.params = params
        self.encoder = Encoder(resolution=params.resolution, in_channels=params.in_channels, ch=params.ch,
                               ch_mult=params.ch_mult, num_res_blocks=params.num_res_blocks,
                               z_channels=params.z_channels)
        self.decoder = Decoder(resolution=params.resolution, in_channels=params.in_channels, ch=params.ch,
                               out_ch=params.out_ch, ch_mult=params.ch_mult, num_res_blocks=params.num_res_blocks,
                               z_channels=params.z_channels)
        self.reg = DiagonalGaussian()
        self.scale_factor = params.scale_factor
        self.shift_factor = params.shift_factor

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        # the decoder is not on the training path: a checkpoint without `decoder.*` tensors (encoder-only export) loads
        # under strict=True as well
        if strict and not any(k.startswith("decoder.") for k in state_dict):
            state_dict = {**state_dict, **{f"decoder.{k}": v for k, v in self.decoder.state_dict().items()}}
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    @torch.no_grad()
    def encode_patchified(self, img: Tensor, mean: float = 0.0, std: float = 1.0, noise: Tensor | None = None) -> Tensor:
        """Fused training-step form: raw image (normalisation folded in) -> x_1 fp32 [B, (h/2)(w/2), 4z], i.e.
        ``rearrange(vae.encode(norm(img)), 'b c (h ph) (w pw) -> b (h w) (c ph pw)')`` of
        train_SigLIP_stage1.py:243,246.  ``noise`` defaults to ``torch.randn`` of the reference's shape/dtype
        ([B,z,h,w] fp32 on the device) so the device RNG stream is consumed exactly as ``randn_like(mean)`` does."""
        mom = self.encoder.moments_nhwc(img, mean, std)
        B, h, w, z2 = mom.shape
        if self.reg.sample:
            if noise is None:
                noise = torch.randn(B, z2 // 2, h, w, dtype=F32, device=mom.device)
        else:
            noise = torch.zeros(B, z2 // 2, h, w, dtype=F32, device=mom.device)
            mom = mom.clone()
            mom[..., z2 // 2:] = 0  # std = exp(0) multiplies a zero noise
        return K.ae_sample_patchify(mom, noise.float(), self.scale_factor, self.shift_factor)

    @torch.no_grad()
    def encode(self, x: Tensor, noise: Tensor | None = None) -> Tensor:
        """x: normalised image [B,3,H,W] -> z [B, z, H/8, W/8] fp32 (autoencoder.py:302-305)."""
        B, _, H, Wd = x.shape
        x1 = self.encode_patchified(x, 0.0, 1.0, noise)
        z = self.params.z_channels
        h2, w2 = H // 16, Wd // 16
        return x1.view(B, h2, w2, z, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, z, h2 * 2, w2 * 2)

    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        """z [B, z, h, w] -> image [B, out_ch, 8h, 8w] (autoencoder.py:307-309): z / scale_factor + shift_factor."""
        return self.decoder(z.float() / self.scale_factor + self.shift_factor)

    def forward(self, x: Tensor) -> Tensor:
        return self.decode(self.encode(x))
