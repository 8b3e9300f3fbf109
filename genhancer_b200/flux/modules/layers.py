"""Parameter containers with the reference's module tree / state_dict names
(/root/reference/Continuous/src/flux/modules/layers.py), plus the small standalone operators.

TRAINING goes through ``Flux.forward``, which hands the whole parameter set to the fused engine
(``genhancer_b200.flux.engine``: one explicit forward / backward schedule for all blocks at once).  Every module
here also has the reference's own ``forward`` (same arguments, same return structure), built from the same sm_100a
kernels, so ``dit.time_in(x)``, ``block(img, txt, vec, pe)``, ``dit.final_layer(x, vec)`` ... can be called one at a
time as in the reference: ``MLPEmbedder`` is autograd-aware (``ops.linear``); the blocks / attention / last layer are
forward-only (``torch.no_grad``: their backward lives in the engine).
Dead reference classes (LoRA / IP-adapter processors, ImageProjModel) are not reproduced.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from ... import kernels as K
from ... import ops
from ...kernels import ACT_GELU_TANH, ACT_SILU, BF16


def _cs_from_pe(pe: Tensor) -> Tensor:
    """The reference's rotation table [B|1, 1, L, D/2, 2, 2] ((cos, -sin), (sin, cos)) -> what the fused QK-norm + RoPE
    kernel takes: [B|1, L, D/2, 2] (cos, sin), fp32."""
    return torch.stack([pe[:, 0, :, :, 0, 0], pe[:, 0, :, :, 1, 0]], dim=-1).float().contiguous()


def _adaln(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """(1 + scale) * LayerNorm(x; eps 1e-6, no affine) + shift, one kernel (shift / scale: [B, C] rows)."""
    return K.layernorm_fwd(x.to(BF16).contiguous(), shift=shift, scale=scale, eps=1e-6, save_stats=False)[0]


def _qkv_to_heads(qkv: Tensor, H: int, norm: "QKNorm", cs: Tensor, q: Tensor, k: Tensor, v: Tensor, l_off: int) -> None:
    K.qk_norm_rope_fwd(qkv, H, norm.query_norm.scale, norm.key_norm.scale, cs, q, k, v, l_off)


class EmbedND(nn.Module):
    """layers.py:11-25.  forward(ids [B,L,3]) -> rotation table [B,1,L,sum(axes)/2,2,2] fp32 (reference layout)."""

    def __init__(self, dim: int, theta: int, axes_dim: list[int]):
        super().__init__()
        self.dim, self.theta, self.axes_dim = dim, theta, axes_dim

    def forward(self, ids: Tensor) -> Tensor:
        cs = K.rope_table(ids, tuple(self.axes_dim), float(self.theta))  # [B, L, half, 2] (cos, sin)
        c, s = cs[..., 0], cs[..., 1]
        return torch.stack([c, -s, s, c], dim=-1).reshape(*c.shape, 2, 2).unsqueeze(1)


def timestep_embedding(t: Tensor, dim: int = 256, max_period: int = 10000, time_factor: float = 1000.0) -> Tensor:
    """layers.py:28-49 for dim=256 (the only size the model uses); returns bf16 [N, 256]."""
    if dim != 256 or max_period != 10000 or time_factor != 1000.0:
        raise NotImplementedError("genhancer_b200 implements the model's timestep_embedding(t, 256) only")
    return K.timestep_embedding(t.float(), round_bf16=(t.dtype != torch.float32))


class MLPEmbedder(nn.Module):  # layers.py:52-60
    def __init__(self, in_dim: int, hidden_dim: int):
        super().__init__()
        self.in_layer = nn.Linear(in_dim, hidden_dim, bias=True)
        self.silu = nn.SiLU()
        self.out_layer = nn.Linear(hidden_dim, hidden_dim, bias=True)

    def forward(self, x: Tensor) -> Tensor:
        h = ops.linear(x, self.in_layer.weight, self.in_layer.bias, act=ACT_SILU)
        return ops.linear(h, self.out_layer.weight, self.out_layer.bias)


class RMSNorm(nn.Module):  # layers.py:63-72
    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))

    def forward(self, x: Tensor) -> Tensor:   # (standalone form; inside the blocks it is fused with RoPE: gh_qk_norm_rope_fwd)
        x_dtype = x.dtype
        x = x.float()
        rrms = torch.rsqrt(torch.mean(x ** 2, dim=-1, keepdim=True) + 1e-6)
        return (x * rrms).to(dtype=x_dtype) * self.scale


class QKNorm(nn.Module):  # layers.py:75-84
    def __init__(self, dim: int):
        super().__init__()
        self.query_norm = RMSNorm(dim)
        self.key_norm = RMSNorm(dim)

    def forward(self, q: Tensor, k: Tensor, v: Tensor) -> tuple[Tensor, Tensor]:
        return self.query_norm(q).to(v), self.key_norm(k).to(v)


class SelfAttention(nn.Module):  # layers.py:142-152
    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.norm = QKNorm(dim // num_heads)
        self.proj = nn.Linear(dim, dim)

    @torch.no_grad()
    def forward(self, x: Tensor, pe: Tensor) -> Tensor:
        B, L, C = x.shape
        H, D = self.num_heads, C // self.num_heads
        qkv = K.gemm(x.to(BF16).reshape(-1, C).contiguous(), self.qkv.weight, bias=self.qkv.bias).view(B, L, 3 * C)
        q = torch.empty(B, H, L, D, dtype=BF16, device=x.device)
        k, v = torch.empty_like(q), torch.empty_like(q)
        _qkv_to_heads(qkv, H, self.norm, _cs_from_pe(pe), q, k, v, 0)
        o = torch.empty(B, L, C, dtype=BF16, device=x.device)
        K.flash_attn_fwd(q, k, v, D ** -0.5, o, want_lse=False)
        return K.gemm(o.view(-1, C), self.proj.weight, bias=self.proj.bias).view(B, L, C)


@dataclass
class ModulationOut:
    shift: Tensor
    scale: Tensor
    gate: Tensor


class Modulation(nn.Module):  # layers.py:162-175
    def __init__(self, dim: int, double: bool):
        super().__init__()
        self.is_double = double
        self.multiplier = 6 if double else 3
        self.lin = nn.Linear(dim, self.multiplier * dim, bias=True)

    def forward(self, vec: Tensor) -> tuple[ModulationOut, ModulationOut | None]:
        out = ops.linear(nn.functional.silu(vec), self.lin.weight, self.lin.bias)[:, None, :].chunk(self.multiplier, dim=-1)
        return ModulationOut(*out[:3]), ModulationOut(*out[3:]) if self.is_double else None


class DoubleStreamBlock(nn.Module):  # layers.py:339-389
    def __init__(self, hidden_size: int, num_heads: int, mlp_ratio: float, qkv_bias: bool = False):
        super().__init__()
        mlp_hidden_dim = int(hidden_size * mlp_ratio)
        self.num_heads, self.hidden_size, self.head_dim = num_heads, hidden_size, hidden_size // num_heads
        self.img_mod = Modulation(hidden_size, double=True)
        self.img_norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.img_attn = SelfAttention(dim=hidden_size, num_heads=num_heads, qkv_bias=qkv_bias)
        self.img_norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.img_mlp = nn.Sequential(nn.Linear(hidden_size, mlp_hidden_dim, bias=True), nn.GELU(approximate="tanh"),
                                     nn.Linear(mlp_hidden_dim, hidden_size, bias=True))
        self.txt_mod = Modulation(hidden_size, double=True)
        self.txt_norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.txt_attn = SelfAttention(dim=hidden_size, num_heads=num_heads, qkv_bias=qkv_bias)
        self.txt_norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.txt_mlp = nn.Sequential(nn.Linear(hidden_size, mlp_hidden_dim, bias=True), nn.GELU(approximate="tanh"),
                                     nn.Linear(mlp_hidden_dim, hidden_size, bias=True))

    @torch.no_grad()
    def forward(self, img: Tensor, txt: Tensor, vec: Tensor, pe: Tensor, image_proj: Tensor | None = None,
                ip_scale: float = 1.0) -> tuple[Tensor, Tensor]:
        """DoubleStreamBlockProcessor.__call__ (layers.py:303-337) on the sm_100a kernels, forward only."""
        if image_proj is not None:
            raise NotImplementedError("IP-adapter inputs are unused by every GenHancer training script")
        B, Li, C = img.shape
        Lt, H, D = txt.shape[1], self.num_heads, self.head_dim
        dev = img.device
        mods = {"img": self.img_mod(vec), "txt": self.txt_mod(vec)}
        xs = {"img": img.to(BF16).contiguous(), "txt": txt.to(BF16).contiguous()}
        cs = _cs_from_pe(pe)
        q = torch.empty(B, H, Lt + Li, D, dtype=BF16, device=dev)
        k, v = torch.empty_like(q), torch.empty_like(q)
        for s, l_off in (("txt", 0), ("img", Lt)):            # joint sequence: txt first, then img (layers.py:323-326)
            m1, att = mods[s][0], getattr(self, f"{s}_attn")
            h = _adaln(xs[s], m1.shift[:, 0], m1.scale[:, 0])
            qkv = K.gemm(h.view(-1, C), att.qkv.weight, bias=att.qkv.bias).view(B, -1, 3 * C)
            _qkv_to_heads(qkv, H, att.norm, cs, q, k, v, l_off)
        attn = {"txt": torch.empty(B, Lt, C, dtype=BF16, device=dev), "img": torch.empty(B, Li, C, dtype=BF16, device=dev)}
        K.flash_attn_fwd(q, k, v, D ** -0.5, attn["img"], attn["txt"], Lt, want_lse=False)
        out = {}
        for s, rpb in (("img", Li), ("txt", Lt)):
            (m1, m2), att, mlp = mods[s], getattr(self, f"{s}_attn"), getattr(self, f"{s}_mlp")
            x = K.gemm(attn[s].view(-1, C), att.proj.weight, bias=att.proj.bias, gate=m1.gate[:, 0], rows_per_batch=rpb,
                       residual=xs[s].view(-1, C)).view(B, rpb, C)
            h2 = _adaln(x, m2.shift[:, 0], m2.scale[:, 0])
            a = K.gemm(h2.view(-1, C), mlp[0].weight, bias=mlp[0].bias, act=ACT_GELU_TANH)
            out[s] = K.gemm(a, mlp[2].weight, bias=mlp[2].bias, gate=m2.gate[:, 0], rows_per_batch=rpb,
                            residual=x.view(-1, C)).view(B, rpb, C)
        return out["img"], out["txt"]


class SingleStreamBlock(nn.Module):  # layers.py:503-557
    def __init__(self, hidden_size: int, num_heads: int, mlp_ratio: float = 4.0, qk_scale: float | None = None):
        super().__init__()
        self.hidden_dim = self.hidden_size = hidden_size
        self.num_heads = num_heads
        self.head_dim = hidden_size // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.mlp_hidden_dim = int(hidden_size * mlp_ratio)
        self.linear1 = nn.Linear(hidden_size, hidden_size * 3 + self.mlp_hidden_dim)
        self.linear2 = nn.Linear(hidden_size + self.mlp_hidden_dim, hidden_size)
        self.norm = QKNorm(self.head_dim)
        self.pre_norm = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp_act = nn.GELU(approximate="tanh")
        self.modulation = Modulation(hidden_size, double=False)

    @torch.no_grad()
    def forward(self, x: Tensor, vec: Tensor, pe: Tensor) -> Tensor:
        """SingleStreamBlockProcessor.__call__ (layers.py:485-501) on the sm_100a kernels, forward only."""
        B, L, C = x.shape
        H, D, mlp = self.num_heads, self.head_dim, self.mlp_hidden_dim
        x = x.to(BF16).contiguous()
        mod, _ = self.modulation(vec)
        h = _adaln(x, mod.shift[:, 0], mod.scale[:, 0])
        w1, b1 = self.linear1.weight, self.linear1.bias
        qkv = K.gemm(h.view(-1, C), w1[:3 * C], bias=b1[:3 * C]).view(B, L, 3 * C)
        cat = torch.empty(B * L, C + mlp, dtype=BF16, device=x.device)       # [attn | gelu(mlp)]: linear2 needs no concat
        K.gemm(h.view(-1, C), w1[3 * C:], bias=b1[3 * C:], act=ACT_GELU_TANH, out=cat[:, C:])
        q = torch.empty(B, H, L, D, dtype=BF16, device=x.device)
        k, v = torch.empty_like(q), torch.empty_like(q)
        _qkv_to_heads(qkv, H, self.norm, _cs_from_pe(pe), q, k, v, 0)
        K.flash_attn_fwd(q, k, v, self.scale, cat.view(B, L, C + mlp)[:, :, :C], want_lse=False)
        return K.gemm(cat, self.linear2.weight, bias=self.linear2.bias, gate=mod.gate[:, 0], rows_per_batch=L,
                      residual=x.view(-1, C)).view(B, L, C)


class LastLayer(nn.Module):  # layers.py:561-572
    def __init__(self, hidden_size: int, patch_size: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))

    @torch.no_grad()
    def forward(self, x: Tensor, vec: Tensor) -> Tensor:
        B, L, C = x.shape
        lin = self.adaLN_modulation[1]
        m = K.gemm(K.act_fwd(vec.to(BF16).contiguous(), ACT_SILU), lin.weight, bias=lin.bias)      # [B, 2C]: shift, scale
        h = _adaln(x, m[:, :C], m[:, C:])
        return K.gemm(h.view(-1, C), self.linear.weight, bias=self.linear.bias).view(B, L, -1)
