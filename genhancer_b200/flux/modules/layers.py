"""Parameter containers with the reference's module tree / state_dict names
(/root/reference/Continuous/src/flux/modules/layers.py), plus the small standalone operators.

The blocks do not run their own ``forward``: ``Flux.forward`` hands the whole parameter set to the fused
engine (``genhancer_b200.flux.engine``), which schedules the sm_100a kernels for all blocks at once.
Dead reference classes (LoRA / IP-adapter processors, ImageProjModel) are not reproduced.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from ... import kernels as K


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


class RMSNorm(nn.Module):  # layers.py:63-72
    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))


class QKNorm(nn.Module):  # layers.py:75-84
    def __init__(self, dim: int):
        super().__init__()
        self.query_norm = RMSNorm(dim)
        self.key_norm = RMSNorm(dim)


class SelfAttention(nn.Module):  # layers.py:142-152
    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.norm = QKNorm(dim // num_heads)
        self.proj = nn.Linear(dim, dim)


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


class LastLayer(nn.Module):  # layers.py:561-572
    def __init__(self, hidden_size: int, patch_size: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))
