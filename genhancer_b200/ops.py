"""Small autograd-aware operators over the sm_100a kernels, for the trainable pieces that sit OUTSIDE the fused
DiT engine: the CLIP projectors (clip_models/CLIP_bank.py:17-28), the VisualPromptAdapter
(train_OpenAICLIP_video_stage1.py:85-97) and LoRA branches.  Activations are bf16; parameters may be fp32
(the reference keeps projector / adapter parameters in fp32) -- they are cast to bf16 for the tensor cores and
their gradients are returned in the parameter's own dtype.
"""
from __future__ import annotations

import torch

from . import kernels as K
from .kernels import ACT_NONE, BF16, F32


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == BF16 else t.to(BF16)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        shp = x.shape
        x2 = _bf(x).reshape(-1, shp[-1]).contiguous()
        wb = _bf(w).contiguous()
        pre = torch.empty(x2.shape[0], w.shape[0], dtype=BF16, device=x.device) if act != ACT_NONE else None
        y = K.gemm(x2, wb, bias=b, act=act, aux_out=pre)
        ctx.save_for_backward(x2, wb, pre)
        ctx.act, ctx.shp, ctx.has_b = act, shp, b is not None
        ctx.dt = (x.dtype, w.dtype, b.dtype if b is not None else None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb, pre = ctx.saved_tensors
        dy2 = _bf(dy).reshape(-1, dy.shape[-1]).contiguous()
        if ctx.act != ACT_NONE:
            dy2 = K.act_bwd(dy2, pre, ctx.act)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dy2, wb, b_mn=True).view(ctx.shp).to(ctx.dt[0])
        if ctx.needs_input_grad[1]:
            dw = K.gemm(dy2, x2, a_mn=True, b_mn=True, out_dtype=F32 if ctx.dt[1] == F32 else BF16)
        if ctx.has_b and ctx.needs_input_grad[2]:
            acc = torch.zeros(wb.shape[0], dtype=F32, device=dy.device)
            K.colsum(dy2, acc)
            db = acc.to(ctx.dt[2])
        return dx, dw, db, None


def linear(x, w, b=None, act: int = ACT_NONE):
    """y = act(x @ w^T + b) on the tcgen05 GEMM (x [..., K], w [N, K]); N % 4 == 0, K % 8 == 0."""
    return _Linear.apply(x, w, b, act)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        x2 = _bf(x).reshape(-1, shp[-1]).contiguous()
        wf, bfp = w.float().contiguous(), b.float().contiguous()
        y, mean, rstd = K.layernorm_fwd(x2, weight=wf, bias=bfp, eps=eps)
        ctx.save_for_backward(x2, wf, mean, rstd)
        ctx.shp, ctx.dt = shp, (x.dtype, w.dtype, b.dtype)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, wf, mean, rstd = ctx.saved_tensors
        dy2 = _bf(dy).reshape(-1, dy.shape[-1]).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.layernorm_bwd_dx(dy2, x2, mean, rstd, weight=wf).view(ctx.shp).to(ctx.dt[0])
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            acc = torch.zeros(2, x2.shape[1], dtype=F32, device=dy.device)
            K.layernorm_bwd_params(dy2, x2, mean, rstd, acc[0], acc[1])
            db, dw = acc[0].to(ctx.dt[2]), acc[1].to(ctx.dt[1])
        return dx, dw, db, None


def layer_norm(x, w, b, eps: float = 1e-5):
    """Affine LayerNorm over the last dim (C % 8 == 0, C <= 4096), bf16 output."""
    return _LayerNorm.apply(x, w, b, eps)
