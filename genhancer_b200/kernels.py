"""Tensor-level wrappers over the C ABI (no autograd here; see ``ops.py``).

Every function launches hand-written sm_100a kernels on the *current* torch CUDA stream and
never synchronises.  ``LAUNCHES`` counts kernel launches for ``bench.py``'s ``gpu_launches``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, GH_BF16, GH_F32, GemmArgs,
                   check)

LAUNCHES = 0
BF16 = torch.bfloat16
F32 = torch.float32


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ensure(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise _lib.GhError("genhancer_b200 kernels need CUDA tensors (no CPU fallback)")
    _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _dt(t: torch.Tensor) -> int:
    if t.dtype == BF16:
        return GH_BF16
    if t.dtype == F32:
        return GH_F32
    raise _lib.GhError(f"unsupported dtype {t.dtype}")


def _rowmajor2d(t: torch.Tensor, what: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.GhError(f"{what}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: torch.Tensor | None = None, out_dtype: torch.dtype = BF16, alpha: float = 1.0,
         bias: torch.Tensor | None = None, act: int = ACT_NONE, act_grad: bool = False,
         aux_in: torch.Tensor | None = None, aux_out: torch.Tensor | None = None,
         gate: torch.Tensor | None = None, rows_per_batch: int = 0,
         residual: torch.Tensor | None = None) -> torch.Tensor:
    """D[M,N] = epilogue(alpha * A @ B^T).

    ``a`` is [M,K] (or [K,M] when ``a_mn``), ``b`` is [N,K] (or [K,N] when ``b_mn``); both bf16 with unit
    inner stride (row pitch may exceed the row length).  See ``gh_gemm_bf16`` in include/genhancer_b200.h.
    """
    _ensure(a)
    _rowmajor2d(a, "gemm a")
    _rowmajor2d(b, "gemm b")
    if a.dtype != BF16 or b.dtype != BF16:
        raise _lib.GhError("gemm operands must be bf16")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise _lib.GhError(f"gemm: reduction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _rowmajor2d(out, "gemm out")
    if tuple(out.shape) != (M, N):
        raise _lib.GhError(f"gemm: out has shape {tuple(out.shape)}, expected {(M, N)}")
    g = GemmArgs()
    g.a, g.lda, g.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn)
    g.b, g.ldb, g.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn)
    g.d, g.ldd, g.d_dtype = out.data_ptr(), out.stride(0), _dt(out)
    g.M, g.N, g.K = M, N, K
    g.alpha = alpha
    if bias is not None:
        g.bias, g.bias_dtype = bias.data_ptr(), _dt(bias)
    g.act, g.act_grad = act, int(act_grad)
    if aux_in is not None:
        _rowmajor2d(aux_in, "gemm aux_in")
        g.aux_in, g.ld_aux_in = aux_in.data_ptr(), aux_in.stride(0)
    if aux_out is not None:
        _rowmajor2d(aux_out, "gemm aux_out")
        g.aux_out, g.ld_aux_out = aux_out.data_ptr(), aux_out.stride(0)
    if gate is not None:
        _rowmajor2d(gate, "gemm gate")
        g.gate, g.gate_ld, g.rows_per_batch = gate.data_ptr(), gate.stride(0), rows_per_batch
    if residual is not None:
        _rowmajor2d(residual, "gemm residual")
        g.residual, g.ld_res, g.res_dtype = residual.data_ptr(), residual.stride(0), _dt(residual)
    check(_lib.lib().gh_gemm_bf16(C.byref(g), _stream()))
    _count()
    return out


def fm_interp(x1: torch.Tensor, x0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """x_t = bf16((1-t) x1 + t x0); x1/x0 fp32 [B, ...], t fp32 [B]."""
    _ensure(x1)
    assert x1.dtype == F32 and x0.dtype == F32 and t.dtype == F32
    x1c, x0c, tc = x1.contiguous(), x0.contiguous(), t.contiguous()
    B = x1c.shape[0]
    per = x1c.numel() // max(B, 1)
    out = torch.empty(x1c.shape, dtype=BF16, device=x1.device)
    check(_lib.lib().gh_fm_interp_fwd(x1c.data_ptr(), x0c.data_ptr(), tc.data_ptr(), out.data_ptr(), B, per, _stream()))
    _count()
    return out


def fm_mse_loss(pred: torch.Tensor, x0: torch.Tensor, x1: torch.Tensor, grad_scale: float = 1.0,
                want_grad: bool = True):
    """Returns (loss[1] fp32, dpred bf16 or None)."""
    _ensure(pred)
    assert pred.dtype == BF16 and x0.dtype == F32 and x1.dtype == F32
    p, a, b = pred.contiguous(), x0.contiguous(), x1.contiguous()
    loss = torch.zeros(1, dtype=F32, device=pred.device)
    dpred = torch.empty_like(p) if want_grad else None
    check(_lib.lib().gh_fm_mse_loss_fwdbwd(p.data_ptr(), a.data_ptr(), b.data_ptr(), loss.data_ptr(),
                                           dpred.data_ptr() if want_grad else None, grad_scale, p.numel(), _stream()))
    _count()
    return loss, dpred
