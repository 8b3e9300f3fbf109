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
# Host-side launch policy (NOT library state): parallel.GradReducer sets it while gradient buckets are in flight, every
# GEMM launched meanwhile asks for the dynamic tile schedule through gh_gemm_args::dynamic_tiles.
DYNAMIC_TILES = False
GEMM_TIMER = None  # bench.py: callable(flops) -> (start_event, stop_event) recorded around every tcgen05 GEMM/conv launch
BF16 = torch.bfloat16
F32 = torch.float32


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    # the raw handle of torch's current stream; torch.cuda.current_stream() builds a Stream object (~14 us per call,
    # 10 ms of host time per step over ~700 launches)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


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
         residual: torch.Tensor | None = None, a2: torch.Tensor | None = None,
         b2: torch.Tensor | None = None, k_splits: int = 0, batch: int = 1) -> torch.Tensor:
    """D[M,N] = epilogue(alpha * (A @ B^T + A2 @ B2^T)).

    ``batch`` > 1: ``a``, ``b`` and ``out`` are ``batch`` equal problems stacked along their FIRST dimension
    ([batch*M, K] / [batch*N, K] -> [batch*M, N], or the ``*_mn`` forms stacked along K); one launch computes them all.

    ``a2`` / ``b2`` (same majors as ``a`` / ``b``, reduction length K2 = the LoRA rank) fold a low-rank branch into
    the base GEMM: one extra 16-deep MMA per tile instead of a second GEMM and a second pass over the output.

    ``a`` is [M,K] (or [K,M] when ``a_mn``), ``b`` is [N,K] (or [K,N] when ``b_mn``); both bf16 with unit
    inner stride (row pitch may exceed the row length).  See ``gh_gemm_bf16`` in include/genhancer_b200.h.
    """
    _ensure(a)
    _rowmajor2d(a, "gemm a")
    _rowmajor2d(b, "gemm b")
    if a.dtype != BF16 or b.dtype != BF16:
        raise _lib.GhError("gemm operands must be bf16")
    if batch < 1 or a.shape[0] % batch or b.shape[0] % batch:
        raise _lib.GhError(f"gemm: batch={batch} does not divide the stacked operands {tuple(a.shape)} / {tuple(b.shape)}")
    a_rows, b_rows = a.shape[0] // batch, b.shape[0] // batch
    M, K = (a.shape[1], a_rows) if a_mn else (a_rows, a.shape[1])
    N, Kb = (b.shape[1], b_rows) if b_mn else (b_rows, b.shape[1])
    if K != Kb:
        raise _lib.GhError(f"gemm: reduction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((batch * M, N), dtype=out_dtype, device=a.device)
    _rowmajor2d(out, "gemm out")
    if tuple(out.shape) != (batch * M, N):
        raise _lib.GhError(f"gemm: out has shape {tuple(out.shape)}, expected {(batch * M, N)}")
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
    if batch > 1:
        g.batch, g.a_batch_rows, g.b_batch_rows, g.d_batch_rows = batch, a_rows, b_rows, M
    g.dynamic_tiles = int(DYNAMIC_TILES)
    g.k_splits = k_splits   # != 0: out (fp32, zeroed or to be accumulated into by the caller) += A @ B^T, split over K
    K2 = 0
    if a2 is not None:
        _rowmajor2d(a2, "gemm a2")
        _rowmajor2d(b2, "gemm b2")
        if a2.dtype != BF16 or b2.dtype != BF16:
            raise _lib.GhError("gemm a2/b2 must be bf16")
        M2, K2 = (a2.shape[1], a2.shape[0]) if a_mn else (a2.shape[0], a2.shape[1])
        N2, K2b = (b2.shape[1], b2.shape[0]) if b_mn else (b2.shape[0], b2.shape[1])
        if (M2, N2) != (M, N) or K2 != K2b:
            raise _lib.GhError(f"gemm: second operand pair is {M2}x{K2} / {N2}x{K2b}, expected M={M} N={N}")
        g.a2, g.lda2, g.b2, g.ldb2, g.K2 = a2.data_ptr(), a2.stride(0), b2.data_ptr(), b2.stride(0), K2
    tm = GEMM_TIMER
    if tm is not None:
        ev0, ev1 = tm(2.0 * batch * M * N * (K + K2), f"gemm M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} d={'f32' if out.dtype == F32 else 'bf16'}"
                      + (f" batch={batch}" if batch > 1 else ""))
        ev0.record()
    check(_lib.lib().gh_gemm_bf16(C.byref(g), _stream()))
    if tm is not None:
        ev1.record()
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


# ----------------------------------------------------------------------------------------------
# row-view helpers (see gh_rows_view in include/genhancer_b200.h)
# ----------------------------------------------------------------------------------------------
from ._lib import AttnOut, AttnTensor, RowsView  # noqa: E402


def _rows_view(t: torch.Tensor) -> tuple[RowsView, int, int, int]:
    """-> (view, rows, C, batches) for a [B, L, C] or [R, C] tensor with unit channel stride."""
    if t.stride(-1) != 1:
        raise _lib.GhError(f"channel dim must be contiguous, got strides {t.stride()}")
    if t.dim() == 3:
        B, L, Cc = t.shape
        return RowsView(L, t.stride(0), t.stride(1)), B * L, Cc, B
    if t.dim() == 2:
        R, Cc = t.shape
        return RowsView(max(R, 1), 0, t.stride(0)), R, Cc, 1
    raise _lib.GhError(f"expected 2-D or 3-D tensor, got {tuple(t.shape)}")


def _p(t):
    return None if t is None else t.data_ptr()


def layernorm_fwd(x, weight=None, bias=None, shift=None, scale=None, eps=1e-6, out=None, save_stats=True):
    """bf16 LayerNorm over the last dim; affine (fp32 w,b) | AdaLN (bf16 shift/scale [B,C] rows of a wider tensor)."""
    _ensure(x)
    xv, rows, Cc, _ = _rows_view(x)
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    yv, rows_y, _, _ = _rows_view(out)
    assert rows_y == rows and x.dtype == BF16 and out.dtype == BF16
    mean = torch.empty(rows, dtype=F32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=F32, device=x.device) if save_stats else None
    mod_ld = 0
    if scale is not None:
        assert shift is not None and scale.dtype == BF16 and shift.dtype == BF16 and scale.stride(-1) == 1
        assert scale.stride(0) == shift.stride(0)
        mod_ld = scale.stride(0)
    if weight is not None:
        assert weight.dtype == F32 and bias is not None and bias.dtype == F32
    check(_lib.lib().gh_layernorm_fwd(x.data_ptr(), C.byref(xv), out.data_ptr(), C.byref(yv), rows, Cc, _p(weight),
                                      _p(bias), _p(shift), _p(scale), mod_ld, eps, _p(mean), _p(rstd), _stream()))
    _count()
    return out, mean, rstd


def layernorm_bwd_dx(dy, x, mean, rstd, weight=None, scale=None, dres=None, out=None):
    _ensure(dy)
    dyv, rows, Cc, _ = _rows_view(dy)
    xv, _, _, _ = _rows_view(x)
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    dxv, _, _, _ = _rows_view(out)
    drv = _rows_view(dres)[0] if dres is not None else None
    mod_ld = scale.stride(0) if scale is not None else 0
    check(_lib.lib().gh_layernorm_bwd_dx(dy.data_ptr(), C.byref(dyv), x.data_ptr(), C.byref(xv), rows, Cc,
                                         mean.data_ptr(), rstd.data_ptr(), _p(weight), _p(scale), mod_ld, _p(dres),
                                         C.byref(drv) if drv is not None else None, out.data_ptr(), C.byref(dxv),
                                         _stream()))
    _count()
    return out


def layernorm_bwd_params(dy, x, mean, rstd, dshift_acc, dscale_acc):
    """dshift_acc/dscale_acc: fp32 [B, C] views (row pitch = stride(0)) accumulated with atomics."""
    _ensure(dy)
    dyv, _, Cc, nb = _rows_view(dy)
    xv = _rows_view(x)[0]
    assert dshift_acc.dtype == F32 and dscale_acc.dtype == F32
    acc_ld = dshift_acc.stride(0) if dshift_acc.dim() == 2 else 0
    if dscale_acc.dim() == 2:
        assert dscale_acc.stride(0) == acc_ld
    check(_lib.lib().gh_layernorm_bwd_params(dy.data_ptr(), C.byref(dyv), x.data_ptr(), C.byref(xv), nb, Cc,
                                             mean.data_ptr(), rstd.data_ptr(), dshift_acc.data_ptr(),
                                             dscale_acc.data_ptr(), acc_ld, _stream()))
    _count()


def gate_bwd(dout, u, gate, dgate_acc, out=None, dbias_acc=None):
    """out = res + gate[b]*u  ->  du = gate[b]*dout (returned), dgate_acc[b,:] += sum_l dout*u;
    ``dbias_acc`` (fp32 [C]) += column sums of du (the bias gradient of the linear that produced u)."""
    _ensure(dout)
    dov, _, Cc, nb = _rows_view(dout)
    uv = _rows_view(u)[0]
    if out is None:
        out = torch.empty(dout.shape, dtype=BF16, device=dout.device)
    duv = _rows_view(out)[0]
    check(_lib.lib().gh_gate_bwd(dout.data_ptr(), C.byref(dov), u.data_ptr(), C.byref(uv), nb, Cc, gate.data_ptr(),
                                 gate.stride(0), out.data_ptr(), C.byref(duv), dgate_acc.data_ptr(),
                                 dgate_acc.stride(0) if dgate_acc.dim() == 2 else 0, _p(dbias_acc), _stream()))
    _count()
    return out


def colsum(dy, acc):
    """acc (fp32 [C] or [B, C]) += column sums of dy ([R, C] or [B, L, C])."""
    _ensure(dy)
    dyv, _, Cc, nb = _rows_view(dy)
    if acc.dim() == 1:
        if dy.dim() == 3:  # reduce over batch too: flatten needs contiguity
            dyc = dy.reshape(-1, Cc)
            dyv, _, Cc, nb = _rows_view(dyc)
            dy = dyc
        acc_ld = 0
    else:
        acc_ld = acc.stride(0)
    check(_lib.lib().gh_colsum(dy.data_ptr(), C.byref(dyv), nb, Cc, acc.data_ptr(), acc_ld, _stream()))
    _count()


def rope_table(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10_000.0) -> torch.Tensor:
    """ids [..., L, 3] (any float dtype, integer-valued) -> fp32 [..., L, sum(axes)/2, 2] (cos, sin)."""
    _ensure(ids)
    idf = ids.to(F32).contiguous()
    n_tok = idf.numel() // 3
    half = sum(axes_dim) // 2
    out = torch.empty(*idf.shape[:-1], half, 2, dtype=F32, device=ids.device)
    check(_lib.lib().gh_rope_table(idf.data_ptr(), out.data_ptr(), n_tok, axes_dim[0], axes_dim[1], axes_dim[2],
                                   float(theta), _stream()))
    _count()
    return out


def qk_norm_rope_fwd(qkv, H, q_scale, k_scale, cs, q, k, v, l_off):
    """qkv [B, L, 3*H*D] (row pitch free) -> writes q,k,v [B, H, Ltot, D] at token offset l_off. cs [B|1, Ltot, D/2, 2]."""
    _ensure(qkv)
    B, L, _ = qkv.shape
    D = q.shape[-1]
    Ltot = q.shape[2]
    assert qkv.stride(2) == 1 and qkv.stride(0) == L * qkv.stride(1)
    assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    csb = cs.stride(0) // 2 if cs.shape[0] > 1 else 0  # in float2 units
    check(_lib.lib().gh_qk_norm_rope_fwd(qkv.data_ptr(), qkv.stride(1), B, L, H, D, Ltot, l_off, q_scale.data_ptr(),
                                         k_scale.data_ptr(), cs.data_ptr(), csb, q.data_ptr(), k.data_ptr(),
                                         v.data_ptr(), _stream()))
    _count()


def qk_norm_rope_bwd(dq, dk, dv, qkv, H, q_scale, k_scale, cs, l_off, dqkv, dscale_q_acc, dscale_k_acc):
    _ensure(qkv)
    B, L, _ = qkv.shape
    D = dq.shape[-1]
    Ltot = dq.shape[2]
    assert dq.is_contiguous() and dk.is_contiguous() and dv.is_contiguous()
    csb = cs.stride(0) // 2 if cs.shape[0] > 1 else 0
    check(_lib.lib().gh_qk_norm_rope_bwd(dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), qkv.data_ptr(), qkv.stride(1),
                                         B, L, H, D, Ltot, l_off, q_scale.data_ptr(), k_scale.data_ptr(),
                                         cs.data_ptr(), csb, dqkv.data_ptr(), dqkv.stride(1),
                                         dscale_q_acc.data_ptr(), dscale_k_acc.data_ptr(), _stream()))
    _count()


def timestep_embedding(t: torch.Tensor, round_bf16: bool = True) -> torch.Tensor:
    _ensure(t)
    tc = t.to(F32).contiguous()
    out = torch.empty(tc.shape[0], 256, dtype=BF16, device=t.device)
    check(_lib.lib().gh_timestep_embedding(tc.data_ptr(), out.data_ptr(), tc.shape[0], int(round_bf16), _stream()))
    _count()
    return out


def act_fwd(x: torch.Tensor, act: int) -> torch.Tensor:
    _ensure(x)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    check(_lib.lib().gh_act_fwd(xc.data_ptr(), out.data_ptr(), xc.numel(), act, _stream()))
    _count()
    return out


def act_bwd(dy: torch.Tensor, x: torch.Tensor, act: int) -> torch.Tensor:
    _ensure(x)
    dyc, xc = dy.contiguous(), x.contiguous()
    out = torch.empty_like(xc)
    check(_lib.lib().gh_act_bwd(dyc.data_ptr(), xc.data_ptr(), out.data_ptr(), xc.numel(), act, _stream()))
    _count()
    return out


def accum_cast(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0, accumulate: bool = False) -> None:
    _ensure(src)
    assert src.dtype == F32 and src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    check(_lib.lib().gh_accum_cast(src.data_ptr(), dst.data_ptr(), _dt(dst), src.numel(), scale, int(accumulate),
                                   _stream()))
    _count()


def _attn_tensor(t: torch.Tensor) -> AttnTensor:
    """t viewed as [B, H, L, D] with unit stride on D."""
    assert t.dim() == 4 and t.stride(3) == 1 and t.dtype == BF16
    return AttnTensor(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def flash_attn_fwd(q, k, v, scale, out1, out0=None, n_split=0, want_lse=True, d_valid=0):
    """q,k,v: [B,H,L,D] views (any strides, D contiguous). out1/out0: token-major [B, L(seg), H*D] (row pitch free).
    Rows l < n_split go to out0.  Returns lse2 [B,H,Lq] fp32 (log2 domain) or None.
    ``d_valid``: leading lanes of D that hold data (heads zero-padded into a wider slot); 0 = all."""
    _ensure(q)
    B, H, Lq, D = q.shape
    Lk = k.shape[2]
    at = [_attn_tensor(x) for x in (q, k, v)]
    o = AttnOut()
    if out0 is not None and n_split > 0:
        assert out0.stride(2) == 1
        o.seg0, o.seg0_batch_stride, o.seg0_row_stride = out0.data_ptr(), out0.stride(0), out0.stride(1)
    assert out1.stride(2) == 1
    o.seg1, o.seg1_batch_stride, o.seg1_row_stride = out1.data_ptr(), out1.stride(0), out1.stride(1)
    o.n_split = n_split
    lse = torch.empty(B, H, Lq, dtype=F32, device=q.device) if want_lse else None
    check(_lib.lib().gh_flash_attn_fwd(C.byref(at[0]), C.byref(at[1]), C.byref(at[2]), B, H, Lq, Lk, D, _dvalid(d_valid, D),
                                       float(scale), C.byref(o), _p(lse), _stream()))
    _count()
    return lse


def conv2d_nhwc(x, w, KH, KW, stride=1, pad=0, Ho=None, Wo=None, bias=None, act=ACT_NONE, residual=None,
                out_dtype=BF16, out=None):
    """x bf16 NHWC [B,H,W,Cin]; w bf16 [Cout, KH*KW*Cin] (k = (kh*KW+kw)*Cin + ci) -> [B,Ho,Wo,Cout]."""
    from ._lib import ConvArgs
    _ensure(x)
    assert x.dtype == BF16 and w.dtype == BF16 and x.is_contiguous() and w.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == KH * KW * Cin
    if Ho is None:
        Ho = (H + 2 * pad - KH) // stride + 1
        Wo = (W + 2 * pad - KW) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, Cout, dtype=out_dtype, device=x.device)
    a = ConvArgs()
    a.x, a.w, a.y, a.y_dtype = x.data_ptr(), w.data_ptr(), out.data_ptr(), _dt(out)
    a.B, a.H, a.W, a.Cin, a.Cout, a.KH, a.KW, a.stride, a.pad, a.Ho, a.Wo = B, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo
    if bias is not None:
        a.bias, a.bias_dtype = bias.data_ptr(), _dt(bias)
    a.act = act
    if residual is not None:
        assert residual.is_contiguous() and tuple(residual.shape) == tuple(out.shape)
        a.residual, a.res_dtype = residual.data_ptr(), _dt(residual)
    tm = GEMM_TIMER
    if tm is not None:
        ev0, ev1 = tm(2.0 * B * Ho * Wo * Cout * KH * KW * Cin, f"conv B={B} H={H} W={W} Cin={Cin} Cout={Cout} k={KH} s={stride}")
        ev0.record()
    check(_lib.lib().gh_conv2d_nhwc(C.byref(a), _stream()))
    if tm is not None:
        ev1.record()
    _count()
    return out


def is_u8_image(img: torch.Tensor) -> bool:
    """A decoded RGB batch as the loader's workers produce it: uint8 [B, H, W, 3]."""
    return img.dtype == torch.uint8 and img.dim() == 4 and img.shape[-1] == 3


def image_bhw(img: torch.Tensor) -> tuple[int, int, int]:
    """(B, H, W) of an fp32 NCHW [B,3,H,W] or uint8 HWC [B,H,W,3] batch."""
    if is_u8_image(img):
        return img.shape[0], img.shape[1], img.shape[2]
    if img.dim() != 4 or img.shape[1] != 3:
        raise _lib.GhError(f"expected an image batch [B,3,H,W] (float) or [B,H,W,3] (uint8), got {img.dtype} {tuple(img.shape)}")
    return img.shape[0], img.shape[2], img.shape[3]


def patch_im2col(img, patch, ld, mean=None, std=None):
    """img fp32 NCHW [B,3,S,S] in [0,1] -- or the decoded uint8 HWC [B,S,S,3] batch itself (u8 / 255 on the fly) --
    -> bf16 [B*G*G, ld] view of width 3*p*p (row pitch ld)."""
    _ensure(img)
    u8 = is_u8_image(img)
    assert (u8 or img.dtype == F32) and img.is_contiguous()
    B, S, _ = image_bhw(img)
    G = S // patch
    buf = torch.empty(B * G * G, ld, dtype=BF16, device=img.device)
    m3 = (C.c_float * 3)(*mean) if mean is not None else None
    s3 = (C.c_float * 3)(*std) if std is not None else None
    fn = _lib.lib().gh_patch_im2col_u8hwc if u8 else _lib.lib().gh_patch_im2col
    check(fn(img.data_ptr(), buf.data_ptr(), B, S, patch, ld, m3, s3, _stream()))
    _count()
    return buf[:, : 3 * patch * patch]


def patch_embed(img, w, patch, bias=None, mean=None, std=None):
    """Patch-embed conv as an implicit GEMM (gh_patch_embed_fwd): img fp32 NCHW [B,3,S,S] in [0,1] or uint8 HWC
    [B,S,S,3]; w bf16 [D, ld] view of width 3*p*p (pad columns zero); -> bf16 [B*(S/p)^2, D]."""
    _ensure(img)
    u8 = is_u8_image(img)
    assert (u8 or img.dtype == F32) and img.is_contiguous() and w.dtype == BF16 and w.stride(1) == 1
    B, S, _ = image_bhw(img)
    G, D = S // patch, w.shape[0]
    out = torch.empty(B * G * G, D, dtype=BF16, device=img.device)
    m3 = (C.c_float * 3)(*mean) if mean is not None else None
    s3 = (C.c_float * 3)(*std) if std is not None else None
    if bias is not None:
        assert bias.dtype == F32 and bias.is_contiguous()
    check(_lib.lib().gh_patch_embed_fwd(img.data_ptr(), int(u8), w.data_ptr(), w.stride(0), _p(bias), out.data_ptr(), D, B, S,
                                        patch, D, m3, s3, _stream()))
    _count()
    return out


def im2col3x3_c3(img, mean=0.0, std=1.0):
    """fp32 NCHW [B,3,H,W] or uint8 HWC [B,H,W,3] -> bf16 [B*H*W, 32] (the AE conv_in gather, normalisation fused)."""
    _ensure(img)
    u8 = is_u8_image(img)
    assert (u8 or img.dtype == F32) and img.is_contiguous()
    B, H, W = image_bhw(img)
    out = torch.empty(B * H * W, 32, dtype=BF16, device=img.device)
    fn = _lib.lib().gh_im2col3x3_c3_u8hwc if u8 else _lib.lib().gh_im2col3x3_c3
    check(fn(img.data_ptr(), out.data_ptr(), B, H, W, float(mean), float(std), _stream()))
    _count()
    return out


def embed_assemble(patch, cls, pos, B, T, D):
    _ensure(patch)
    out = torch.empty(B, T, D, dtype=BF16, device=patch.device)
    check(_lib.lib().gh_embed_assemble(patch.data_ptr(), _p(cls), pos.data_ptr(), out.data_ptr(), B, T, D,
                                       0 if cls is None else 1, _stream()))
    _count()
    return out


def groupnorm_swish_nhwc(x, weight, bias, eps=1e-6, swish=True):
    _ensure(x)
    assert x.dtype == BF16 and x.is_contiguous() and weight.dtype == F32 and bias.dtype == F32
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    ws = torch.empty(_lib.lib().gh_groupnorm_ws_bytes(B, H * W) // 4, dtype=F32, device=x.device)
    check(_lib.lib().gh_groupnorm_swish_nhwc(x.data_ptr(), y.data_ptr(), B, H * W, Cc, weight.data_ptr(), bias.data_ptr(),
                                             eps, int(swish), ws.data_ptr(), _stream()))
    _count(3)
    return y


def u8hwc_to_f32chw(x: torch.Tensor) -> torch.Tensor:
    """uint8 [B,H,W,3] (a decoded RGB batch) -> fp32 [B,3,H,W] in [0,1]: torchvision's ToTensor on the device."""
    _ensure(x)
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
        raise _lib.GhError(f"u8hwc_to_f32chw: expected uint8 [B,H,W,3], got {x.dtype} {tuple(x.shape)}")
    xc = x.contiguous()
    B, H, W, _ = xc.shape
    y = torch.empty(B, 3, H, W, dtype=F32, device=x.device)
    check(_lib.lib().gh_u8hwc_to_f32chw(xc.data_ptr(), y.data_ptr(), B, H, W, _stream()))
    _count()
    return y


def upsample2x_nhwc(x: torch.Tensor) -> torch.Tensor:
    _ensure(x)
    assert x.dtype == BF16 and x.is_contiguous() and x.dim() == 4
    B, H, W, Cc = x.shape
    y = torch.empty(B, 2 * H, 2 * W, Cc, dtype=BF16, device=x.device)
    check(_lib.lib().gh_upsample2x_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, Cc, _stream()))
    _count()
    return y


def softmax_rows(s, n, scale, ld_out):
    _ensure(s)
    assert s.dtype == F32 and s.dim() == 2 and s.stride(1) == 1
    rows = s.shape[0]
    p = torch.empty(rows, ld_out, dtype=BF16, device=s.device)
    check(_lib.lib().gh_softmax_rows(s.data_ptr(), s.stride(0), p.data_ptr(), ld_out, rows, n, float(scale), _stream()))
    _count()
    return p


def ae_sample_patchify(moments_nhwc, noise_nchw, scale_factor, shift_factor):
    _ensure(moments_nhwc)
    assert moments_nhwc.dtype == F32 and noise_nchw.dtype == F32 and moments_nhwc.is_contiguous()
    B, h, w, z2 = moments_nhwc.shape
    z = z2 // 2
    nz = noise_nchw.contiguous()
    out = torch.empty(B, (h // 2) * (w // 2), z * 4, dtype=F32, device=moments_nhwc.device)
    check(_lib.lib().gh_ae_sample_patchify(moments_nhwc.data_ptr(), nz.data_ptr(), out.data_ptr(), B, h, w, z,
                                           float(scale_factor), float(shift_factor), _stream()))
    _count()
    return out


def _attn_out(out1, out0, n_split) -> AttnOut:
    o = AttnOut()
    if out0 is not None and n_split > 0:
        assert out0.stride(2) == 1
        o.seg0, o.seg0_batch_stride, o.seg0_row_stride = out0.data_ptr(), out0.stride(0), out0.stride(1)
    assert out1.stride(2) == 1
    o.seg1, o.seg1_batch_stride, o.seg1_row_stride = out1.data_ptr(), out1.stride(0), out1.stride(1)
    o.n_split = n_split
    return o


def _dvalid(d_valid: int, D: int) -> int:
    return 0 if not d_valid or d_valid >= D else (int(d_valid) + 15) // 16 * 16


def flash_attn_bwd(q, k, v, lse, scale, o1, do1, dq, dk, dv, o0=None, do0=None, n_split=0, d_valid=0):
    """Backward of flash_attn_fwd.  o*/do* token-major (same split as forward); dq/dk/dv [B,H,L,D] views (any
    strides, D contiguous) are written."""
    _ensure(q)
    B, H, Lq, D = q.shape
    Lk = k.shape[2]
    at = [_attn_tensor(x) for x in (q, k, v, dq, dk, dv)]
    o = _attn_out(o1, o0, n_split)
    do = _attn_out(do1, do0, n_split)
    ws_do = torch.empty(_lib.lib().gh_flash_attn_bwd_workspace_bytes(B, H, Lq, D, 0) // 2, dtype=BF16, device=q.device)
    ws_delta = torch.empty(_lib.lib().gh_flash_attn_bwd_workspace_bytes(B, H, Lq, D, 1) // 4, dtype=F32, device=q.device)
    check(_lib.lib().gh_flash_attn_bwd(C.byref(at[0]), C.byref(at[1]), C.byref(at[2]), C.byref(o), C.byref(do),
                                       lse.data_ptr(), B, H, Lq, Lk, D, _dvalid(d_valid, D), float(scale), C.byref(at[3]), C.byref(at[4]),
                                       C.byref(at[5]), ws_do.data_ptr(), ws_delta.data_ptr(), _stream()))
    _count(3)


_SUMSQ_WS: dict = {}


def sumsq_workspace(device) -> torch.Tensor:
    """The zeroed scratch gh_sumsq_accum reduces through (block partials + a ticket counter), one per device; calls
    that share it must be ordered on one stream (the optimizer's)."""
    ws = _SUMSQ_WS.get(device)
    if ws is None:
        ws = _SUMSQ_WS[device] = torch.zeros(_lib.lib().gh_sumsq_workspace_bytes() // 4, dtype=F32, device=device)
    return ws


def sumsq_accum(g: torch.Tensor, acc: torch.Tensor, ws: torch.Tensor | None = None) -> None:
    """acc[0] (fp32) += sum(g^2) over a flat contiguous bf16 / fp32 buffer; bit-reproducible (fixed summation order)."""
    _ensure(g)
    assert g.is_contiguous() and acc.dtype == F32
    if ws is None:
        ws = sumsq_workspace(g.device)
    check(_lib.lib().gh_sumsq_accum(g.data_ptr(), _dt(g), g.numel(), acc.data_ptr(), ws.data_ptr(), _stream()))
    _count()


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, gnorm_sq=None, max_norm=0.0, grad_scale=1.0,
               dev_state=None):
    """In-place AdamW on flat contiguous buffers of one dtype; optional fused clip-by-global-norm.
    ``dev_state`` (int32[2] on the device: step count, enable flag) replaces the by-value ``step`` -- the form a
    CUDA graph can replay (see gh_adamw_step)."""
    _ensure(p)
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    assert p.dtype == g.dtype == m.dtype == v.dtype and p.numel() == g.numel() == m.numel() == v.numel()
    check(_lib.lib().gh_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _dt(p), p.numel(), lr, beta1,
                                   beta2, eps, weight_decay, int(step), _p(gnorm_sq), float(max_norm),
                                   float(grad_scale), _p(dev_state), _stream()))
    _count()


def euler_cfg_step(x: torch.Tensor, pred: torch.Tensor, neg_pred: torch.Tensor | None, dt: float, true_gs: float) -> None:
    """x += dt * (neg + true_gs * (pred - neg) | pred), bf16, in place (one Euler step of the flow sampler)."""
    _ensure(x)
    assert x.dtype == BF16 and pred.dtype == BF16 and x.is_contiguous() and pred.is_contiguous() and x.numel() == pred.numel()
    if neg_pred is not None:
        assert neg_pred.dtype == BF16 and neg_pred.is_contiguous() and neg_pred.numel() == x.numel()
    check(_lib.lib().gh_euler_cfg_step(x.data_ptr(), pred.data_ptr(), _p(neg_pred), float(dt), float(true_gs), x.numel(),
                                       _stream()))
    _count()


def dropout_fwd(x: torch.Tensor, p: float, seed: int, offset: int, offset_base: torch.Tensor | None = None) -> torch.Tensor:
    """bf16 inverted dropout with a counter-based mask (seed, offset [+ *offset_base, an int64 device scalar]): see
    gh_dropout_fwd."""
    _ensure(x)
    assert x.dtype == BF16 and x.is_contiguous()
    y = torch.empty_like(x)
    check(_lib.lib().gh_dropout_fwd(x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed), int(offset),
                                    _p(offset_base), _stream()))
    _count()
    return y


def dropout_bwd_add(t: torch.Tensor, dx: torch.Tensor, p: float, seed: int, offset: int,
                    offset_base: torch.Tensor | None = None) -> None:
    """dx += mask(seed, offset) * t / (1 - p), in place."""
    _ensure(t)
    assert t.dtype == BF16 and dx.dtype == BF16 and t.is_contiguous() and dx.is_contiguous() and t.numel() == dx.numel()
    check(_lib.lib().gh_dropout_bwd_add(t.data_ptr(), dx.data_ptr(), t.numel(), float(p), int(seed), int(offset),
                                        _p(offset_base), _stream()))
    _count()


def lora_dropout_fwd(x: torch.Tensor, a_ext: torch.Tensor, r: int, alpha: float, p: float, seed: int, offset: int,
                     offset_base: torch.Tensor | None = None):
    """-> (xd, u): xd = dropout(x) [M,K] and u [M, a_ext.shape[0]] with u[:, :r] = alpha * xd @ a_ext[:r]^T and, when a_ext
    has 16 more rows than r, u[:, r:] = [1, 0, ...] -- one pass over x (gh_lora_dropout_fwd)."""
    _ensure(x)
    assert x.dtype == BF16 and x.is_contiguous() and x.dim() == 2 and a_ext.dtype == BF16 and a_ext.is_contiguous()
    M, Kd = x.shape
    rw = a_ext.shape[0]
    xd = torch.empty_like(x)
    u = torch.empty(M, rw, dtype=BF16, device=x.device)
    check(_lib.lib().gh_lora_dropout_fwd(x.data_ptr(), xd.data_ptr(), a_ext.data_ptr(), u.data_ptr(), M, Kd, r, rw - r, rw,
                                         float(alpha), float(p), int(seed), int(offset), _p(offset_base), _stream()))
    _count()
    return xd, u


def lora_dropout_bwd(du: torch.Tensor, a: torch.Tensor, dx: torch.Tensor, p: float, seed: int, offset: int,
                     offset_base: torch.Tensor | None = None, act_pre: torch.Tensor | None = None, act: int = ACT_NONE) -> None:
    """dx += mask(seed, offset) * (du @ a) / (1 - p), in place; du [M,R] (unit inner stride), a [R,K] contiguous.
    ``act_pre`` ([M,K] bf16, contiguous): the sum is then multiplied by act'(act_pre) (dgrad through an MLP activation)."""
    _ensure(du)
    assert du.dtype == BF16 and a.dtype == BF16 and dx.dtype == BF16 and a.is_contiguous() and dx.is_contiguous()
    assert du.dim() == 2 and du.stride(1) == 1 and dx.dim() == 2 and dx.shape == (du.shape[0], a.shape[1]) and du.shape[1] == a.shape[0]
    assert act_pre is None or (act_pre.dtype == BF16 and act_pre.is_contiguous() and act_pre.shape == dx.shape)
    check(_lib.lib().gh_lora_dropout_bwd(du.data_ptr(), a.data_ptr(), dx.data_ptr(), dx.shape[0], dx.shape[1], a.shape[0],
                                         du.stride(0), float(p), int(seed), int(offset), _p(offset_base), _p(act_pre),
                                         int(act), _stream()))
    _count()


class CopyTable:
    """A device-resident table of ``gh_copy_desc`` built once (pointers are stable: parameters and gradients live in
    the flat buffers of ``optim.flatten``); ``run()`` is ONE launch for all of them."""

    def __init__(self, device):
        self.device = device
        self._descs: list = []
        self._dev = None
        self._keep: list = []

    def add(self, src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, scale: float = 1.0,
            accumulate: bool = False, src_rows=None, src_cols=None, dst_rows=None, dst_cols=None) -> None:
        """Logical [rows, cols] matrix from ``src`` to ``dst`` (2-D views, unit inner stride; the storage of a side may
        be wider / taller than the logical matrix when that side is grouped: ``*_rows`` / ``*_cols`` = (group, pitch))."""
        from ._lib import CopyDesc
        for t in (src, dst):
            if t.dim() != 2 or (t.stride(1) != 1 and t.shape[1] != 1):
                raise _lib.GhError("CopyTable: 2-D views with unit inner stride expected")
        d = CopyDesc()
        d.src, d.dst, d.rows, d.cols = src.data_ptr(), dst.data_ptr(), rows, cols
        d.src_ld, d.dst_ld = src.stride(0), dst.stride(0)
        d.src_dtype, d.dst_dtype = _dt(src), _dt(dst)
        d.src_row_group, d.src_row_pitch = src_rows or (0, 0)
        d.src_col_group, d.src_col_pitch = src_cols or (0, 0)
        d.dst_row_group, d.dst_row_pitch = dst_rows or (0, 0)
        d.dst_col_group, d.dst_col_pitch = dst_cols or (0, 0)
        d.scale, d.accumulate = scale, int(accumulate)
        self._descs.append(d)
        self._keep += [src, dst]
        self._dev = None

    def __len__(self):
        return len(self._descs)

    def run(self, blocks_per_desc: int = 4) -> None:
        if not self._descs:
            return
        from ._lib import CopyDesc
        if self._dev is None:
            arr = (CopyDesc * len(self._descs))(*self._descs)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._dev = host.to(self.device)
        _lib.init(self._dev.device.index if self._dev.device.index is not None else torch.cuda.current_device())
        for i in range(0, len(self._descs), 65535):
            n = min(65535, len(self._descs) - i)
            check(_lib.lib().gh_batched_copy(self._dev.data_ptr() + i * C.sizeof(CopyDesc), n, blocks_per_desc, _stream()))
            _count()
