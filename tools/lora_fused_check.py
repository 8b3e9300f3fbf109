"""Fused LoRA-dropout kernels (csrc/lora_fused.cu) against the unfused pair they replace: same mask, same numbers; timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import kernels as K

dev, BF = "cuda", torch.bfloat16


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def case(M, Kd, R, RX=16, p=0.1):
    g = torch.Generator(device=dev).manual_seed(M + Kd + R)
    x = torch.randn(M, Kd, device=dev, generator=g).to(BF)
    a_ext = torch.zeros(R + RX, Kd, device=dev, dtype=BF)
    a_ext[:R] = (torch.randn(R, Kd, device=dev, generator=g) * Kd ** -0.5).to(BF)
    ub = torch.zeros(R + RX, device=dev)
    if RX:
        ub[R] = 1.0
    base = torch.full((1,), 12345, dtype=torch.int64, device=dev)
    drop = (p, 0x1234567, 7, base)
    xd_ref = K.dropout_fwd(x, *drop)
    u_ref = K.gemm(xd_ref, a_ext, alpha=2.0, bias=ub if RX else None)
    xd, u = K.lora_dropout_fwd(x, a_ext, R, 2.0, *drop)
    e_xd, e_u = (xd != xd_ref).sum().item(), rel(u[:, :R], u_ref[:, :R])
    ones_ok = (not RX) or (bool((u[:, R] == 1).all()) and float(u[:, R + 1:].abs().max()) == 0.0)
    du = torch.randn(M, R + RX, device=dev, generator=g).to(BF)[:, :R]
    dx0 = torch.randn(M, Kd, device=dev, generator=g).to(BF)
    dx_ref = dx0.clone()
    K.dropout_bwd_add(K.gemm(du, a_ext[:R], b_mn=True), dx_ref, *drop)
    dx = dx0.clone()
    K.lora_dropout_bwd(du, a_ext[:R], dx, *drop)
    e_dx = rel(dx, dx_ref)     # (the unfused pair rounds du A to bf16 before the add: agreement to bf16 precision)
    # the same with the MLP's act' applied in the pass (fc2's input gradient)
    pre = torch.randn(M, Kd, device=dev, generator=g).to(BF)
    dxa_ref = dx0.clone()
    K.dropout_bwd_add(K.gemm(du, a_ext[:R], b_mn=True), dxa_ref, *drop)
    dxa_ref = K.act_bwd(dxa_ref, pre, K.ACT_GELU_TANH)
    dxa = dx0.clone()
    K.lora_dropout_bwd(du, a_ext[:R], dxa, *drop, act_pre=pre, act=K.ACT_GELU_TANH)
    e_dxa = rel(dxa, dxa_ref)
    t_f_ref = timeit(lambda: K.gemm(K.dropout_fwd(x, *drop), a_ext, alpha=2.0, bias=ub if RX else None))
    t_f = timeit(lambda: K.lora_dropout_fwd(x, a_ext, R, 2.0, *drop))
    t_b_ref = timeit(lambda: K.dropout_bwd_add(K.gemm(du, a_ext[:R], b_mn=True), dx_ref, *drop))
    t_b = timeit(lambda: K.lora_dropout_bwd(du, a_ext[:R], dx, *drop))
    ok = e_xd == 0 and e_u < 5e-3 and ones_ok and e_dx < 4e-3 and e_dxa < 6e-3
    gb = M * Kd * 2 / 1e9
    print(f"{'PASS' if ok else 'FAIL'} M={M} K={Kd} R={R}+{RX}: xd mismatches {e_xd}, u relerr {e_u:.2e}, ones column {ones_ok}, "
          f"dx relerr {e_dx:.2e} (with act' {e_dxa:.2e}) | fwd {t_f:.0f} us ({2 * gb / t_f * 1e3:.2f} TB/s) vs unfused {t_f_ref:.0f}; "
          f"bwd {t_b:.0f} us ({2 * gb / t_b * 1e3:.2f} TB/s) vs unfused {t_b_ref:.0f}")
    return ok


if __name__ == "__main__":
    ok = True
    for args in [(23328, 1152, 16), (23328, 1152, 48), (23328, 2048, 16), (23328, 4304, 16), (1000, 136, 16), (33, 72, 32, 0), (32, 1152, 16)]:
        ok &= case(*args)
    print("== done", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
