"""Bring-up self-test for the sm_100a kernels (run on the GPU box through gpurun).

Prints one line per case so a failure localises itself; writes the same to
gpurun_out/selftest.log.  Not a pytest file: tests/ holds the graded parity tests.
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from genhancer_b200 import kernels as K

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/selftest.log", "a")


LINES: list[str] = []  # every line said, so tests/test_kernels_gpu.py can assert that none starts with FAIL


def say(*a):
    s = " ".join(str(x) for x in a)
    LINES.append(s)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def relerr(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def gemm_case(M, N, Kd, a_mn=False, b_mn=False, pad=0, seed=0, **epi):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    A = torch.randn(M, Kd, device=dev, generator=g).to(torch.bfloat16)
    B = torch.randn(N, Kd, device=dev, generator=g).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    if pad:
        def padded(t):
            buf = torch.zeros(t.shape[0], t.shape[1] + pad, device=dev, dtype=t.dtype)
            buf[:, : t.shape[1]] = t
            return buf[:, : t.shape[1]]
        a_in, b_in = padded(a_in), padded(b_in)
    kw = {}
    if epi.get("bias"):
        bias = torch.randn(N, device=dev, generator=g).to(torch.bfloat16)
        kw["bias"] = bias
        ref = ref + bias.float()
    if epi.get("aux_out"):
        kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        aux_ref = ref.clone()
    act = epi.get("act", 0)
    if epi.get("act_grad"):
        aux_in = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
        kw["aux_in"] = aux_in
        kw["act_grad"] = True
        x = aux_in.float().requires_grad_(True)
        y = {1: lambda v: torch.nn.functional.gelu(v, approximate="tanh"), 2: lambda v: v * torch.sigmoid(1.702 * v),
             3: torch.nn.functional.gelu, 4: torch.nn.functional.silu}[act](x)
        (dact,) = torch.autograd.grad(y.sum(), x)
        ref = ref * dact
    elif act:
        ref = {1: lambda v: torch.nn.functional.gelu(v, approximate="tanh"), 2: lambda v: v * torch.sigmoid(1.702 * v),
               3: torch.nn.functional.gelu, 4: torch.nn.functional.silu}[act](ref)
    if epi.get("gate"):
        rpb = epi["gate"]
        nb = (M + rpb - 1) // rpb
        gate = torch.randn(nb, N, device=dev, generator=g).to(torch.bfloat16)
        kw["gate"] = gate
        kw["rows_per_batch"] = rpb
        ref = ref * gate.float().repeat_interleave(rpb, dim=0)[:M]
    if epi.get("residual"):
        res = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
        kw["residual"] = res
        ref = ref + res.float()
    out_dtype = torch.float32 if epi.get("f32") else torch.bfloat16
    out = K.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, act=act, out_dtype=out_dtype, **kw)
    torch.cuda.synchronize()
    e = relerr(out, ref)
    tag = f"gemm M={M} N={N} K={Kd} a_mn={int(a_mn)} b_mn={int(b_mn)} pad={pad} epi={epi}"
    ok = e < 1e-2
    if epi.get("aux_out"):
        e2 = relerr(kw["aux_out"], aux_ref)
        ok = ok and e2 < 1e-2
        tag += f" aux_err={e2:.2e}"
    say(("PASS" if ok else "FAIL"), tag, f"relerr={e:.3e}")
    return ok


def lora_gemm_case(M, N, Kd, R, a_mn=False, b_mn=False, seed=0, f32=False, **kw):
    """D = A B^T + A2 B2^T (second operand pair of reduction length R: the LoRA branch folded into the base GEMM)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda r, c: torch.randn(r, c, device="cuda", generator=g).to(torch.bfloat16)
    A, B, A2, B2 = mk(M, Kd), mk(N, Kd), mk(M, R), mk(N, R)
    ref = A.float() @ B.float().t() + A2.float() @ B2.float().t()
    ta = (lambda t: t.t().contiguous()) if a_mn else (lambda t: t)
    tb = (lambda t: t.t().contiguous()) if b_mn else (lambda t: t)
    if kw.get("residual"):
        res = mk(M, N)
        ref = ref + res.float()
        kw["residual"] = res
    out = K.gemm(ta(A), tb(B), a_mn=a_mn, b_mn=b_mn, a2=ta(A2), b2=tb(B2), out_dtype=torch.float32 if f32 else torch.bfloat16,
                 **kw)
    torch.cuda.synchronize()
    e = relerr(out, ref)
    say("PASS" if e < 1e-2 else "FAIL", f"lora_gemm M={M} N={N} K={Kd} R={R} a_mn={int(a_mn)} b_mn={int(b_mn)} f32={f32}",
        f"relerr={e:.3e}")


def copy_table_cases():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    tab = K.CopyTable(dev)
    # fp32 [6*72, 16] -> bf16 head slots of 128 rows inside a wider block-diagonal operand
    src = torch.randn(6 * 72, 16, device=dev, generator=g)
    dst = torch.zeros(6 * 128, 48, device=dev, dtype=torch.bfloat16)
    tab.add(src, dst[:, 16:32], 6 * 72, 16, dst_rows=(72, 128), scale=0.5)
    # fp32 [16, 6*72] -> bf16 column slots
    src2 = torch.randn(16, 6 * 72, device=dev, generator=g)
    dst2 = torch.zeros(16, 6 * 128, device=dev, dtype=torch.bfloat16)
    tab.add(src2, dst2, 16, 6 * 72, dst_cols=(72, 128))
    # gradient scatter: fp32 slots -> fp32 dense, accumulating
    src3 = torch.randn(6 * 128, device=dev, generator=g)
    dst3 = torch.ones(6 * 72, device=dev)
    tab.add(src3.unsqueeze(1), dst3.unsqueeze(1), 6 * 72, 1, accumulate=True, src_rows=(72, 128))
    tab.run()
    torch.cuda.synchronize()
    r1 = torch.zeros_like(dst)
    r1.view(6, 128, 48)[:, :72, 16:32] = (0.5 * src).view(6, 72, 16).to(torch.bfloat16)
    r2 = torch.zeros_like(dst2)
    r2.view(16, 6, 128)[:, :, :72] = src2.view(16, 6, 72).to(torch.bfloat16)
    r3 = 1.0 + src3.view(6, 128)[:, :72].reshape(-1)
    ok = torch.equal(dst, r1) and torch.equal(dst2, r2) and torch.equal(dst3, r3)
    say("PASS" if ok else "FAIL", "batched_copy: row slots / column slots / accumulate (bit-exact)")


def time_gemm(M, N, Kd, a_mn=False, b_mn=False, iters=20):
    dev = "cuda"
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        K.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * Kd / ms / 1e9
    # library yardstick (cuBLAS through torch) on the same shape, for context only
    Ar = A.t() if a_mn else A
    Br = B.t() if b_mn else B
    for _ in range(3):
        torch.matmul(Ar, Br.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(Ar, Br.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    say(f"TIME gemm M={M} N={N} K={Kd} a_mn={int(a_mn)} b_mn={int(b_mn)}: {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS {ms2:.3f} ms "
        f"{2.0 * M * N * Kd / ms2 / 1e9:.0f} TFLOP/s")


def batched_cases():
    """gh_gemm_bf16 batch mode (the AE mid-block attention, autoencoder.py:37-52) against torch.bmm, and the split-K
    dgrad of a 32-row Modulation (layers.py:169-175) against a plain matmul."""
    g = torch.Generator(device="cuda").manual_seed(11)
    for (Bn, L, C) in [(3, 1764, 512), (5, 200, 64), (32, 444, 128), (2, 128, 256)]:
        qkv = (torch.randn(Bn * L, 3 * C, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        s = K.gemm(q, k, out_dtype=torch.float32, batch=Bn)                       # [Bn*L, L]
        ref_s = torch.bmm(q.float().view(Bn, L, C), k.float().view(Bn, L, C).transpose(1, 2)).view(Bn * L, L)
        e = relerr(s, ref_s)
        say("PASS" if e < 2e-3 else "FAIL", f"batched QK^T B={Bn} L={L} C={C}", f"relerr={e:.3e}")
        ldp = (L + 7) // 8 * 8
        p = K.softmax_rows(s, L, C ** -0.5, ldp)
        o = K.gemm(p[:, :L], v, b_mn=True, batch=Bn)                               # [Bn*L, C]
        ref_o = torch.bmm(p[:, :L].float().view(Bn, L, L), v.float().view(Bn, L, C)).view(Bn * L, C)
        e = relerr(o, ref_o)
        say("PASS" if e < 1e-2 else "FAIL", f"batched PV   B={Bn} L={L} C={C}", f"relerr={e:.3e}")
        # bias + residual through the batched epilogue rows
        bias = torch.randn(C, device="cuda", generator=g).to(torch.bfloat16)
        res = torch.randn(Bn * L, C, device="cuda", generator=g).to(torch.bfloat16)
        o2 = K.gemm(p[:, :L], v, b_mn=True, batch=Bn, bias=bias, residual=res)
        e = relerr(o2, ref_o + bias.float() + res.float())
        say("PASS" if e < 1e-2 else "FAIL", f"batched PV + bias + residual B={Bn} L={L} C={C}", f"relerr={e:.3e}")
    for (M, N, Kd) in [(32, 3072, 18432), (32, 3072, 9216), (32, 3072, 6144), (2, 768, 1536)]:
        A = torch.randn(M, Kd, device="cuda", generator=g).to(torch.bfloat16)
        W = (torch.randn(Kd, N, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        base = torch.randn(M, N, device="cuda", generator=g)
        out = base.clone()
        K.gemm(A, W, b_mn=True, out=out, k_splits=-1)
        ref = base + A.float() @ W.float()
        e = relerr(out, ref)
        say("PASS" if e < 2e-3 else "FAIL", f"split-K dgrad M={M} N={N} K={Kd}", f"relerr={e:.3e}")


def fm_cases():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    B, L = 4, 441
    x1 = torch.randn(B, L, 64, device=dev, generator=g)
    x0 = torch.randn(B, L, 64, device=dev, generator=g)
    t = torch.sigmoid(torch.randn(B, device=dev, generator=g))
    xt = K.fm_interp(x1, x0, t)
    ref = ((1 - t[:, None, None]) * x1 + t[:, None, None] * x0).to(torch.bfloat16)
    say("PASS" if torch.equal(xt, ref) else "FAIL", "fm_interp bit-exact:", torch.equal(xt, ref),
        "maxdiff", (xt.float() - ref.float()).abs().max().item())
    pred = torch.randn(B, L, 64, device=dev, generator=g).to(torch.bfloat16)
    loss, dpred = K.fm_mse_loss(pred, x0, x1)
    pr = pred.float().requires_grad_(True)
    lref = torch.nn.functional.mse_loss(pr, (x0 - x1))
    lref.backward()
    e1 = abs(loss.item() - lref.item()) / abs(lref.item())
    e2 = relerr(dpred, pr.grad)
    say("PASS" if (e1 < 1e-5 and e2 < 5e-3) else "FAIL", f"fm_mse loss relerr={e1:.2e} dpred relerr={e2:.2e}")


def main():
    say("== genhancer_b200 selftest", time.strftime("%H:%M:%S"), torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["fm", "gemm", "mn", "epi", "time"]
    if "fm" in which:
        fm_cases()
    if "gemm" in which:
        gemm_case(128, 64, 64)
        gemm_case(128, 256, 64)
        gemm_case(128, 256, 256)
        gemm_case(256, 512, 512)
        gemm_case(200, 328, 588, pad=4)
        gemm_case(32, 3072, 768)
        gemm_case(1000, 64, 3072)
        gemm_case(3000, 3072, 1024)
    if "mn" in which:
        ok_b = gemm_case(128, 256, 256, b_mn=True)
        ok_a = gemm_case(128, 256, 256, a_mn=True)
        gemm_case(128, 64, 128, a_mn=True, b_mn=True)
        gemm_case(304, 520, 328, a_mn=True, b_mn=True)
        gemm_case(1000, 3072, 777 + 7, b_mn=True)
        if not (ok_a and ok_b):
            for cand in ["1024,8192,2048", "8192,1024,4096", "128,1024,2048", "1024,128,2048"]:
                os.environ["GH_DEBUG_MN_DESC"] = cand
                say("-- retry MN descriptors with", cand)
                gemm_case(128, 256, 256, b_mn=True)
                gemm_case(128, 256, 256, a_mn=True)
            del os.environ["GH_DEBUG_MN_DESC"]
    if "epi" in which:
        gemm_case(300, 512, 256, bias=True)
        gemm_case(300, 512, 256, bias=True, act=1)
        gemm_case(300, 512, 256, bias=True, act=2)
        gemm_case(300, 512, 256, bias=True, act=3)
        gemm_case(300, 512, 256, bias=True, act=4)
        gemm_case(300, 512, 256, bias=True, act=1, aux_out=True)
        gemm_case(300, 512, 256, act=1, act_grad=True)
        gemm_case(300, 512, 256, act=2, act_grad=True)             # lean TMA epilogue with the aux_in tile
        gemm_case(1000, 768, 320, bias=True, act=1, act_grad=True)
        gemm_case(300, 512, 256, act=4, act_grad=True)             # SiLU': general path
        gemm_case(300, 512, 256, bias=True, gate=100, residual=True)
        gemm_case(300, 512, 256, bias=True, f32=True)
    if "lora" in which:
        lora_gemm_case(300, 512, 256, 16)
        lora_gemm_case(300, 512, 256, 48, residual=True)
        lora_gemm_case(1000, 3456 // 8 * 8, 1152, 48)
        lora_gemm_case(392, 1152, 4304, 16, residual=True)
        lora_gemm_case(300, 520, 256, 16, b_mn=True)            # dgrad: dx = dy W + du A
        lora_gemm_case(300, 1152, 432, 48, b_mn=True)
        lora_gemm_case(300, 512, 256, 80, b_mn=True)            # rank spilling into a second 64-wide k block
        lora_gemm_case(16, 1152, 392, 16, a_mn=True, b_mn=True, f32=True)   # wgrad shapes dA = du^T x (no second pair needed,
        gemm_case(16, 1152, 392, a_mn=True, b_mn=True, f32=True)            #  but the skinny M = 16 / N = 16 tiles are new)
        gemm_case(432, 16, 392, a_mn=True, b_mn=True, f32=True)
        gemm_case(392, 16, 1152)
        gemm_case(392, 48, 432, b_mn=True)
        # split-K: D (fp32, pre-filled) += A^T B over k slices (LoRA wgrads)
        for (M, N, Kd) in [(16, 1152, 23328), (48, 1152, 5000), (4304, 16, 23328), (6144, 48, 3000), (304, 520, 777)]:
            g = torch.Generator(device="cuda").manual_seed(4)
            A = torch.randn(Kd, M, device="cuda", generator=g).to(torch.bfloat16)
            B = torch.randn(Kd, N, device="cuda", generator=g).to(torch.bfloat16)
            base = torch.randn(M, N, device="cuda", generator=g)
            out = base.clone()
            K.gemm(A, B, a_mn=True, b_mn=True, out=out, k_splits=-1)
            torch.cuda.synchronize()
            ref = base + A.float().t() @ B.float()
            e = relerr(out, ref)
            say("PASS" if e < 2e-3 else "FAIL", f"split-K wgrad M={M} N={N} K={Kd}", f"relerr={e:.3e}")
        copy_table_cases()
    if "batched" in which:
        batched_cases()
    if "dyn" in which:
        # dynamic tile schedule (gh_gemm_args::dynamic_tiles): same results, counters hand themselves back zeroed (repeat)
        K.DYNAMIC_TILES = True
        try:
            for rep in range(3):
                gemm_case(3000, 3072, 1024)                          # CTA pairs, several tiles per pair
                gemm_case(14112, 3072, 512, bias=True, act=1)        # > 1 wave of 256-row pair tiles
                gemm_case(200, 328, 588, pad=4)                      # single CTAs, fewer tiles than SMs
                gemm_case(1000, 3072, 784, b_mn=True)
                gemm_case(304, 520, 328, a_mn=True, b_mn=True)
                gemm_case(300, 512, 256, bias=True, gate=100, residual=True)
            batched_cases()
        finally:
            K.DYNAMIC_TILES = False
        gemm_case(3000, 3072, 1024)
    if "time" in which:
        time_gemm(4096, 4096, 4096)
        time_gemm(8192, 8192, 8192)
        time_gemm(14112, 9216, 3072)
        time_gemm(14112, 3072, 12288)
        time_gemm(14112, 12288, 3072)
        time_gemm(14112, 3072, 9216, b_mn=True)
        time_gemm(9216, 3072, 14112, a_mn=True, b_mn=True)
        time_gemm(18464, 3072, 1024)
        time_gemm(32, 18432, 3072)
    say("== done")


if __name__ == "__main__":
    main()
