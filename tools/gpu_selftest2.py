"""Bring-up self-test #2: norm / rope / attention kernels vs plain torch (run through gpurun)."""
from __future__ import annotations

import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from genhancer_b200 import kernels as K
from oracle import genhancer_oracle as O

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/selftest2.log", "a")
BF = torch.bfloat16
dev = "cuda"


LINES: list[str] = []  # every line said, so tests/test_kernels_gpu.py can assert that none starts with FAIL


def say(*a):
    s = " ".join(str(x) for x in a)
    LINES.append(s)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def rel(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def check(name, got, ref, tol):
    e = rel(got, ref)
    say("PASS" if e < tol else "FAIL", name, f"relerr={e:.3e}")


def test_ln():
    g = torch.Generator(device=dev).manual_seed(0)
    for (B, L, C) in [(3, 50, 3072), (2, 17, 1024), (2, 9, 768), (1, 5, 4096), (2, 33, 1152)]:
        x = (torch.randn(B, L, C, device=dev, generator=g) * 2 + 0.5).to(BF)
        mod = torch.randn(B, 6 * C, device=dev, generator=g).to(BF) * 0.5
        shift, scale = mod[:, :C], mod[:, C:2 * C]
        y, mean, rstd = K.layernorm_fwd(x, shift=shift, scale=scale, eps=1e-6)
        xr = x.float().requires_grad_(True)
        sh = shift.float().requires_grad_(True)
        sc = scale.float().requires_grad_(True)
        ref = (1 + sc[:, None]) * F.layer_norm(xr, (C,), eps=1e-6) + sh[:, None]
        check(f"ln_adaln fwd {B}x{L}x{C}", y, ref, 6e-3)
        dy = torch.randn(B, L, C, device=dev, generator=g).to(BF)
        dres = torch.randn(B, L, C, device=dev, generator=g).to(BF)
        ref.backward(dy.float())
        dx = K.layernorm_bwd_dx(dy, x, mean, rstd, scale=scale, dres=dres)
        check(f"ln_adaln bwd dx(+dres) {C}", dx, xr.grad + dres.float(), 8e-3)
        acc = torch.zeros(B, 2 * C, device=dev)
        K.layernorm_bwd_params(dy, x, mean, rstd, acc[:, :C], acc[:, C:])
        check(f"ln_adaln dshift {C}", acc[:, :C], sh.grad, 5e-3)
        check(f"ln_adaln dscale {C}", acc[:, C:], sc.grad, 5e-3)
        # affine
        w = torch.randn(C, device=dev, generator=g) * 0.2 + 1
        bb = torch.randn(C, device=dev, generator=g) * 0.2
        y2, m2, r2 = K.layernorm_fwd(x, weight=w, bias=bb, eps=1e-5)
        xr2 = x.float().requires_grad_(True)
        wr, br = w.clone().requires_grad_(True), bb.clone().requires_grad_(True)
        ref2 = F.layer_norm(xr2, (C,), wr, br, 1e-5)
        check(f"ln_affine fwd {C}", y2, ref2, 6e-3)
        ref2.backward(dy.float())
        dx2 = K.layernorm_bwd_dx(dy, x, m2, r2, weight=w)
        check(f"ln_affine bwd dx {C}", dx2, xr2.grad, 8e-3)
        acc2 = torch.zeros(2, C, device=dev)
        K.layernorm_bwd_params(dy.reshape(-1, C), x.reshape(-1, C), m2, r2, acc2[0], acc2[1])
        check(f"ln_affine dbias {C}", acc2[0], br.grad, 5e-3)
        check(f"ln_affine dweight {C}", acc2[1], wr.grad, 5e-3)
    # sliced input (drop txt tokens): x[:, 3:, :]
    B, L, C = 2, 20, 3072
    xf = torch.randn(B, L, C, device=dev, generator=g).to(BF)
    mod = torch.randn(B, 2 * C, device=dev, generator=g).to(BF)
    y, _, _ = K.layernorm_fwd(xf[:, 3:], shift=mod[:, :C], scale=mod[:, C:])
    ref = (1 + mod[:, None, C:].float()) * F.layer_norm(xf[:, 3:].float(), (C,), eps=1e-6) + mod[:, None, :C].float()
    check("ln sliced view", y, ref, 6e-3)


def test_gate_colsum():
    g = torch.Generator(device=dev).manual_seed(1)
    B, L, C = 3, 77, 3072
    dout = torch.randn(B, L, C, device=dev, generator=g).to(BF)
    u = torch.randn(B, L, C, device=dev, generator=g).to(BF)
    gate = torch.randn(B, 3 * C, device=dev, generator=g).to(BF)[:, C:2 * C]
    acc = torch.zeros(B, C, device=dev)
    dbias = torch.zeros(acc.shape[-1], dtype=torch.float32, device=dout.device)
    du = K.gate_bwd(dout, u, gate, acc, dbias_acc=dbias)
    check("gate_bwd du", du, gate.float()[:, None] * dout.float(), 5e-3)
    check("gate_bwd dbias", dbias, (gate.float()[:, None] * dout.float()).sum((0, 1)), 5e-3)
    check("gate_bwd dgate", acc, (dout.float() * u.float()).sum(1), 5e-3)
    accb = torch.zeros(C, device=dev)
    K.colsum(dout, accb)
    check("colsum", accb, dout.float().sum((0, 1)), 5e-3)


def test_rope_qknorm():
    g = torch.Generator(device=dev).manual_seed(2)
    B, H, D = 2, 3, 128
    n_txt, n_img = 5, 3 * 4
    Ltot = n_txt + n_img
    img_ids = O.make_img_ids(B, 3, 4).to(dev)
    txt_ids = torch.cat([O.create_spatio_temporal_ids(B, 2, 1, 5)], 1).to(dev)
    ids = torch.cat([txt_ids, img_ids], 1)
    cs = K.rope_table(ids)
    pe = O.rope_table(ids, (16, 56, 56), 10000)  # [B,1,L,64,2,2]
    check("rope_table cos", cs[..., 0], pe[:, 0, :, :, 0, 0], 1e-6)
    check("rope_table sin", cs[..., 1], pe[:, 0, :, :, 1, 0], 1e-6)
    qs = (1 + 0.1 * torch.randn(D, device=dev, generator=g)).to(BF)
    ks = (1 + 0.1 * torch.randn(D, device=dev, generator=g)).to(BF)
    q = torch.zeros(B, H, Ltot, D, device=dev, dtype=BF)
    k = torch.zeros_like(q)
    v = torch.zeros_like(q)
    qkv_t = torch.randn(B, n_txt, 3 * H * D, device=dev, generator=g).to(BF)
    qkv_i = torch.randn(B, n_img, 3 * H * D, device=dev, generator=g).to(BF)
    K.qk_norm_rope_fwd(qkv_t, H, qs, ks, cs, q, k, v, 0)
    K.qk_norm_rope_fwd(qkv_i, H, qs, ks, cs, q, k, v, n_txt)
    # reference (bf16 semantics of the reference code)
    qkv = torch.cat([qkv_t, qkv_i], 1).float().requires_grad_(True)
    qsf, ksf = qs.float().requires_grad_(True), ks.float().requires_grad_(True)
    rq, rk, rv = O._heads(qkv, H)
    rq, rk = O._rms(rq, qsf), O._rms(rk, ksf)
    rq, rk = O.apply_rope(rq, rk, pe)
    check("qk_norm_rope q", q, rq, 6e-3)
    check("qk_norm_rope k", k, rk, 6e-3)
    check("qk_norm_rope v", v, rv, 1e-6)
    dq = torch.randn(B, H, Ltot, D, device=dev, generator=g).to(BF)
    dk = torch.randn(B, H, Ltot, D, device=dev, generator=g).to(BF)
    dv = torch.randn(B, H, Ltot, D, device=dev, generator=g).to(BF)
    (rq * dq.float()).sum().add((rk * dk.float()).sum()).add((rv * dv.float()).sum()).backward()
    dqkv_t = torch.empty_like(qkv_t)
    dqkv_i = torch.empty_like(qkv_i)
    accq, acck = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    K.qk_norm_rope_bwd(dq, dk, dv, qkv_t, H, qs, ks, cs, 0, dqkv_t, accq, acck)
    K.qk_norm_rope_bwd(dq, dk, dv, qkv_i, H, qs, ks, cs, n_txt, dqkv_i, accq, acck)
    check("qk_norm_rope bwd dqkv", torch.cat([dqkv_t, dqkv_i], 1), qkv.grad, 8e-3)
    check("qk_norm_rope bwd dscale_q", accq, qsf.grad, 8e-3)
    check("qk_norm_rope bwd dscale_k", acck, ksf.grad, 8e-3)


def test_misc():
    g = torch.Generator(device=dev).manual_seed(3)
    t = torch.sigmoid(torch.randn(7, device=dev, generator=g))
    e = K.timestep_embedding(t, round_bf16=True)
    ref = O.timestep_embedding(t.to(BF), 256)
    check("timestep_embedding (bf16 t)", e, ref, 4e-3)
    e2 = K.timestep_embedding(torch.full((3,), 4.0, device=dev), round_bf16=True)
    check("timestep_embedding guidance", e2, O.timestep_embedding(torch.full((3,), 4.0, device=dev, dtype=BF), 256), 4e-3)
    x = torch.randn(5, 3072, device=dev, generator=g).to(BF)
    check("act silu", K.act_fwd(x, K.ACT_SILU), F.silu(x.float()), 4e-3)
    dy = torch.randn(5, 3072, device=dev, generator=g).to(BF)
    xr = x.float().requires_grad_(True)
    F.silu(xr).backward(dy.float())
    check("act silu bwd", K.act_bwd(dy, x, K.ACT_SILU), xr.grad, 5e-3)
    src = torch.randn(1000, device=dev, generator=g)
    dst = torch.ones(1000, device=dev, dtype=BF)
    K.accum_cast(src, dst, 0.5, True)
    check("accum_cast", dst, 1 + 0.5 * src, 5e-3)


def attn_case(B, H, L, D, n_split=0, fused_qkv=False, seed=5):
    g = torch.Generator(device=dev).manual_seed(seed)
    if fused_qkv:  # ViT layout: [B, L, 3, H, D] fused projection output
        qkv = torch.randn(B, L, 3, H, D, device=dev, generator=g).to(BF)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    else:
        q, k, v = (torch.randn(B, H, L, D, device=dev, generator=g).to(BF) for _ in range(3))
    scale = D ** -0.5
    out1 = torch.full((B, L - n_split, H * D), float("nan"), device=dev, dtype=BF)
    out0 = torch.full((B, max(n_split, 1), H * D), float("nan"), device=dev, dtype=BF) if n_split else None
    lse = K.flash_attn_fwd(q, k, v, scale, out1, out0, n_split)
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float())
    ref = ref.transpose(1, 2).reshape(B, L, H * D)
    got = torch.cat([out0, out1], 1) if n_split else out1
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    lref = torch.logsumexp(s, -1) / math.log(2.0)
    e, e2 = rel(got, ref), rel(lse, lref)
    say("PASS" if (e < 8e-3 and e2 < 1e-3) else "FAIL", f"flash_fwd B={B} H={H} L={L} D={D} split={n_split} fused={fused_qkv}",
        f"relerr={e:.3e} lse relerr={e2:.3e} nan={torch.isnan(got.float()).any().item()}")


def attn_bwd_case(B, H, L, D, n_split=0, fused_qkv=False, seed=7):
    g = torch.Generator(device=dev).manual_seed(seed)
    if fused_qkv:
        qkv = torch.randn(B, L, 3, H, D, device=dev, generator=g).to(BF)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        dqkv = torch.full_like(qkv, float("nan"))
        dq, dk, dv = (dqkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    else:
        q, k, v = (torch.randn(B, H, L, D, device=dev, generator=g).to(BF) for _ in range(3))
        dq, dk, dv = (torch.full((B, H, L, D), float("nan"), device=dev, dtype=BF) for _ in range(3))
    scale = D ** -0.5
    out1 = torch.empty(B, L - n_split, H * D, device=dev, dtype=BF)
    out0 = torch.empty(B, max(n_split, 1), H * D, device=dev, dtype=BF) if n_split else None
    lse = K.flash_attn_fwd(q, k, v, scale, out1, out0, n_split)
    do_full = torch.randn(B, L, H * D, device=dev, generator=g).to(BF)
    do0 = do_full[:, :n_split].contiguous() if n_split else None
    do1 = do_full[:, n_split:].contiguous()
    K.flash_attn_bwd(q, k, v, lse, scale, out1, do1, dq, dk, dv, out0, do0, n_split)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, L, H * D)
    ref.backward(do_full.float())
    e = [rel(dq, qf.grad), rel(dk, kf.grad), rel(dv, vf.grad)]
    say("PASS" if max(e) < 1.2e-2 else "FAIL", f"flash_bwd B={B} H={H} L={L} D={D} split={n_split} fused={fused_qkv}",
        f"dq={e[0]:.3e} dk={e[1]:.3e} dv={e[2]:.3e}")


def attn_cross_bwd_case(B, H, Lq, Lk, D, seed=9):
    """Lq != Lk (SigLIP MAP head: one probe query over all tokens), k/v read off a fused [B, Lk, 2, H, D] buffer."""
    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.randn(B, H, Lq, D, device=dev, generator=g).to(BF)
    kv = torch.randn(B, Lk, 2, H, D, device=dev, generator=g).to(BF)
    k, v = (kv[:, :, i].permute(0, 2, 1, 3) for i in range(2))
    dkv = torch.full_like(kv, float("nan"))
    dk, dv = (dkv[:, :, i].permute(0, 2, 1, 3) for i in range(2))
    dq = torch.full((B, H, Lq, D), float("nan"), device=dev, dtype=BF)
    scale = D ** -0.5
    out = torch.empty(B, Lq, H * D, device=dev, dtype=BF)
    lse = K.flash_attn_fwd(q, k, v, scale, out)
    do = torch.randn(B, Lq, H * D, device=dev, generator=g).to(BF)
    K.flash_attn_bwd(q, k, v, lse, scale, out, do, dq, dk, dv)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Lq, H * D)
    ref.backward(do.float())
    e = [rel(out, ref), rel(dq, qf.grad), rel(dk, kf.grad), rel(dv, vf.grad)]
    say("PASS" if max(e) < 1.2e-2 else "FAIL", f"flash_cross B={B} H={H} Lq={Lq} Lk={Lk} D={D}",
        f"o={e[0]:.3e} dq={e[1]:.3e} dk={e[2]:.3e} dv={e[3]:.3e}")


def time_attn_bwd(B, H, L, D, iters=10):
    q, k, v = (torch.randn(B, H, L, D, device=dev).to(BF) for _ in range(3))
    dq, dk, dv = (torch.empty_like(q) for _ in range(3))
    out = torch.empty(B, L, H * D, device=dev, dtype=BF)
    do = torch.randn(B, L, H * D, device=dev).to(BF)
    lse = K.flash_attn_fwd(q, k, v, D ** -0.5, out)
    for _ in range(3):
        K.flash_attn_bwd(q, k, v, lse, D ** -0.5, out, do, dq, dk, dv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.flash_attn_bwd(q, k, v, lse, D ** -0.5, out, do, dq, dk, dv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 10.0 * B * H * L * L * D
    say(f"TIME flash_bwd B={B} H={H} L={L} D={D}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s (algorithmic 5 GEMMs)")


def time_attn(B, H, L, D, iters=10):
    q, k, v = (torch.randn(B, H, L, D, device=dev).to(BF) for _ in range(3))
    out = torch.empty(B, L, H * D, device=dev, dtype=BF)
    for _ in range(3):
        K.flash_attn_fwd(q, k, v, D ** -0.5, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.flash_attn_fwd(q, k, v, D ** -0.5, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * L * L * D
    for _ in range(3):
        F.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        F.scaled_dot_product_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    say(f"TIME flash_fwd B={B} H={H} L={L} D={D}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s | torch SDPA {ms2:.3f} ms {fl / ms2 / 1e9:.0f} TFLOP/s")


def time_hbm_kernels():
    """HBM-bound kernels at the step's sizes: achieved GB/s of algorithmic bytes (inputs rotate over > L2)."""
    def run(name, fn, nbytes, iters=20):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        say(f"TIME {name}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s")
    for (B, L, C) in [(32, 442, 3072), (32, 577, 1024)]:
        xs = [torch.randn(B, L, C, device=dev).to(BF) for _ in range(4)]
        dys = [torch.randn(B, L, C, device=dev).to(BF) for _ in range(4)]
        mod = torch.randn(B, 2 * C, device=dev).to(BF)
        out = torch.empty(B, L, C, device=dev, dtype=BF)
        n = B * L * C * 2
        _, mean, rstd = K.layernorm_fwd(xs[0], shift=mod[:, :C], scale=mod[:, C:])
        run(f"ln_fwd adaln {B}x{L}x{C}", lambda i: K.layernorm_fwd(xs[i % 4], shift=mod[:, :C], scale=mod[:, C:], out=out), 2 * n)
        run(f"ln_bwd_dx(+dres) {B}x{L}x{C}", lambda i: K.layernorm_bwd_dx(dys[i % 4], xs[i % 4], mean, rstd, scale=mod[:, C:],
                                                                        dres=dys[(i + 1) % 4], out=out), 4 * n)
        acc = torch.zeros(C, device=dev)
        run(f"colsum {B * L}x{C}", lambda i: K.colsum(dys[i % 4].view(-1, C), acc), n)


def main():
    say("== selftest2", torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["ln", "gate", "rope", "misc", "attn", "time"]
    if "ln" in which:
        test_ln()
    if "gate" in which:
        test_gate_colsum()
    if "rope" in which:
        test_rope_qknorm()
    if "misc" in which:
        test_misc()
    if "attn" in which:
        attn_case(1, 1, 64, 128)
        attn_case(1, 1, 128, 128)
        attn_case(1, 2, 200, 128)
        attn_case(2, 3, 442, 128)
        attn_case(2, 3, 442, 128, n_split=1)
        attn_case(2, 2, 1017, 128, n_split=576)
        attn_case(1, 1, 64, 64)
        attn_case(2, 4, 577, 64, fused_qkv=True)
        attn_case(2, 4, 257, 64, fused_qkv=True)
    if "attnbwd" in which or "attnbwd_cases" in which:
        attn_bwd_case(1, 1, 64, 128)
        attn_bwd_case(1, 1, 128, 128)
        attn_bwd_case(1, 2, 200, 128)
        attn_bwd_case(2, 3, 442, 128, n_split=1)
        attn_bwd_case(2, 2, 1017, 128, n_split=576)
        attn_bwd_case(1, 1, 64, 64)
        attn_bwd_case(2, 4, 577, 64, fused_qkv=True)
        attn_cross_bwd_case(2, 2, 1, 16, 128)
        attn_cross_bwd_case(3, 16, 1, 729, 128)
        attn_cross_bwd_case(2, 4, 70, 333, 64)
    if "attnbwd" in which:
        time_attn_bwd(32, 24, 442, 128)
        time_attn_bwd(32, 16, 577, 64)
    if "hbm" in which:
        time_hbm_kernels()
    if "time" in which:
        time_attn(32, 24, 442, 128)
        time_attn(32, 16, 577, 64)
        time_attn(8, 24, 2169, 128)
    say("== done")


if __name__ == "__main__":
    main()
