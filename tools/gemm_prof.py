"""Where does a GEMM / conv launch spend its time?  Runs the shapes of the stage-1 step with the kernel's per-CTA
pipeline counters on (gh_debug_gemm_prof) and prints, averaged over CTAs: issuer cycles, the share it waited for
TMA data, the share it waited for a free TMEM accumulator (= epilogue back-pressure), epilogue waiting for MMA,
producer waiting for a free smem slot.  Timing (CUDA events, counters off) is printed next to it."""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import _lib, kernels as K

BF = torch.bfloat16
dev = "cuda"
os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/gemm_prof.log", "a")


def say(s):
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def prof(fn, name, flops):
    ms = timed(fn)
    buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    _lib.lib().gh_debug_gemm_prof(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    _lib.lib().gh_debug_gemm_prof(None)
    c = buf.view(148, 8).double()
    act = c[:, 3] > 0
    c = c[act]
    tot, wf, wa, nt, et, ew, ps, wl = (c[:, i].mean().item() for i in range(8))
    say(f"{name:58s} {ms * 1e3:8.1f} us {flops / ms / 1e9:7.0f} TF/s | issuer {tot:9.0f} cyc, tiles/CTA {nt:5.1f}, "
        f"cyc/tile {tot / max(nt, 1):7.0f} | wait-TMA {100 * wf / tot:4.1f}% wait-acc {100 * wa / tot:4.1f}% | "
        f"epi wait-mma {100 * ew / max(et, 1):4.1f}% lean/tile {wl / max(nt, 1):6.0f} | producer wait-slot {100 * ps / tot:4.1f}%")


def gemm(M, N, Kd, a_mn=False, b_mn=False, bias=False, act=0, res=False, f32=False, name=""):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(BF)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(BF)
    kw = {}
    if bias:
        kw["bias"] = torch.randn(N, device=dev)
    if res:
        kw["residual"] = torch.randn(M, N, device=dev).to(BF)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else BF)
    fn = lambda: K.gemm(A, B, a_mn=a_mn, b_mn=b_mn, act=act, out=out, **kw)
    prof(fn, f"{name} M={M} N={N} K={Kd} mn={int(a_mn)}{int(b_mn)} b={int(bias)} act={act} res={int(res)}", 2.0 * M * N * Kd)


def conv(B, H, C, Co, name=""):
    x = torch.randn(B, H, H, C, device=dev).to(BF)
    w = torch.randn(Co, 9 * C, device=dev).to(BF) * 0.02
    bias = torch.randn(Co, device=dev)
    fn = lambda: K.conv2d_nhwc(x, w, 3, 3, 1, 1, bias=bias)
    prof(fn, f"{name} conv B={B} H={H} Cin={C} Cout={Co}", 2.0 * B * H * H * Co * 9 * C)


def main():
    say("== gemm_prof " + torch.cuda.get_device_name(0))
    gemm(18464, 4096, 1024, bias=True, act=2, name="vit fc1")
    gemm(18464, 3072, 1024, bias=True, name="vit qkv")
    gemm(18464, 1024, 1024, bias=True, res=True, name="vit out")
    gemm(18464, 1024, 4096, bias=True, res=True, name="vit fc2")
    gemm(14144, 3072, 15360, bias=True, name="dit linear2")
    gemm(14144, 12288, 3072, bias=True, act=1, name="dit mlp fc1")
    gemm(14144, 12288, 3072, b_mn=True, name="dit dgrad")
    gemm(21504, 3072, 14144, a_mn=True, b_mn=True, name="dit wgrad")
    gemm(8192, 8192, 8192, name="square")
    conv(32, 336, 128, 128, "ae")
    conv(32, 168, 256, 256, "ae")
    conv(32, 84, 512, 512, "ae")


if __name__ == "__main__":
    main()
