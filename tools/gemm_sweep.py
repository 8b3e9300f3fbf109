"""Timing sweep of gh_gemm_bf16 on shapes of the stage-1 step (CUDA events, L2 flushed between iterations by
rotating over distinct buffers).  `python tools/gemm_sweep.py [ncu]` -- with `ncu` only the capture shape runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import kernels as K

dev = "cuda"
BF = torch.bfloat16


def run(M, N, Kd, bias=False, act=0, residual=False, b_mn=False, a_mn=False, iters=10, nbuf=4, f32bias=True, tag=""):
    As = [torch.randn((Kd, M) if a_mn else (M, Kd), device=dev, dtype=BF) for _ in range(nbuf)]
    Bs = [torch.randn((Kd, N) if b_mn else (N, Kd), device=dev, dtype=BF) for _ in range(nbuf)]
    outs = [torch.empty(M, N, device=dev, dtype=BF) for _ in range(nbuf)]
    kw = {}
    if bias:
        kw["bias"] = torch.randn(N, device=dev, dtype=torch.float32 if f32bias else BF)
    res = [torch.randn(M, N, device=dev, dtype=BF) for _ in range(nbuf)] if residual else None
    for i in range(3):
        K.gemm(As[i % nbuf], Bs[i % nbuf], a_mn=a_mn, b_mn=b_mn, out=outs[i % nbuf], act=act,
               residual=res[i % nbuf] if res else None, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        K.gemm(As[i % nbuf], Bs[i % nbuf], a_mn=a_mn, b_mn=b_mn, out=outs[i % nbuf], act=act,
               residual=res[i % nbuf] if res else None, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{tag:28s} M={M} N={N} K={Kd} bias={int(bias)} act={act} res={int(residual)} a_mn={int(a_mn)} b_mn={int(b_mn)}: "
          f"{ms:.4f} ms  {2.0 * M * N * Kd / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        run(18464, 4096, 1024, bias=True, act=2, iters=2, nbuf=2, tag="vit fc1 (capture)")
        run(14144, 3072, 15360, bias=True, iters=2, nbuf=2, tag="dit linear2 (capture)")
        sys.exit(0)
    for Kd in (256, 512, 1024, 2048, 4096):
        run(18464, 4096, Kd, tag="pure")
    for Kd in (256, 512, 1024, 2048, 4096):
        run(18464, 4096, Kd, bias=True, act=2, tag="bias+quick_gelu")
    run(18464, 4096, 1024, bias=True, tag="bias only")
    run(18464, 1024, 1024, bias=True, residual=True, tag="vit out_proj")
    run(18464, 1024, 4096, bias=True, residual=True, tag="vit fc2")
    run(18464, 3072, 1024, bias=True, tag="vit qkv")
    run(14144, 3072, 15360, bias=True, tag="dit linear2")
    run(14144, 12288, 3072, bias=True, act=1, tag="dit fc1")
    run(8192, 8192, 8192, tag="square")
    run(128 * 148, 256, 1024, tag="one wave, 1 tile/SM")
    run(128 * 148, 512, 1024, tag="two tiles/SM")
    run(128 * 148, 1024, 1024, tag="four tiles/SM")
