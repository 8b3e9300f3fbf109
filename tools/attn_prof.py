"""Attention forward / backward: per-kernel durations (CUPTI through torch.profiler) at the shapes of the BASELINE configs.
(The in-kernel phase counters that led to the second form of the backward -- gh_debug_attn_prof, A/B build
-DGH_ATTN_BWD_V1 -- are recorded in profiles/r02_attn_phase_counters.txt.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from genhancer_b200 import _lib, kernels as K

dev, BF = "cuda", torch.bfloat16
shapes = [(32, 24, 442, 128, 0), (32, 16, 729, 128, 80), (10, 24, 1593, 128, 0), (10, 24, 2169, 128, 0), (32, 16, 577, 64, 0)]
for (B, H, L, D, dvalid) in shapes:
    g = torch.Generator(device=dev).manual_seed(L)
    q, k, v = (torch.randn(B, H, L, D, device=dev, generator=g).to(BF) for _ in range(3))
    o = torch.empty(B, L, H * D, device=dev, dtype=BF)
    do = torch.randn(B, L, H * D, device=dev, generator=g).to(BF)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    lse = K.flash_attn_fwd(q, k, v, D ** -0.5, o, d_valid=dvalid)
    for _ in range(3):
        K.flash_attn_bwd(q, k, v, lse, D ** -0.5, o, do, dq, dk, dv, d_valid=dvalid)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            K.flash_attn_fwd(q, k, v, D ** -0.5, o, d_valid=dvalid)
            K.flash_attn_bwd(q, k, v, lse, D ** -0.5, o, do, dq, dk, dv, d_valid=dvalid)
        torch.cuda.synchronize()
    dreal = dvalid or D
    fl = 2.0 * B * H * L * L * dreal
    line = []
    for e in prof.key_averages():
        if "flash" in e.key or "attn" in e.key:
            us = e.device_time_total / e.count
            n = {"flash_fwd": 2, "dkv": 4, "dq": 3}
            gemms = next((c for s, c in n.items() if s in e.key), 0)
            name = e.key.split("(")[0].replace("void gh::", "")
            line.append(f"{name} {us:.0f} us" + (f" ({gemms * fl / us / 1e6:.0f} TFLOP/s executed)" if gemms else ""))
    print(f"B={B} H={H} L={L} D={D} dvalid={dvalid}: " + "; ".join(line))
