"""Pipeline phase counters of the attention dK/dV kernel (gh_debug_attn_prof): where one CTA's cycles go."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import _lib, kernels as K

dev, BF = "cuda", torch.bfloat16
for (B, H, L, D, dvalid) in [(32, 24, 442, 128, 0), (32, 16, 729, 128, 80), (32, 16, 577, 64, 0)]:
    g = torch.Generator(device=dev).manual_seed(L)
    q, k, v = (torch.randn(B, H, L, D, device=dev, generator=g).to(BF) for _ in range(3))
    o = torch.empty(B, L, H * D, device=dev, dtype=BF)
    do = torch.randn(B, L, H * D, device=dev, generator=g).to(BF)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    lse = K.flash_attn_fwd(q, k, v, D ** -0.5, o, d_valid=dvalid)
    K.flash_attn_bwd(q, k, v, lse, D ** -0.5, o, do, dq, dk, dv, d_valid=dvalid)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().gh_debug_attn_prof(buf.data_ptr()))
    K.flash_attn_bwd(q, k, v, lse, D ** -0.5, o, do, dq, dk, dv, d_valid=dvalid)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().gh_debug_attn_prof(None))
    c = buf.tolist()
    nq = max(c[7], 1)
    print(f"dkv B={B} H={H} L={L} D={D} dvalid={dvalid}: {nq} query blocks; compute warp 0 per block: wait S/dP {c[0] / nq:.0f}, "
          f"stat store + bar.sync {c[1] / nq:.0f}, TMEM ld + math {c[2] / nq:.0f}, wait prev dV/dK {c[3] / nq:.0f}, "
          f"smem store + fence + arrive {c[4] / nq:.0f}; loop {c[5]} ({c[5] / nq:.0f} per block), epilogue {c[6]}; "
          f"control: K/V + first Q/dO load {c[11]}, wait Q/dO {c[8] / nq:.0f}, wait P/dS {c[9] / nq:.0f} per block, total {c[10]}")
