"""One launch (after one warm-up) of every kernel VERDICT r01 asked ncu evidence for, at the shapes of the step:

    ncu --set full --clock-control none --import-source on -k regex:"flash|conv3x3|ln_fwd|attn_bwd_prep" \
        -o gpurun_out/prof_r02_kernels python tools/ncu_targets.py

  flash_fwd<64>      ViT-L/14-336 self-attention          B=32 H=16 L=577 D=64
  flash_fwd<128>     DiT joint attention                  B=32 H=24 L=442 D=128
  flash_bwd_*<128>   its backward (prep + dK/dV + dQ)
  flash_*<128>       SigLIP-so400m-384: 72-wide heads in 128-lane slots, d_valid = 80   B=32 H=16 L=729
  conv3x3_res        AE level-0 ResnetBlock conv          B=32 336x336 128 -> 128
  ln_fwd             DiT AdaLN                            14144 x 3072
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import kernels as K

dev, BF = "cuda", torch.bfloat16


def attn(B, H, L, D, d_valid=0, bwd=True):
    g = torch.Generator(device=dev).manual_seed(L)
    dr = d_valid or D
    def mk():
        t = torch.zeros(B, H, L, D, device=dev, dtype=BF)
        t[..., :dr] = torch.randn(B, H, L, dr, device=dev, generator=g).to(BF)
        return t
    q, k, v = mk(), mk(), mk()
    o = torch.empty(B, L, H * D, device=dev, dtype=BF)
    do = torch.randn(B, L, H * D, device=dev, generator=g).to(BF)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    for _ in range(2):
        lse = K.flash_attn_fwd(q, k, v, dr ** -0.5, o, d_valid=d_valid)
        if bwd:
            K.flash_attn_bwd(q, k, v, lse, dr ** -0.5, o, do, dq, dk, dv, d_valid=d_valid)
    torch.cuda.synchronize()


def main():
    attn(32, 16, 577, 64)
    attn(32, 24, 442, 128)
    attn(32, 16, 729, 128, d_valid=80)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(32, 336, 336, 128, device=dev, generator=g).to(BF)
    w = (torch.randn(128, 9 * 128, device=dev, generator=g) * 0.03).to(BF)
    b = torch.randn(128, device=dev, generator=g)
    for _ in range(2):
        y = K.conv2d_nhwc(x, w, 3, 3, 1, 1, bias=b, residual=x)
    del y
    xx = torch.randn(32, 442, 3072, device=dev, generator=g).to(BF)
    mod = torch.randn(32, 2 * 3072, device=dev, generator=g).to(BF)
    for _ in range(2):
        K.layernorm_fwd(xx, shift=mod[:, :3072], scale=mod[:, 3072:], eps=1e-6)
    torch.cuda.synchronize()
    print("ncu_targets done")


if __name__ == "__main__":
    main()
