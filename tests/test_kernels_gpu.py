"""Kernel-level parity on the B200, through the C ABI: every sm_100a kernel against a plain PyTorch fp32 reference
of the same op (tolerances are the bf16 rounding floor of the outputs; integer/fp32 paths are checked exactly).
The case generators live in tools/gpu_selftest*.py (also runnable standalone for bring-up / timing)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(modname, groups):
    import sys
    argv, sys.argv = sys.argv, [modname] + groups
    try:
        mod = importlib.import_module(modname)
        del mod.LINES[:]
        mod.main()
        torch.cuda.synchronize()
        lines = list(mod.LINES)
    finally:
        sys.argv = argv
    fails = [l for l in lines if l.startswith("FAIL")]
    passes = [l for l in lines if l.startswith("PASS")]
    assert not fails, "\n".join(fails)
    assert passes, "no case ran"
    return lines


@pytest.mark.parametrize("group", ["fm", "gemm", "mn", "epi", "lora", "batched", "dyn"])
def test_gemm_and_flow_matching_kernels(group):
    _run("tools.gpu_selftest", [group])


@pytest.mark.parametrize("group", ["ln", "gate", "rope", "misc", "attn", "attnbwd_cases"])
def test_norm_rope_attention_kernels(group):
    _run("tools.gpu_selftest2", [group])


@pytest.mark.parametrize("shape", [(23328, 1152, 48, 16), (2000, 4304, 16, 16), (1000, 136, 16, 16), (33, 72, 32, 0), (32, 1152, 16, 16)])
def test_fused_lora_dropout_kernels_match_the_unfused_pair(shape):
    """gh_lora_dropout_fwd / _bwd (mask applied in registers around a warp-level MMA) against dropout kernel + skinny GEMM:
    the SAME mask bit for bit, u and dx to bf16 precision, the ones column of the bias-gradient trick."""
    from tools import lora_fused_check
    assert lora_fused_check.case(*shape)


def test_fm_interp_is_bit_exact_and_loss_matches():
    from genhancer_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    x1 = torch.randn(32, 441, 64, device="cuda", generator=g)
    x0 = torch.randn(32, 441, 64, device="cuda", generator=g)
    t = torch.sigmoid(torch.randn(32, device="cuda", generator=g))
    ref = ((1 - t[:, None, None]) * x1 + t[:, None, None] * x0).to(torch.bfloat16)
    assert torch.equal(K.fm_interp(x1, x0, t), ref)
    pred = torch.randn(32, 441, 64, device="cuda", generator=g).to(torch.bfloat16)
    loss, dpred = K.fm_mse_loss(pred, x0, x1)
    ref_loss = torch.nn.functional.mse_loss(pred.float(), x0 - x1)
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-5
    # linearity of the gradient in grad_scale (size-independent property)
    _, d2 = K.fm_mse_loss(pred, x0, x1, grad_scale=2.0)
    assert torch.allclose(d2.float(), 2 * dpred.float(), rtol=1e-2, atol=1e-9)


def test_empty_and_bad_arguments_fail_loudly():
    from genhancer_b200 import _lib, kernels as K
    a = torch.zeros(16, 24, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.GhError):
        K.gemm(a, torch.zeros(16, 32, device="cuda", dtype=torch.bfloat16))  # K mismatch
    with pytest.raises(_lib.GhError):
        b = torch.zeros(16, 20, device="cuda", dtype=torch.bfloat16)
        K.gemm(b, b)  # row pitch 20 elements: not a multiple of 8 (TMA needs 16-byte row pitch)
    with pytest.raises(_lib.GhError):
        K.gemm(a.float(), a.float())


def test_u8_to_tensor_is_bit_identical_to_torch():
    """gh_u8hwc_to_f32chw == torchvision ToTensor as the reference's DataLoader runs it -- on the CPU, where
    ``.div(255)`` is an IEEE division (torch's CUDA kernel multiplies by the reciprocal instead and differs in the
    last bit for 126 of the 256 byte values).  Every byte value, odd sizes, B = 0."""
    from genhancer_b200 import _lib, kernels as K
    g = torch.Generator(device="cuda").manual_seed(5)
    for shape in [(2, 16, 16, 3), (3, 37, 53, 3), (32, 336, 336, 3), (0, 8, 8, 3)]:
        x = torch.randint(0, 256, shape, device="cuda", dtype=torch.uint8, generator=g)
        if x.numel() >= 256 * 3:
            x.view(-1)[:256] = torch.arange(256, device="cuda", dtype=torch.uint8)     # every value at least once
        y = K.u8hwc_to_f32chw(x)
        assert y.shape == (shape[0], 3, shape[1], shape[2]) and y.dtype == torch.float32
        assert torch.equal(y.cpu(), x.cpu().permute(0, 3, 1, 2).to(torch.float32).div(255))
    with pytest.raises(_lib.GhError):
        K.u8hwc_to_f32chw(torch.zeros(1, 3, 8, 8, device="cuda", dtype=torch.uint8))       # CHW is not the input layout


@pytest.mark.parametrize("B,H,W,res", [(4, 90, 100, False), (4, 90, 100, True), (4, 112, 112, True), (2, 336, 336, True)])
def test_conv3x3_resident_weight_kernel_matches_torch(B, H, W, res):
    """conv3x3_res_kernel (Cin = Cout = 128, 3x3 / s1 / p1: weights resident in smem, one input box per column shift,
    taps as 1024-byte offsets into it) against torch conv2d, incl. ragged right / bottom patches and zero padding."""
    import torch.nn.functional as F
    from genhancer_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    C = 128
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w4 = (torch.randn(C, C, 3, 3, device="cuda", generator=g) * 0.05)
    wk = w4.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous().to(torch.bfloat16)     # k = (kh * 3 + kw) * Cin + ci
    bias = torch.randn(C, device="cuda", generator=g)
    r = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16) if res else None
    y = K.conv2d_nhwc(x, wk, 3, 3, 1, 1, bias=bias, residual=r)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wk.float().view(C, 3, 3, C).permute(0, 3, 1, 2), bias, padding=1)
    ref = ref.permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    err = (y.float() - ref).norm() / ref.norm()
    assert err < 6e-3, err
    # borders (zero padding through TMA out-of-bounds fill) and the ragged last patches on their own
    for a, b in zip((y[:, 0], y[:, -1], y[:, :, 0], y[:, :, -1]), (ref[:, 0], ref[:, -1], ref[:, :, 0], ref[:, :, -1])):
        assert (a.float() - b).norm() / b.norm() < 6e-3


def test_uint8_hwc_gathers_are_bit_identical_to_the_fp32_path():
    """f-4: the decoded uint8 HWC batch feeds the patch-embed / conv_in gathers directly (u8 / 255 on the fly):
    same bits as ToTensor on the device followed by the fp32 gathers, at 1 byte per value."""
    from genhancer_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    u8 = torch.randint(0, 256, (3, 56, 56, 3), device="cuda", generator=g, dtype=torch.uint8)
    f32 = K.u8hwc_to_f32chw(u8)
    assert torch.equal(f32.cpu(), u8.cpu().permute(0, 3, 1, 2).float().div(255))   # the CPU ToTensor of the reference's loader
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    a = K.patch_im2col(f32, 14, 592, mean, std)
    b = K.patch_im2col(u8, 14, 592, mean, std)
    assert a.shape == b.shape == (3 * 16, 588) and torch.equal(a, b)
    c = K.im2col3x3_c3(f32, 0.5, 0.5)
    d = K.im2col3x3_c3(u8, 0.5, 0.5)
    assert c.shape == d.shape == (3 * 56 * 56, 32) and torch.equal(c, d)
    assert K.image_bhw(u8) == (3, 56, 56) and K.image_bhw(f32) == (3, 56, 56)


@pytest.mark.parametrize("B,S,D,bias,u8", [(3, 56, 128, False, False), (2, 224, 1024, False, False), (2, 84, 1152, True, False),
                                           (5, 42, 200, True, True), (33, 336, 1024, False, True)])
def test_patch_embed_implicit_gemm_matches_conv2d(B, S, D, bias, u8):
    """gh_patch_embed_fwd (north_star bullet 1: the patch-embed conv as an implicit GEMM, operand gathered from the image
    into the MMA's smem layout, ToTensor / Normalize folded in) against torch conv2d on the same bf16-rounded operands,
    and against the older gather + GEMM pair (same bits up to fp32 accumulation order)."""
    import torch.nn.functional as F
    from genhancer_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(B * 100 + S)
    p = 14
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    if u8:
        img = torch.randint(0, 256, (B, S, S, 3), device="cuda", generator=g, dtype=torch.uint8)
        f32 = K.u8hwc_to_f32chw(img)
    else:
        img = torch.rand(B, 3, S, S, device="cuda", generator=g)
        f32 = img
    kdim, ld = 3 * p * p, (3 * p * p + 7) // 8 * 8
    w4 = torch.randn(D, 3, p, p, device="cuda", generator=g) * 0.05
    wk = torch.zeros(D, ld, device="cuda", dtype=torch.bfloat16)
    wk[:, :kdim] = w4.reshape(D, kdim).to(torch.bfloat16)
    bv = torch.randn(D, device="cuda", generator=g) if bias else None
    out = K.patch_embed(img, wk[:, :kdim], p, bias=bv, mean=mean, std=std)
    G = S // p
    assert out.shape == (B * G * G, D)
    # reference: normalise in fp32, round operands to bf16 as the kernels do, conv in fp32
    m = torch.tensor(mean, device="cuda").view(1, 3, 1, 1)
    s = torch.tensor(std, device="cuda").view(1, 3, 1, 1)
    xn = ((f32 - m) * (1.0 / s)).to(torch.bfloat16).float()
    ref = F.conv2d(xn, wk[:, :kdim].float().view(D, 3, p, p), bv, stride=p).permute(0, 2, 3, 1).reshape(B * G * G, D)
    assert (out.float() - ref).norm() / ref.norm() < 5e-3
    old = K.gemm(K.patch_im2col(img, p, ld, mean, std), wk[:, :kdim], bias=bv)
    assert (out.float() - old.float()).abs().max().item() <= 2e-2 * ref.abs().max().item()
