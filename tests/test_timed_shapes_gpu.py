"""Parity at the shapes that are BENCHMARKED (VERDICT r01 "pin parity at the shapes that are timed"):

  * BASELINE configs[1]: the whole stage-1 step at 336x336, batch 32 (577-token ViT-L/14, 442-token DiT, M = 14144) against
    the oracle port running on the same GPU through stock PyTorch ops with the SAME weights, inputs and RNG draws;
  * the full SigLIP-so400m-384 tower (27 x 1152, 729 tokens, head_dim 72 in 128-lane slots, MAP head);
  * attention forward + backward at the sequence lengths of the video modes (L = 1593, 2169) and of SigLIP (729), for
    head_dim 128 and for 72-wide heads zero-padded into 128-wide slots, against torch SDPA autograd;
  * MetaCLIP-H geometry (head_dim 80) against the reference's MetaCLIP wrapper (fixture);
  * ``prepare_clip`` against the reference's (fixture); the sliding-window gather on DEVICE tensors, bit for bit.
"""
import math
import random

import pytest
import torch
import torch.nn.functional as F

from conftest import cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _bench():
    import bench
    return bench


def test_cfg2_step_336_batch32_matches_oracle_on_gpu():
    """The step bench.py times: loss within 1e-2 (north_star), vec / txt cosine >= 0.999, gradient cosines >= 0.99."""
    from oracle import genhancer_oracle as O
    bench = _bench()
    dev = torch.device("cuda", 0)
    B, S = 32, 336
    w = bench.build_workload("img336_stage1", dev, B)
    gen = torch.Generator(device=dev).manual_seed(1234)
    x = w.make_inputs(gen)
    par = bench.oracle_parity(w, x, dev)
    print("cfg[1] parity at 336^2, B=32:", par)
    assert par["parity_rel"] <= 1e-2
    assert par["cos_vec"] >= 0.999 and par["cos_txt"] >= 0.999 and par["cos_x1"] >= 0.999
    assert par["cos_pred"] >= 0.995
    # gradients at the timed shape: ours vs autograd of the oracle port (fp32 tower + AE, bf16 DiT as the reference)
    h = S // 8
    g2 = torch.Generator(device=dev).manual_seed(99)
    noise = torch.randn(B, 16, h, h, device=dev, generator=g2)
    t = torch.sigmoid(torch.randn(B, device=dev, generator=g2))
    x_0 = torch.randn(B, (h // 2) ** 2, 64, device=dev, generator=g2)
    for p in w.dit.parameters():
        p.grad = None
    w.dit._grad_overwrite = True
    loss = w.step(x[0], ae_noise=noise, t=t, x_0=x_0)
    loss.backward()
    names = ["final_layer.linear.weight", "single_blocks.3.linear2.weight", "double_blocks.0.img_attn.qkv.weight",
             "double_blocks.1.img_mlp.2.bias", "single_blocks.0.linear2.bias", "double_blocks.0.img_attn.proj.bias"]
    P = dict(w.dit.named_parameters())
    ours = {n: P[n].grad.detach().float().clone() for n in names}
    ours["project_t5.3.weight"] = w.clip_vis.project_t5[3].weight.grad.detach().float().clone()
    ours_loss = float(loss)
    tc, fc, ac = O.openai_vit_l14(S), O.FluxCfg(), O.AECfg()
    sd_t = {k: v.detach().float() for k, v in w.clip_vis.model.state_dict().items()}
    sd_w = {k: v.detach().float().clone().requires_grad_(True) for k, v in w.clip_vis.state_dict().items() if k.startswith("project_")}
    sd_d = {k: v.detach().clone().requires_grad_(True) for k, v in w.dit.state_dict().items()}
    sd_a = {k: v.detach().float() for k, v in w.vae.encoder.state_dict().items()}
    del w, P
    torch.cuda.empty_cache()
    out = O.stage1_image_step(sd_t, sd_w, sd_d, sd_a, x[0], tc, fc, ac, (0.48145466, 0.4578275, 0.40821073),
                              (0.26862954, 0.26130258, 0.27577711), noise, t, x_0, dit_dtype=torch.bfloat16)
    out.loss.backward()
    assert abs(ours_loss - float(out.loss)) / float(out.loss) <= 1e-2
    for n in names:
        c = cosine(ours[n], sd_d[n].grad)
        print(f"  grad cosine {n}: {c:.5f}")
        assert c >= 0.99, (n, c)
    c = cosine(ours["project_t5.3.weight"], sd_w["project_t5.3.weight"].grad)
    print(f"  grad cosine project_t5.3.weight: {c:.5f}")
    assert c >= 0.99


def test_siglip_so400m_384_full_tower_matches_oracle_on_gpu():
    """SigLIP-so400m-384 at FULL size (27 layers x 1152, 729 tokens, 16 heads of 72 in 128-lane slots, MAP pooling head)
    against the oracle tower (fp32, stock PyTorch ops) on the same GPU, same random-init weights."""
    from genhancer_b200.clip_models.build_CLIP import load_clip_model_SigLIP
    from oracle import genhancer_oracle as O
    import contextlib, io, warnings
    dev = torch.device("cuda", 0)

    class C:
        clip_image_size, clip_dim, t5_dim = 384, 768, 4096
    torch.manual_seed(5)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        wrap = load_clip_model_SigLIP(C, dev)
    img = torch.rand(4, 3, 384, 384, device=dev)
    x = (img - 0.5) / 0.5
    with torch.no_grad():
        out = wrap.model.vision_model(x, output_hidden_states=True)
        cls, pc, pt5 = wrap(x)
        cls2, _, _ = wrap(img, _norm=((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)))
        tc = O.siglip_so400m(384)
        sd_t = {k: v.detach().float() for k, v in wrap.model.state_dict().items()}
        sd_w = {k: v.detach().float() for k, v in wrap.state_dict().items() if k.startswith("project_")}
        lhs, pooled = O.tower_forward(sd_t, x, tc)
        ocls, opc, opt5 = O.clip_wrapper_forward(sd_t, sd_w, x, tc)
    assert out.last_hidden_state.shape == (4, 729, 1152)
    cs = dict(lhs=cosine(out.last_hidden_state, lhs), pooled=cosine(out.pooler_output, pooled), cls=cosine(cls, ocls),
              pc=cosine(pc, opc), pt5=cosine(pt5, opt5), fold=cosine(cls2, cls))
    print("SigLIP-so400m-384 full tower cosines:", cs)
    assert min(cs["lhs"], cs["pooled"], cs["cls"], cs["pc"], cs["pt5"]) >= 0.999 and cs["fold"] >= 0.9999


def _sdpa_ref(q, k, v, do, scale):
    q, k, v = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, scale=scale)
    o.backward(do.float())
    return o.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("L,Dreal", [(729, 128), (1593, 128), (2169, 128), (729, 72), (577, 80)])
def test_attention_fwd_bwd_at_benchmarked_lengths(L, Dreal):
    """flash_fwd / flash_bwd<128> at the joint-attention lengths of the video modes (441 + 1152, 441 + 1728) and of
    SigLIP-384 (729 tokens); 72- and 80-wide heads run zero-padded in 128-lane slots with scale = Dreal^-0.5."""
    from genhancer_b200 import kernels as K
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(L + Dreal)
    B, H, D = 2, 3, 128
    def mk():
        t = torch.zeros(B, H, L, D, device=dev, dtype=torch.bfloat16)
        t[..., :Dreal] = torch.randn(B, H, L, Dreal, device=dev, generator=g).to(torch.bfloat16)
        return t
    q, k, v = mk(), mk(), mk()
    scale = Dreal ** -0.5
    o = torch.empty(B, L, H * D, device=dev, dtype=torch.bfloat16)
    lse = K.flash_attn_fwd(q, k, v, scale, o, d_valid=Dreal)          # (72 -> 80 lanes of work, as the towers call it)
    do = torch.zeros(B, L, H, D, device=dev, dtype=torch.bfloat16)
    do[..., :Dreal] = torch.randn(B, L, H, Dreal, device=dev, generator=g).to(torch.bfloat16)
    do = do.view(B, L, H * D)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    K.flash_attn_bwd(q, k, v, lse, scale, o, do, dq, dk, dv, d_valid=Dreal)
    do_h = do.view(B, L, H, D).permute(0, 2, 1, 3)
    ro, rq, rk, rv = _sdpa_ref(q[..., :Dreal], k[..., :Dreal], v[..., :Dreal], do_h[..., :Dreal], scale)
    o_h = o.view(B, L, H, D).permute(0, 2, 1, 3)
    assert rel_err(o_h[..., :Dreal], ro) < 1e-2
    assert rel_err(dq[..., :Dreal], rq) < 2e-2 and rel_err(dk[..., :Dreal], rk) < 2e-2 and rel_err(dv[..., :Dreal], rv) < 2e-2
    if Dreal < D:       # the padding lanes stay exactly zero (so the 128-slot form is exact, not approximate)
        assert float(o_h[..., Dreal:].abs().max()) == 0.0 and float(dq[..., Dreal:].abs().max()) == 0.0
        assert float(dk[..., Dreal:].abs().max()) == 0.0 and float(dv[..., Dreal:].abs().max()) == 0.0
    # log-sum-exp (log2 domain) against the reference's
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * scale
    ref_lse2 = torch.logsumexp(s, dim=-1) / math.log(2.0)
    assert (lse - ref_lse2).abs().max().item() < 2e-2


def test_metaclip_h_geometry_matches_reference():
    """MetaCLIP-H/14 geometry (head_dim 80 = 1280 / 16) at reduced size: fixture from HF CLIPModel inside the
    reference's MetaCLIP(clip_type='huge') wrapper (CLIP_bank.py:76-122)."""
    from test_tower_ae_gpu import OPENAI_MEAN, OPENAI_STD, _build_wrapper
    fx = load_golden("tower_metaclip_h_small.pt")
    assert fx["cfg"]["hidden"] // fx["cfg"]["heads"] == 80
    wrap = _build_wrapper(fx)
    img = fx["img"].to("cuda")
    mean = torch.tensor(OPENAI_MEAN, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD, device="cuda").view(1, 3, 1, 1)
    x = (img - mean) / std
    out = wrap.model.vision_model(x, output_hidden_states=True)
    assert cosine(out.last_hidden_state, fx["last_hidden_state"]) >= 0.999
    assert cosine(out.pooler_output, fx["pooler_output"]) >= 0.999
    cls, pc, pt5 = wrap(x)
    assert cosine(cls, fx["class_token"]) >= 0.999 and cosine(pc, fx["projection_clip"]) >= 0.999
    assert cosine(pt5, fx["projection_t5"]) >= 0.999
    (pc.float().square().mean() + pt5.float().square().mean()).backward()
    assert cosine(wrap.project_t5[1].weight.grad, fx["grad_project_t5_1_weight"]) >= 0.99


def test_prepare_clip_matches_reference():
    """clip_models.sampling.prepare_clip (reference: clip_models/sampling.py:9-42): img / img_ids / txt_ids bit-equal,
    txt / vec at the tower tolerance."""
    from genhancer_b200.clip_models.sampling import prepare_clip
    from test_tower_ae_gpu import _build_wrapper
    fx = load_golden("prepare_clip_small.pt")
    wrap = _build_wrapper(fx)
    with torch.no_grad():
        out = prepare_clip(clip=wrap, original_img=fx["original_img"].to("cuda"), img=fx["latent"].to("cuda"))
    ref = fx["out"]
    assert set(out) == set(ref) == {"img", "img_ids", "txt", "txt_ids", "vec"}
    for k in ref:
        assert out[k].shape == ref[k].shape and out[k].is_cuda, k
    assert torch.equal(out["img"].cpu(), ref["img"])
    assert torch.equal(out["img_ids"].cpu().float(), ref["img_ids"])
    assert torch.equal(out["txt_ids"].cpu().float(), ref["txt_ids"])
    assert cosine(out["txt"], ref["txt"]) >= 0.999 and cosine(out["vec"], ref["vec"]) >= 0.999


def test_window_gather_on_device_is_bit_equal_to_reference():
    """build_windows_with_mask on DEVICE tensors (ragged masks) == the reference's per-frame stack loops (fixture)."""
    from genhancer_b200.video import build_windows_with_mask
    fx = load_golden("video_step_small.pt")
    w = fx["windows"]
    fr, mask = fx["frames"].to("cuda"), fx["frame_mask"].to("cuda")
    random.seed(0)
    c0, c1, c2, tgt, avg_nw, bs_eff = build_windows_with_mask(fr, mask, 3, 1, 8)
    assert all(t.is_cuda for t in (c0, c1, c2, tgt))
    assert torch.equal(c0.cpu(), w["cond0"]) and torch.equal(c2.cpu(), w["cond2"]) and torch.equal(tgt.cpu(), w["target"])
    assert avg_nw == w["avg_nw"] and bs_eff == w["bs_eff"]
    assert build_windows_with_mask(fr[:, :3], mask[:, :3], 3, 1, 8) is None
