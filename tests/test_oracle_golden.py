"""CPU: the oracle (oracle/genhancer_oracle.py) against the golden fixtures, which hold outputs of the REFERENCE
itself (minted by oracle/make_golden.py from /root/reference in the build container).  This is what pins the
oracle; the GPU tests then compare the CUDA path with the same fixtures."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import genhancer_oracle as O

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def _norm(img, kind):
    mean = torch.tensor(OPENAI_MEAN if kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD if kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    return (img - mean) / std


def _tower(name):
    fx = load_golden(name)
    tc = O.TowerCfg(**fx["cfg"])
    sd_t = O.synth_state_dict(fx["key_shapes_tower"], fx["seed"])
    sd_w = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(fx["key_shapes_wrap"], fx["seed"] + 1).items()}
    x = _norm(fx["img"], tc.kind)
    lhs, pooled = O.tower_forward(sd_t, x, tc)
    cls, pc, pt5 = O.clip_wrapper_forward(sd_t, sd_w, x, tc)
    (pc.square().mean() + pt5.square().mean()).backward()
    assert rel_err(lhs, fx["last_hidden_state"]) < 1e-5
    assert rel_err(pooled, fx["pooler_output"]) < 1e-5
    assert rel_err(cls, fx["class_token"]) < 1e-5
    assert rel_err(pc, fx["projection_clip"]) < 1e-5
    assert rel_err(pt5, fx["projection_t5"]) < 1e-5
    assert rel_err(sd_w["project_t5.1.weight"].grad, fx["grad_project_t5_1_weight"]) < 1e-4
    assert rel_err(sd_w["project_clip.3.bias"].grad, fx["grad_project_clip_3_bias"]) < 1e-4


def test_oracle_tower_clip():
    _tower("tower_clip_small.pt")


def test_oracle_tower_siglip():
    _tower("tower_siglip_small.pt")


def test_oracle_ae_encoder():
    fx = load_golden("ae_small.pt")
    ac = O.AECfg(**fx["cfg"])
    sd = O.synth_state_dict(fx["key_shapes"], fx["seed"])
    x = (fx["img"] - 0.5) / 0.5
    with torch.no_grad():
        assert rel_err(O.ae_encoder_forward(sd, x, ac), fx["moments"]) < 1e-5
        assert rel_err(O.ae_encode(sd, x, ac, fx["noise"]), fx["z"]) < 1e-5


def _flux(name):
    fx = load_golden(name)
    fc = O.FluxCfg(**fx["cfg"])
    sd = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(fx["key_shapes"], fx["seed"]).items()}
    img, txt, y = (fx[k].clone().requires_grad_(True) for k in ("img", "txt", "y"))
    pred = O.flux_forward(sd, fc, img, fx["img_ids"], txt, fx["txt_ids"], fx["t"], y, fx["guidance"])
    loss = F.mse_loss(pred.float(), fx["target"])
    loss.backward()
    assert rel_err(pred, fx["pred"]) < 1e-5
    assert abs(loss.item() - fx["loss"].item()) < 1e-6 * max(1.0, abs(fx["loss"].item()))
    assert rel_err(img.grad, fx["d_img"]) < 1e-4
    assert rel_err(txt.grad, fx["d_txt"]) < 1e-4
    assert rel_err(y.grad, fx["d_y"]) < 1e-4
    for k, g in fx["grads"].items():
        assert rel_err(sd[k].grad, g) < 2e-4, k


def test_oracle_flux_image_mode():
    _flux("flux_img.pt")


def test_oracle_flux_video_ids():
    _flux("flux_video.pt")


def test_oracle_stage1_step_small():
    fx = load_golden("step_small.pt")
    tc, fc, ac = O.TowerCfg(**fx["tower_cfg"]), O.FluxCfg(**fx["flux_cfg"]), O.AECfg(**fx["ae_cfg"])
    ks, seed = fx["key_shapes"], fx["seed"]
    sd_t = O.synth_state_dict(ks["tower"], seed)
    sd_w = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["wrap"], seed + 1).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["dit"], seed + 2).items()}
    sd_a = O.synth_state_dict(ks["ae"], seed + 3)
    out = O.stage1_image_step(sd_t, sd_w, sd_d, sd_a, fx["img"], tc, fc, ac, OPENAI_MEAN, OPENAI_STD, fx["ae_noise"],
                              fx["t"], fx["x_0"])
    out.loss.backward()
    assert rel_err(out.x_1, fx["x_1"]) < 1e-5
    assert rel_err(out.x_t, fx["x_t"]) < 1e-5
    assert rel_err(out.pred, fx["pred"]) < 5e-5
    assert abs(out.loss.item() - fx["loss"].item()) / fx["loss"].item() < 1e-5
    assert rel_err(sd_w["project_t5.3.weight"].grad, fx["grad_project_t5_3_weight"]) < 5e-4
    assert rel_err(sd_d["txt_in.weight"].grad, fx["grad_txt_in_weight"]) < 5e-4
    assert rel_err(sd_d["final_layer.linear.weight"].grad, fx["grad_final_linear_weight"]) < 5e-4


def test_rng_draw_order_reproduces_reference_draws():
    """t and x_0 in the fixture came out of the reference's own torch.randn calls after manual_seed(seed): the
    restated draw order (AE noise -> t -> x_0) must give the identical bits."""
    fx = load_golden("step_small.pt")
    torch.manual_seed(fx["seed"])
    noise = torch.randn(2, 16, 14, 14)
    t = torch.sigmoid(torch.randn((2,)) * 1.0)
    x_0 = torch.randn(2, 49, 64)
    assert torch.equal(noise, fx["ae_noise"]) and torch.equal(t, fx["t"]) and torch.equal(x_0, fx["x_0"])


def test_window_builder_and_ids():
    frames = torch.arange(2 * 8).float().view(2, 8, 1, 1, 1).expand(2, 8, 3, 2, 2)
    mask = torch.tensor([[1] * 8, [1] * 5 + [0] * 3])
    conds, tgt, counts = O.build_windows_with_mask(frames, mask, 3, 1, 8)
    assert counts == [5, 2] and tgt.shape[0] == 7 and len(conds) == 3
    assert tgt[:, 0, 0, 0].tolist() == [3, 4, 5, 6, 7, 11, 12]
    assert conds[0][:, 0, 0, 0].tolist() == [0, 1, 2, 3, 4, 8, 9]
    ids = O.create_spatio_temporal_ids(2, 2, 3, 4)
    assert ids.shape == (2, 12, 3) and ids[0, 5].tolist() == [2.0, 1.0, 1.0]
    assert O.lora_merge(torch.zeros(4, 3), torch.ones(2, 3), torch.ones(4, 2), 0.5).eq(1.0).all()


def test_oracle_video_step_and_window_builder():
    import random
    fx = load_golden("video_step_small.pt")
    tc, fc, ac = O.TowerCfg(**fx["tower_cfg"]), O.FluxCfg(**fx["flux_cfg"]), O.AECfg(**fx["ae_cfg"])
    ks, seed = fx["key_shapes"], fx["seed"]
    sd_t = O.synth_state_dict(ks["tower"], seed)
    sd_ad = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["adapter"], seed + 4).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["dit"], seed + 2).items()}
    sd_a = O.synth_state_dict(ks["ae"], seed + 3)
    fr = fx["frames"]
    out = O.stage1_video_step(sd_t, sd_ad, sd_d, sd_a, [fr[:, 0], fr[:, 1]], fr[:, 2], tc, fc, ac, OPENAI_MEAN, OPENAI_STD,
                              fx["cond_times"], fx["target_time"], fx["ae_noise"], fx["t"], fx["x_0"])
    out.loss.backward()
    assert torch.equal(out.extras["txt_ids"], fx["txt_ids"]) and torch.equal(out.extras["img_ids"], fx["img_ids"])
    assert rel_err(out.txt, fx["txt"]) < 1e-5 and rel_err(out.vec, fx["vec"]) < 1e-5
    assert rel_err(out.pred, fx["pred"]) < 5e-5
    assert abs(out.loss.item() - fx["loss"].item()) / fx["loss"].item() < 1e-5
    assert rel_err(sd_ad["proj.2.weight"].grad, fx["grad_adapter_proj2_weight"]) < 5e-4
    assert rel_err(sd_d["txt_in.weight"].grad, fx["grad_txt_in_weight"]) < 5e-4
    # window builder vs the reference's own function (fixture) -- oracle and product host code
    conds, tgt, counts = O.build_windows_with_mask(fr, fx["frame_mask"].long(), 3, 1, 8, rng=random)
    w = fx["windows"]
    assert torch.equal(conds[0], w["cond0"]) and torch.equal(conds[2], w["cond2"]) and torch.equal(tgt, w["target"])
    from genhancer_b200.video import build_windows_with_mask, create_spatio_temporal_ids
    c0, c1, c2, t2, avg_nw, bs_eff = build_windows_with_mask(fr, fx["frame_mask"], 3, 1, 8)
    assert torch.equal(c0, w["cond0"]) and torch.equal(c2, w["cond2"]) and torch.equal(t2, w["target"])
    assert avg_nw == w["avg_nw"] and bs_eff == w["bs_eff"]
    assert build_windows_with_mask(fr[:, :3], fx["frame_mask"][:, :3], 3, 1, 8) is None      # too short for a window
    g = int(round((fx["txt_ids"].shape[1] // 2) ** 0.5))
    ids = torch.cat([create_spatio_temporal_ids(g, g, t, "cpu") for t in fx["cond_times"]], 0).float()
    assert torch.equal(ids, fx["txt_ids"][0])
    # more windows than the cap: random.sample with the caller's RNG state, then sorted (reference semantics)
    long_fr = torch.arange(14).float().view(1, 14, 1, 1, 1).expand(1, 14, 3, 2, 2)
    random.seed(3)
    a = build_windows_with_mask(long_fr, torch.ones(1, 14, dtype=torch.bool), 3, 1, 4)
    random.seed(3)
    starts = sorted(random.sample(list(range(0, 11)), k=4))
    assert a[3][:, 0, 0, 0].tolist() == [float(s + 3) for s in starts] and a[5] == 4


def test_sampler_schedule_unpack_and_oracle_denoise_match_reference():
    """Fixture: the reference's src/flux/sampling.py driving the reference's Flux (make_golden.py::golden_sampler)."""
    from genhancer_b200.flux import sampling as S
    fx = load_golden("sampler_small.pt")
    n_img = fx["img"].shape[1]
    # host-side pieces of the drop-in module: exact
    assert S.get_schedule(4, n_img, shift=True) == fx["schedule"]
    assert S.get_schedule(7, 1024, shift=False) == fx["schedule_plain"]
    assert S.get_schedule(25, 4096) == fx["schedule_big"]
    assert fx["schedule"][0] == 1.0 and fx["schedule"][-1] == 0.0
    assert torch.equal(S.get_noise(2, fx["height"], fx["width"], torch.device("cpu"), torch.float32, fx["seed"]), fx["noise"])
    assert torch.equal(S.unpack(fx["denoised"], fx["height"], fx["width"]), fx["unpacked"])
    with pytest.raises(NotImplementedError):
        S.prepare(None, None, fx["img"], "a prompt")
    # the oracle restatement
    assert O.get_schedule(4, n_img) == fx["schedule"]
    assert torch.equal(O.unpack(fx["denoised"], fx["height"], fx["width"]), fx["unpacked"])
    sd = O.synth_state_dict(fx["key_shapes"], fx["seed"])
    with torch.no_grad():
        out = O.denoise(sd, O.FluxCfg(**fx["cfg"]), fx["img"], fx["img_ids"], fx["txt"], fx["txt_ids"], fx["vec"], fx["neg_txt"],
                        fx["txt_ids"], fx["neg_vec"], fx["schedule"], 4.0, fx["true_gs"], fx["start_cfg"])
    assert (out - fx["denoised"]).abs().max().item() <= 5e-5 * fx["denoised"].abs().max().item()


def test_oracle_ae_decoder():
    fx = load_golden("ae_decoder_small.pt")
    ac = O.AECfg(**fx["cfg"])
    assert {k: tuple(v) for k, v in O.ae_decoder_key_shapes(ac).items()} == {k: tuple(v) for k, v in fx["key_shapes"].items()}
    sd = O.synth_state_dict(fx["key_shapes"], fx["seed"])
    with torch.no_grad():
        img = O.ae_decode(sd, fx["z"], ac)
    assert img.shape == fx["image"].shape == (2, 3, 48, 80)
    assert (img - fx["image"]).abs().max().item() <= 2e-5 * fx["image"].abs().max().item()


def test_oracle_metaclip_h_geometry_and_prepare_clip():
    """Fixtures minted from the reference's MetaCLIP(clip_type='huge') wrapper (head_dim 80) and from the reference's
    prepare_clip (clip_models/sampling.py:9-42)."""
    fx = load_golden("tower_metaclip_h_small.pt")
    tc = O.TowerCfg(**fx["cfg"])
    assert tc.hidden // tc.heads == 80
    sd_t = O.synth_state_dict(fx["key_shapes_tower"], fx["seed"])
    sd_w = O.synth_state_dict(fx["key_shapes_wrap"], fx["seed"] + 1)
    mean = torch.tensor(OPENAI_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD).view(1, 3, 1, 1)
    with torch.no_grad():
        lhs, pooled = O.tower_forward(sd_t, (fx["img"] - mean) / std, tc)
        cls, pc, pt5 = O.clip_wrapper_forward(sd_t, sd_w, (fx["img"] - mean) / std, tc)
    assert rel_err(lhs, fx["last_hidden_state"]) < 1e-5 and rel_err(pooled, fx["pooler_output"]) < 1e-5
    assert rel_err(cls, fx["class_token"]) < 1e-5 and rel_err(pc, fx["projection_clip"]) < 1e-5
    assert rel_err(pt5, fx["projection_t5"]) < 1e-5
    fp = load_golden("prepare_clip_small.pt")
    ref = fp["out"]
    B, _, h, w = fp["latent"].shape
    assert torch.equal(O.patchify(fp["latent"]), ref["img"])
    assert torch.equal(O.make_img_ids(B, h // 2, w // 2), ref["img_ids"])
    assert ref["txt_ids"].shape == (B, 1, 3) and not ref["txt_ids"].any()
    # the product's host-side pieces of prepare_clip on CPU tensors (no kernels involved): bit-equal
    from genhancer_b200.clip_models.sampling import make_img_ids, patchify
    assert torch.equal(patchify(fp["latent"]), ref["img"])
    assert torch.equal(make_img_ids(B, h // 2, w // 2, "cpu"), ref["img_ids"])
