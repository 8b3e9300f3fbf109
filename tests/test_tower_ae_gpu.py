"""GPU parity of the vision tower (+ projectors, with gradients) and the FLUX AE encoder against golden vectors
produced by the reference's own modules (HF CLIPModel inside the reference's OpenAICLIP wrapper; the reference's
AutoEncoder) -- fixtures minted by oracle/make_golden.py."""
import pytest
import torch

from conftest import cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def _build_wrapper(fx):
    from genhancer_b200.clip_models import CLIP_bank, vision_tower as vt
    from oracle import genhancer_oracle as O
    c = fx["cfg"]
    cfg = vt.TowerConfig(c["kind"], c["hidden"], c["layers"], c["heads"], c["mlp"], c["image_size"], c["patch"],
                         c["proj_dim"], c["eps"], "quick_gelu" if c["act"] == "quick_gelu" else "gelu_pytorch_tanh")
    model = vt.VisionLanguageModel(cfg)
    sd_t = O.synth_state_dict(fx["key_shapes_tower"], fx["seed"])
    missing, unexpected = model.load_state_dict(sd_t, strict=False)
    assert not unexpected and all(k.startswith("text_projection") for k in missing), (missing, unexpected)

    class Cfg:
        clip_dim, t5_dim = fx["clip_dim"], fx["t5_dim"]
    wrap = CLIP_bank._Wrapper()
    wrap._finish(model, Cfg, c["proj_dim"] if c["kind"] == "clip" else c["hidden"])
    sd_w = O.synth_state_dict(fx["key_shapes_wrap"], fx["seed"] + 1)
    missing, unexpected = wrap.load_state_dict(sd_w, strict=False)
    assert not unexpected
    for n, p in wrap.named_parameters():
        p.requires_grad_("project_clip" in n or "project_t5" in n)
    return wrap.to("cuda").float()


def test_clip_tower_and_projectors_match_reference():
    fx = load_golden("tower_clip_small.pt")
    wrap = _build_wrapper(fx)
    img = fx["img"].to("cuda")
    mean = torch.tensor(OPENAI_MEAN, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD, device="cuda").view(1, 3, 1, 1)
    x = (img - mean) / std
    out = wrap.model.vision_model(x, output_hidden_states=True)
    # north_star: CLIP features at cosine >= 0.999
    assert cosine(out.last_hidden_state, fx["last_hidden_state"]) >= 0.999
    assert cosine(out.pooler_output, fx["pooler_output"]) >= 0.999
    assert rel_err(out.last_hidden_state, fx["last_hidden_state"]) < 2e-2
    cls, pc, pt5 = wrap(x)
    assert cls.shape == fx["class_token"].shape and pt5.shape == fx["projection_t5"].shape
    assert cosine(cls, fx["class_token"]) >= 0.999
    assert cosine(pc, fx["projection_clip"]) >= 0.999
    assert cosine(pt5, fx["projection_t5"]) >= 0.999
    (pc.float().square().mean() + pt5.float().square().mean()).backward()
    assert cosine(wrap.project_t5[1].weight.grad, fx["grad_project_t5_1_weight"]) >= 0.99
    assert cosine(wrap.project_clip[3].bias.grad, fx["grad_project_clip_3_bias"]) >= 0.99
    # folding transforms.Normalize into the im2col gather gives the same features as normalising first
    cls2, _, _ = wrap(img, _norm=(OPENAI_MEAN, OPENAI_STD))
    assert cosine(cls2, cls) >= 0.9999


def test_siglip_tower_map_head_and_projectors_match_reference():
    """SigLIP: biased patch conv, no CLS / pre-LN, tanh-GELU, eps 1e-6, post-LN on all tokens, MAP pooling head;
    head_dim 72 runs zero-padded to 128 lanes (fixture: HF SiglipModel inside the reference's SigLIP wrapper)."""
    fx = load_golden("tower_siglip_small.pt")
    assert fx["cfg"]["hidden"] // fx["cfg"]["heads"] == 72
    wrap = _build_wrapper(fx)
    x = ((fx["img"] - 0.5) / 0.5).to("cuda")
    out = wrap.model.vision_model(x, output_hidden_states=True)
    assert out.last_hidden_state.shape == fx["last_hidden_state"].shape
    assert cosine(out.last_hidden_state, fx["last_hidden_state"]) >= 0.999
    assert cosine(out.pooler_output, fx["pooler_output"]) >= 0.999
    cls, pc, pt5 = wrap(x)
    assert cosine(cls, fx["class_token"]) >= 0.999
    assert cosine(pc, fx["projection_clip"]) >= 0.999 and cosine(pt5, fx["projection_t5"]) >= 0.999
    (pc.float().square().mean() + pt5.float().square().mean()).backward()
    assert cosine(wrap.project_t5[1].weight.grad, fx["grad_project_t5_1_weight"]) >= 0.99
    cls2, _, _ = wrap(fx["img"].to("cuda"), _norm=((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)))
    assert cosine(cls2, cls) >= 0.9999


def test_tower_rejects_unsupported_head_dim_loudly():
    from genhancer_b200.clip_models import vision_tower as vt
    m = vt.VisionLanguageModel(vt.TowerConfig("clip", 72, 1, 2, 128, 28, 14, 32)).to("cuda")  # head_dim 36
    with pytest.raises(NotImplementedError):
        m.vision_model(torch.zeros(1, 3, 28, 28, device="cuda"))


def _build_ae(fx):
    from genhancer_b200.flux.modules.autoencoder import AutoEncoder, AutoEncoderParams
    from oracle import genhancer_oracle as O
    c = dict(fx["cfg"])
    c["ch_mult"] = list(c["ch_mult"])
    ae = AutoEncoder(AutoEncoderParams(**c))
    ae.encoder.load_state_dict(O.synth_state_dict(fx["key_shapes"], fx["seed"]), strict=True)
    return ae.to("cuda")


def test_ae_encoder_matches_reference():
    fx = load_golden("ae_small.pt")
    ae = _build_ae(fx)
    x = ((fx["img"] - 0.5) / 0.5).to("cuda")
    mom = ae.encoder(x)
    assert mom.shape == fx["moments"].shape
    assert cosine(mom, fx["moments"]) >= 0.999
    assert rel_err(mom, fx["moments"]) < 3e-2
    z = ae.encode(x, noise=fx["noise"].to("cuda"))
    assert z.shape == fx["z"].shape
    assert rel_err(z, fx["z"]) < 3e-2
    # fused form: raw image with the normalisation folded into the conv_in gather, patchified output
    from oracle import genhancer_oracle as O
    x1 = ae.encode_patchified(fx["img"].to("cuda"), 0.5, 0.5, noise=fx["noise"].to("cuda"))
    assert rel_err(x1, O.patchify(fx["z"])) < 3e-2
    assert torch.equal(O.patchify(z.cpu()), ae.encode_patchified(x, 0.0, 1.0, noise=fx["noise"].to("cuda")).cpu())


def test_ae_noise_draw_matches_reference_rng_order():
    """encode() without an explicit noise must consume the device RNG exactly like randn_like(mean)."""
    fx = load_golden("ae_small.pt")
    ae = _build_ae(fx)
    x = ((fx["img"] - 0.5) / 0.5).to("cuda")
    torch.manual_seed(5)
    z1 = ae.encode(x)
    after = torch.randn(3, device="cuda")
    torch.manual_seed(5)
    noise = torch.randn(2, 16, 6, 6, device="cuda")
    after_ref = torch.randn(3, device="cuda")
    z2 = ae.encode(x, noise=noise)
    assert torch.equal(z1, z2) and torch.equal(after, after_ref)


def test_ae_decoder_matches_reference_and_round_trip_shapes():
    """Decoder (autoencoder.py:183-259) + AutoEncoder.decode: unfolded conv_in, mid attention, three nearest-2x
    upsample + conv stages, 3-channel conv_out (zero-padded to 8 output channels for the kernel)."""
    from genhancer_b200.flux.modules.autoencoder import AutoEncoder, AutoEncoderParams
    from oracle import genhancer_oracle as O
    fx = load_golden("ae_decoder_small.pt")
    ac = dict(fx["cfg"])
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=ac["ch"], out_ch=3, ch_mult=list(ac["ch_mult"]),
                                       num_res_blocks=ac["num_res_blocks"], z_channels=ac["z_channels"],
                                       scale_factor=ac["scale_factor"], shift_factor=ac["shift_factor"]))
    ae.decoder.load_state_dict(O.synth_state_dict(fx["key_shapes"], fx["seed"]), strict=True)
    assert sorted(ae.decoder.state_dict()) == sorted(fx["key_shapes"])          # the reference's key names
    ae = ae.to("cuda")
    img = ae.decode(fx["z"].to("cuda"))
    assert img.shape == fx["image"].shape and img.dtype == torch.float32
    assert cosine(img, fx["image"]) >= 0.999
    assert rel_err(img, fx["image"]) < 3e-2
    # upsample kernel: exact
    from genhancer_b200 import kernels as K
    x = torch.randn(2, 5, 7, 64, device="cuda").to(torch.bfloat16)
    ref = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    assert torch.equal(K.upsample2x_nhwc(x), ref)
    # encode -> decode keeps the image geometry (random weights: only shapes / finiteness are meaningful)
    rec = ae.decode(ae.encode(torch.rand(1, 3, 64, 96, device="cuda") * 2 - 1))
    assert rec.shape == (1, 3, 64, 96) and torch.isfinite(rec).all()
    with pytest.raises(ValueError):
        ae.decoder(torch.zeros(1, 8, 4, 4, device="cuda"))
