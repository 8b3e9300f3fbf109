"""GPU parity of the video-mode step (per-patch conditioning through the VisualPromptAdapter, spatio-temporal RoPE
ids) against the reference's own pieces (fixture: oracle/make_golden.py::golden_video_small -- the adapter and id
helpers are lifted out of the reference train scripts by AST)."""
import pytest
import torch

from conftest import cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def _build(fx):
    from genhancer_b200.clip_models import CLIP_bank, vision_tower as vt
    from genhancer_b200.flux.model import Flux, FluxParams
    from genhancer_b200.flux.modules.autoencoder import AutoEncoder, AutoEncoderParams
    from genhancer_b200.video import SuperModel, VideoStep
    from oracle import genhancer_oracle as O
    c, ks, seed = fx["tower_cfg"], fx["key_shapes"], fx["seed"]
    cfg = vt.TowerConfig(c["kind"], c["hidden"], c["layers"], c["heads"], c["mlp"], c["image_size"], c["patch"],
                         c["proj_dim"], c["eps"], "quick_gelu")
    model = vt.VisionLanguageModel(cfg)
    model.load_state_dict(O.synth_state_dict(ks["tower"], seed), strict=False)

    class Cfg:
        clip_dim, t5_dim = 32, 48
    wrap = CLIP_bank._Wrapper()
    wrap._finish(model, Cfg, c["proj_dim"])
    wrap.requires_grad_(False)
    fc = dict(fx["flux_cfg"])
    fc["axes_dim"] = list(fc["axes_dim"])
    dit = Flux(FluxParams(**fc))
    dit.load_state_dict(O.synth_state_dict(ks["dit"], seed + 2), strict=True)
    sm = SuperModel(wrap, dit, adapter_in_dim=c["hidden"], adapter_out_dim=fc["context_in_dim"])
    sm.visual_adapter.load_state_dict(O.synth_state_dict(ks["adapter"], seed + 4), strict=True)
    sm = sm.to("cuda")
    sm.clip_vis.float()
    sm.visual_adapter.float()
    sm.dit.to(torch.bfloat16)
    ac = dict(fx["ae_cfg"])
    ac["ch_mult"] = list(ac["ch_mult"])
    ae = AutoEncoder(AutoEncoderParams(**ac))
    ae.encoder.load_state_dict(O.synth_state_dict(ks["ae"], seed + 3), strict=True)
    ae = ae.to("cuda").requires_grad_(False)
    return VideoStep(sm, ae, cond_times=fx["cond_times"], target_time=fx["target_time"]), sm


def test_video_step_matches_reference():
    fx = load_golden("video_step_small.pt")
    step, sm = _build(fx)
    dev = "cuda"
    fr = fx["frames"].to(dev)
    loss, parts = step([fr[:, 0], fr[:, 1]], fr[:, 2], ae_noise=fx["ae_noise"].to(dev), t=fx["t"].to(dev),
                       x_0=fx["x_0"].to(dev), return_parts=True)
    assert torch.equal(parts["txt_ids"].cpu(), fx["txt_ids"]) and torch.equal(parts["img_ids"].cpu(), fx["img_ids"])
    assert cosine(parts["txt"], fx["txt"]) >= 0.999 and cosine(parts["vec"], fx["vec"]) >= 0.999
    assert rel_err(parts["x_1"], fx["x_1"]) < 3e-2
    assert rel_err(parts["pred"], fx["pred"]) < 5e-2
    assert abs(loss.item() - fx["loss"].item()) / fx["loss"].item() < 1e-2
    loss.backward()
    cs = dict(proj2=cosine(sm.visual_adapter.proj[2].weight.grad, fx["grad_adapter_proj2_weight"]),
              proj3_bias=cosine(sm.visual_adapter.proj[3].bias.grad, fx["grad_adapter_proj3_bias"]),
              txt_in=cosine(dict(sm.dit.named_parameters())["txt_in.weight"].grad, fx["grad_txt_in_weight"]))
    # The fixture's gradients are the reference's fp32 run; the reference itself trains with a bf16 DiT
    # (train_OpenAICLIP_video_stage1.py: dit.to(bfloat16)).  Floor = what that bf16 run achieves against its own fp32
    # gradients on these exact tensors (oracle on the CPU, same weights / inputs / draws): the gate is SURVEY.md 8(d)'s
    # 0.99, or within 0.01 of the reference's own bf16 floor where bf16 arithmetic alone sits below it.
    from oracle import genhancer_oracle as O
    from test_step_gpu import OPENAI_MEAN, OPENAI_STD
    tc, fc, ac = O.TowerCfg(**fx["tower_cfg"]), O.FluxCfg(**fx["flux_cfg"]), O.AECfg(**fx["ae_cfg"])
    ks, seed = fx["key_shapes"], fx["seed"]
    sd_ad = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["adapter"], seed + 4).items()}
    sd_d = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(ks["dit"], seed + 2).items()}
    f = fx["frames"]
    out = O.stage1_video_step(O.synth_state_dict(ks["tower"], seed), sd_ad, sd_d, O.synth_state_dict(ks["ae"], seed + 3),
                              [f[:, 0], f[:, 1]], f[:, 2], tc, fc, ac, OPENAI_MEAN, OPENAI_STD, fx["cond_times"],
                              fx["target_time"], fx["ae_noise"], fx["t"], fx["x_0"], dit_dtype=torch.bfloat16)
    out.loss.backward()
    floor = dict(proj2=cosine(sd_ad["proj.2.weight"].grad, fx["grad_adapter_proj2_weight"]),
                 proj3_bias=cosine(sd_ad["proj.3.bias"].grad, fx["grad_adapter_proj3_bias"]),
                 txt_in=cosine(sd_d["txt_in.weight"].grad, fx["grad_txt_in_weight"]))
    print("GRADCOS video_step_small ours", cs, "reference-bf16 floor", floor)
    for k in cs:
        assert cs[k] >= min(0.99, floor[k] - 0.01), (k, cs[k], floor[k])


def test_video_step_argument_checks_and_adapter_keys():
    fx = load_golden("video_step_small.pt")
    step, sm = _build(fx)
    fr = fx["frames"].to("cuda")
    with pytest.raises(ValueError, match="conditioning frames"):
        step([fr[:, 0]], fr[:, 2])
    assert sorted(sm.visual_adapter.state_dict()) == ["proj.0.bias", "proj.0.weight", "proj.2.bias", "proj.2.weight",
                                                      "proj.3.bias", "proj.3.weight"]
