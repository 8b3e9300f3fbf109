"""End-to-end on the B200: the train_*.py host (genhancer_b200/trainer.py) runs real optimizer steps at full model
size (1.31 B-parameter DiT, ViT-L/14-224, full AE) on synthetic data, writes the reference's checkpoint files,
resumes from them, and the loss goes down."""
import math
import os
import shutil

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = """model_name: "flux-dev"
data_config:
  train_batch_size: 2
  num_workers: 0
  img_size: 224
  {dirkey}: synthetic
  seed: 0
  patch_size: 1
clip_config:
  clip_image_size: 224
  clip_dim: 768
  t5_dim: 4096
scale_factor: 1.0
output_dir: {out}
max_train_steps: {steps}
learning_rate: 1e-4
adam_beta1: 0.9
adam_beta2: 0.999
adam_weight_decay: 0.01
adam_epsilon: 1e-8
max_grad_norm: 1.0
checkpointing_steps: 2
resume_from_checkpoint: latest
gradient_accumulation_steps: {ga}
window_cond: 3
window_stride: 1
max_windows_per_video: 2
"""


def _run(tmp_path, family, mode, steps, ga, dirkey):
    from genhancer_b200 import trainer
    out = str(tmp_path / "out")
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(CFG.format(dirkey=dirkey, out=out, steps=steps, ga=ga))
    return trainer.main(family, mode, "stage1", argv=["--config", str(cfg)]), out


def test_image_stage1_trains_checkpoints_and_resumes(tmp_path):
    import warnings
    warnings.simplefilter("ignore")
    res, out = _run(tmp_path, "OpenAICLIP", "image", steps=4, ga=2, dirkey="img_dir")
    assert res.global_step == 4 and all(math.isfinite(l) for l in res.losses)
    files = set(os.listdir(out))
    for n in (2, 4):
        assert {f"checkpoint-dit-{n}.bin", f"checkpoint-project-clip-{n}.bin", f"checkpoint-project-t5-{n}.bin",
                f"optimizer-state-{n}.bin"} <= files
    sd = torch.load(os.path.join(out, "checkpoint-dit-4.bin"), weights_only=True)
    assert len(sd) == 100 and sd["img_in.weight"].dtype == torch.bfloat16
    p5 = torch.load(os.path.join(out, "checkpoint-project-t5-4.bin"), weights_only=True)
    assert sorted(p5) == ["0.bias", "0.weight", "1.bias", "1.weight", "3.bias", "3.weight"]
    osd = torch.load(os.path.join(out, "optimizer-state-4.bin"), weights_only=False)
    assert osd["param_groups"][0]["lr"] == 1e-4 and len(osd["state"]) == 100 + 12
    # weights moved between the two checkpoints
    sd2 = torch.load(os.path.join(out, "checkpoint-dit-2.bin"), weights_only=True)
    assert not torch.equal(sd2["final_layer.linear.weight"], sd["final_layer.linear.weight"])
    del res
    torch.cuda.empty_cache()
    # resume: picks up at step 4, runs to 6
    from genhancer_b200 import trainer
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(cfg.read_text().replace("max_train_steps: 4", "max_train_steps: 6"))
    res2 = trainer.main("OpenAICLIP", "image", "stage1", argv=["--config", str(cfg)])
    assert res2.global_step == 6 and "checkpoint-dit-6.bin" in os.listdir(out)
    assert res2.opt.step_count == 6
    shutil.rmtree(out, ignore_errors=True)      # (8 GB per checkpoint: the box's scratch disk is small)


def test_graph_trainer_and_eager_trainer_agree(tmp_path):
    """The entry points run the measured path (graph.PipelinedTrainStep: update of step n captured under the frozen
    forward of step n + 1); with ``cuda_graph: false`` they run the sequential eager loop.  Same config, same seed:
    same losses and the same weights after 4 optimizer steps (up to fp32-atomic ordering)."""
    import warnings
    from genhancer_b200 import trainer
    warnings.simplefilter("ignore")
    res = {}
    for tag, flag in (("graph", "true"), ("eager", "false")):
        out = str(tmp_path / f"out_{tag}")
        cfg = tmp_path / f"cfg_{tag}.yaml"
        cfg.write_text(CFG.format(dirkey="img_dir", out=out, steps=4, ga=1).replace("checkpointing_steps: 2", "checkpointing_steps: 100")
                       + f"cuda_graph: {flag}\n")
        r = trainer.main("OpenAICLIP", "image", "stage1", argv=["--config", str(cfg)])
        res[tag] = dict(losses=r.losses, w={k: v.detach().float().cpu() for k, v in r.dit.state_dict().items()
                                           if k in ("final_layer.linear.weight", "single_blocks.2.linear1.weight", "img_in.bias")},
                        p5=r.clip_vis.project_t5[3].weight.detach().float().cpu(),
                        m=r.opt.groups[0].exp_avg[:8_000_000].float().cpu(), graph_steps=r.graph_steps,
                        eager_steps=r.eager_steps, step_count=r.opt.step_count)
        del r
        torch.cuda.empty_cache()
    assert res["graph"]["graph_steps"] == 4 and res["graph"]["eager_steps"] == 0
    assert res["eager"]["graph_steps"] == 0 and res["eager"]["eager_steps"] == 4
    assert res["graph"]["step_count"] == res["eager"]["step_count"] == 4
    for a, b in zip(res["graph"]["losses"], res["eager"]["losses"]):
        assert abs(a - b) <= 2e-3 * abs(b), (res["graph"]["losses"], res["eager"]["losses"])
    from conftest import cosine
    for k in res["graph"]["w"]:
        assert cosine(res["graph"]["w"][k], res["eager"]["w"][k]) >= 0.99999, k
    assert cosine(res["graph"]["p5"], res["eager"]["p5"]) >= 0.99999
    # Adam's first moments carry the gradients of all 4 steps: a skipped, doubled or stale update would show here
    assert cosine(res["graph"]["m"], res["eager"]["m"]) >= 0.999
    shutil.rmtree(str(tmp_path), ignore_errors=True)


@pytest.mark.parametrize("mode", ["use2frames_nextpredic", "sliding_windows_nextpredic"])
def test_video_stage1_modes_train(tmp_path, mode):
    import warnings
    warnings.simplefilter("ignore")
    res, out = _run(tmp_path, "OpenAICLIP", mode, steps=2, ga=1, dirkey="video_dir")
    assert res.global_step == 2 and all(math.isfinite(l) for l in res.losses)
    files = set(os.listdir(out))
    assert {"checkpoint-dit-2.bin", "checkpoint-visual-adapter-2.bin", "optimizer-state-2.bin"} <= files
    assert "checkpoint-project-t5-2.bin" not in files
    ad = torch.load(os.path.join(out, "checkpoint-visual-adapter-2.bin"), weights_only=True)
    assert tuple(ad["proj.0.weight"].shape) == (2048, 1024) and tuple(ad["proj.2.weight"].shape) == (4096, 2048)
    shutil.rmtree(out, ignore_errors=True)


LORA = """lora_config:
  r: 16
  lora_alpha: 16
  lora_dropout: 0.1
  bias: "lora_only"
load_dir: {load}
load_step: {load_step}
"""


def _run2(tmp_path, family, mode, stage, steps, dirkey, load="none", load_step=0, extra=""):
    from genhancer_b200 import trainer
    out = str(tmp_path / f"out_{stage}")
    cfg = tmp_path / f"cfg_{stage}.yaml"
    cfg.write_text(CFG.format(dirkey=dirkey, out=out, steps=steps, ga=1) + LORA.format(load=load, load_step=load_step) + extra)
    return trainer.main(family, mode, stage, argv=["--config", str(cfg)]), out


def test_image_stage2_all_and_only_train_lora_and_export_merged_tower(tmp_path):
    """train_SigLIP_stage2_{all,only}.py: LoRA (r16, dropout 0.1, bias lora_only) on the tower; stage 1 -> stage 2
    hand-over through the flat checkpoint files; output = HF directory with LoRA merged (pytorch_model.bin)."""
    import warnings
    warnings.simplefilter("ignore")
    res1, out1 = _run(tmp_path, "SigLIP", "image", steps=2, ga=1, dirkey="img_dir")
    del res1
    torch.cuda.empty_cache()
    res, out = _run2(tmp_path, "SigLIP", "image", "stage2_all", 2, "img_dir", load=out1, load_step=2)
    assert res.global_step == 2 and all(math.isfinite(l) for l in res.losses)
    d = os.path.join(out, "siglip-so400m-patch14-224-2")
    assert sorted(os.listdir(d)) == ["config.json", "pytorch_model.bin"]
    sd = torch.load(os.path.join(d, "pytorch_model.bin"), weights_only=True)
    assert not any("lora" in k for k in sd)
    k = "vision_model.encoder.layers.3.self_attn.q_proj.weight"
    base = res.clip_vis.model.state_dict()[k].cpu()
    assert tuple(sd[k].shape) == (1152, 1152) and not torch.equal(sd[k], base)     # B moved off zero -> merge changed W
    pair = res.clip_vis.model.lora["vision_model/encoder/layers/3/self_attn/q_proj"]
    assert pair.B.abs().max().item() > 0 and torch.isfinite(pair.A).all()
    assert not torch.equal(sd["vision_model.encoder.layers.3.mlp.fc1.bias"],
                           torch.zeros_like(sd["vision_model.encoder.layers.3.mlp.fc1.bias"]))
    n_lora = sum(p.numel() for p in res.clip_vis.model.lora.parameters())
    assert n_lora == 27 * 16 * (8 * 1152 + 2 * (1152 + 4304)) + 16 * 2 * (1152 + 4304)
    dit_w = res.dit.final_layer.linear.weight.detach().clone()
    del res
    torch.cuda.empty_cache()
    res, out = _run2(tmp_path, "SigLIP", "image", "stage2_only", 2, "img_dir", load=out1, load_step=2)
    assert res.global_step == 2 and all(math.isfinite(l) for l in res.losses)
    assert not any(p.requires_grad for p in res.dit.parameters())
    ref = torch.load(os.path.join(out1, "checkpoint-dit-2.bin"), weights_only=True)["final_layer.linear.weight"]
    assert torch.equal(res.dit.final_layer.linear.weight.detach().cpu(), ref)       # DiT frozen in stage2_only
    assert not torch.equal(dit_w.cpu(), ref)                                         # ... and trained in stage2_all
    assert os.path.isdir(os.path.join(out, "siglip-so400m-patch14-224-2"))
    assert "checkpoint-tower-lora-2.bin" in os.listdir(out) and "checkpoint-dit-2.bin" not in os.listdir(out)
    shutil.rmtree(str(tmp_path), ignore_errors=True)


def test_video_stage2_all_trains_tower_lora_through_the_adapter(tmp_path):
    """train_OpenAICLIP_use2frames_nextpredic_stage2_all.py: target_modules='all-linear', gradients reach the LoRA
    pairs through last_hidden_state -> VisualPromptAdapter -> DiT txt stream and through visual_projection."""
    import warnings
    warnings.simplefilter("ignore")
    res, out = _run2(tmp_path, "OpenAICLIP", "use2frames_nextpredic", "stage2_all", 2, "video_dir")
    assert res.global_step == 2 and all(math.isfinite(l) for l in res.losses)
    files = set(os.listdir(out))
    assert {"checkpoint-dit-2.bin", "checkpoint-visual-adapter-2.bin", "checkpoint-project-clip-2.bin",
            "optimizer-state-2.bin", "clip-vit-large-patch14-2"} <= files
    lo = res.clip_vis.model.lora
    assert "visual_projection" in lo and lo["visual_projection"].B.abs().max().item() > 0
    assert lo["vision_model/encoder/layers/0/mlp/fc2"].B.abs().max().item() > 0
    # stage-2 resume is real: the un-merged LoRA pairs + trainable biases + optimizer state come back
    b_before = lo["vision_model/encoder/layers/0/mlp/fc2"].B.detach().cpu().clone()
    assert "checkpoint-tower-lora-2.bin" in files
    del res, lo
    torch.cuda.empty_cache()
    from genhancer_b200 import trainer
    cfg = tmp_path / "cfg_stage2_all.yaml"
    res2 = trainer.main("OpenAICLIP", "use2frames_nextpredic", "stage2_all", argv=["--config", str(cfg)])   # nothing left to do:
    assert res2.global_step == 2 and res2.opt.step_count == 2                                # what was loaded is what we see
    b_after = res2.clip_vis.model.lora["vision_model/encoder/layers/0/mlp/fc2"].B.detach().cpu()
    assert torch.equal(b_after, b_before), "resume did not restore the un-merged LoRA pairs"
    m = res2.opt.groups[0].exp_avg
    assert m.abs().max().item() > 0, "resume did not restore the optimizer moments"
    shutil.rmtree(str(tmp_path), ignore_errors=True)
