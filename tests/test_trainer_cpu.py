"""Host logic of the train_*.py entry points that needs no GPU: YAML -> config, entry-point inventory, checkpoint
file naming / resume discovery."""
import os

import torch

from conftest import ROOT
from genhancer_b200 import trainer

REFERENCE_ENTRY_POINTS = [  # every train script of the reference fork + the three OpenAI image scripts its docs name
    "train_MetaCLIP_stage1.py", "train_MetaCLIP_stage2_all.py", "train_MetaCLIP_stage2_only.py",
    "train_OpenAICLIP_nextpredic_stage1.py", "train_OpenAICLIP_nextpredic_stage2_all.py",
    "train_OpenAICLIP_sliding_windows_nextpredic_stage1.py", "train_OpenAICLIP_sliding_windows_nextpredic_stage2_all.py",
    "train_OpenAICLIP_use2frames_nextpredic_stage1.py", "train_OpenAICLIP_use2frames_nextpredic_stage2_all.py",
    "train_OpenAICLIP_video_stage1.py", "train_OpenAICLIP_video_stage2_all.py",
    "train_SigLIP_stage1.py", "train_SigLIP_stage2_all.py", "train_SigLIP_stage2_only.py",
    "train_OpenAICLIP_stage1.py", "train_OpenAICLIP_stage2_all.py", "train_OpenAICLIP_stage2_only.py"]


def test_every_reference_entry_point_exists():
    for f in REFERENCE_ENTRY_POINTS:
        src = open(os.path.join(ROOT, f)).read()
        assert "from genhancer_b200.trainer import main" in src and "--config" in src


def test_yaml_config_schema_and_types():
    for name in ("test_OpenAICLIP_336_stage1.yaml", "test_OpenAICLIP_336_video_stage1.yaml",
                 "test_OpenAICLIP_224_stage1_sliding_window.yaml"):
        c = trainer.load_config(os.path.join(ROOT, "train_configs", name))
        assert c.model_name == "flux-dev"
        assert isinstance(c.learning_rate, float) and c.learning_rate == 1e-4      # "1e-4" must not stay a string
        assert isinstance(c.adam_epsilon, float) and c.adam_epsilon == 1e-8
        assert c.clip_config.clip_dim == 768 and c.clip_config.t5_dim == 4096
        assert c.data_config.train_batch_size > 0 and c.get("nope") is None
    assert c.window_cond == 3 and c.max_windows_per_video == 8
    assert trainer.parse_args(["--config", "x.yaml"]) == "x.yaml"


def test_checkpoint_layout_and_resume_discovery(tmp_path):
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 4)

    class Clip(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.project_clip, self.project_t5 = M(), M()

    class Opt:
        def state_dict(self):
            return {"state": {}, "param_groups": []}

        def load_state_dict(self, sd):
            self.loaded = sd

    out = str(tmp_path)
    assert trainer.latest_step(out) is None
    dit, clip, ad, opt = M(), Clip(), M(), Opt()
    trainer.save_checkpoint(out, 50, dit, clip, None, opt, video=False)
    trainer.save_checkpoint(out, 313, dit, clip, ad, opt, video=True, save_project_clip=False)
    files = sorted(os.listdir(out))
    assert files == ["checkpoint-dit-313.bin", "checkpoint-dit-50.bin", "checkpoint-project-clip-50.bin",
                     "checkpoint-project-t5-50.bin", "checkpoint-visual-adapter-313.bin", "optimizer-state-313.bin",
                     "optimizer-state-50.bin"]
    assert trainer.latest_step(out) == 313
    sd = torch.load(os.path.join(out, "checkpoint-dit-50.bin"), weights_only=True)
    assert sorted(sd) == ["lin.bias", "lin.weight"]
    dit2 = M()
    trainer.load_checkpoint(out, 50, dit2, Clip(), None, opt, video=False)
    assert torch.equal(dit2.lin.weight, dit.lin.weight) and opt.loaded == {"state": {}, "param_groups": []}


def test_state_dict_keys_match_the_reference_checkpoint_contract():
    """checkpoint-dit-N.bin: 100 tensors with the names SURVEY.md 8b lists; AE encoder and tower keys likewise."""
    from genhancer_b200.flux.util import configs
    from genhancer_b200.flux.model import Flux
    from oracle import genhancer_oracle as O
    with torch.device("meta"):
        dit = Flux(configs["flux-dev"].params)
    sd = dit.state_dict()
    ks = O.flux_key_shapes(O.FluxCfg())
    assert len(sd) == 100 and set(sd) == set(ks)
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in ks.items())
    assert tuple(sd["single_blocks.3.linear1.weight"].shape) == (21504, 3072)
    assert tuple(sd["double_blocks.1.img_mod.lin.weight"].shape) == (18432, 3072)
    from genhancer_b200.clip_models import vision_tower as vt
    with torch.device("meta"):
        tower = vt.VisionLanguageModel(vt.openai_vit_l14(336))
    tk = O.tower_key_shapes(O.openai_vit_l14(336))
    mine = {k: tuple(v.shape) for k, v in tower.state_dict().items() if not k.startswith("text_projection")}
    assert mine == {k: tuple(v) for k, v in tk.items()}


def test_stage2_export_round_trips_a_full_hf_checkpoint(tmp_path):
    """the merged stage-2 export must be the FULL CLIPModel (text tower, logit_scale, the real
    text_projection, config.json with text_config), as ``merge_and_unload().save_pretrained`` writes it
    (train_SigLIP_stage2_all.py:305-311) -- evaluation loads the directory with CLIPModel.from_pretrained."""
    import json
    import pytest
    from transformers import CLIPConfig, CLIPModel
    from genhancer_b200.clip_models import lora, vision_tower as vt
    vis = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=1, image_size=28,
               patch_size=14, hidden_act="quick_gelu", projection_dim=32)
    txt = dict(hidden_size=48, intermediate_size=96, num_hidden_layers=1, num_attention_heads=2, projection_dim=32,
               vocab_size=50, max_position_embeddings=8)
    torch.manual_seed(0)
    hf = CLIPModel(CLIPConfig(text_config=txt, vision_config=vis, projection_dim=32)).eval()
    src = tmp_path / "hf_src"
    src.mkdir()
    hf_sd = {k: v.clone() for k, v in hf.state_dict().items()}
    torch.save(hf_sd, src / "pytorch_model.bin")
    (src / "config.json").write_text(json.dumps(hf.config.to_dict(), default=str))
    cfg = vt.TowerConfig("clip", 64, 2, 1, 128, 28, 14, 32, 1e-5, "quick_gelu")
    m = vt.VisionLanguageModel.from_pretrained(str(src), cfg)
    assert any(k.startswith("text_model.") for k in m._passthrough_state) and "logit_scale" in m._passthrough_state
    assert tuple(m._passthrough_state["text_projection.weight"].shape) == (32, 48)
    m = lora.get_peft_model(m, lora.LoraConfig(r=16, lora_alpha=16, target_modules="all-linear", bias="lora_only"))
    with torch.no_grad():
        for pair in m.lora.values():
            pair.B.normal_(0, 0.05)
    out = tmp_path / "merged"
    lora.save_pretrained(m, str(out))
    saved = torch.load(out / "pytorch_model.bin", weights_only=True)
    assert set(saved) == set(hf_sd), set(saved) ^ set(hf_sd)                      # no key lost, none invented
    for k, v in hf_sd.items():
        if not (k.startswith("vision_model.") or k.startswith("visual_projection")):
            assert torch.equal(saved[k], v), k                                   # text side: untouched
    k = "vision_model.encoder.layers.1.mlp.fc1.weight"
    pair = m.lora["vision_model/encoder/layers/1/mlp/fc1"]
    assert torch.allclose(saved[k], hf_sd[k] + pair.B @ pair.A, atol=1e-6) and not torch.equal(saved[k], hf_sd[k])
    assert json.loads((out / "config.json").read_text())["text_config"]["hidden_size"] == 48
    # what evaluation does: the directory loads as a complete CLIPModel, text features included
    re = CLIPModel.from_pretrained(str(out)).eval()
    ids = torch.tensor([[1, 5, 7, 2]])
    def text_features(model):   # (a tensor in transformers 4.x, an output object in 5.x)
        o = model.get_text_features(input_ids=ids)
        return o if isinstance(o, torch.Tensor) else o.pooler_output
    with torch.no_grad():
        assert torch.equal(text_features(re), text_features(hf))
    # a checkpoint-loaded tower that lost its passthrough tensors must refuse to export a vision-only file
    m._passthrough_state = None
    with pytest.raises(RuntimeError, match="text tower"):
        lora.save_pretrained(m, str(tmp_path / "bad"))
