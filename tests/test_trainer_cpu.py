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
