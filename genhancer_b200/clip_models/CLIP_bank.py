"""``OpenAICLIP`` / ``SigLIP`` / ``MetaCLIP`` wrappers -- drop-ins for
/root/reference/Continuous/clip_models/CLIP_bank.py:8-122 (same constructor argument, attributes ``model``,
``project_clip``, ``project_t5``, same forward return triple, same state_dict keys for the projectors).

Differences that are deliberate:
  * the HF model object is replaced by ``vision_tower.VisionLanguageModel`` (sm_100a kernels, HF key names);
  * checkpoint locations are not hard-coded (the reference hard-codes the authors' absolute paths,
    CLIP_bank.py:15,48-50,81,97): they come from ``config.clip_path`` or the environment variables below;
    without one the tower is random-initialised (what BASELINE.json's configs ask for) and a warning is printed.
"""
from __future__ import annotations

import os
import warnings

import torch.nn as nn

from .. import ops
from ..kernels import ACT_GELU_ERF
from . import vision_tower as vt


def _projector(in_dim: int, out_dim: int) -> nn.Sequential:  # CLIP_bank.py:17-28
    return nn.Sequential(nn.LayerNorm(in_dim), nn.Linear(in_dim, out_dim), nn.GELU(), nn.Linear(out_dim, out_dim))


def run_projector(seq: nn.Sequential, x):
    """LayerNorm -> Linear -> GELU(erf) -> Linear on the sm_100a kernels (autograd-aware)."""
    h = ops.layer_norm(x, seq[0].weight, seq[0].bias, seq[0].eps)
    h = ops.linear(h, seq[1].weight, seq[1].bias, act=ACT_GELU_ERF)
    return ops.linear(h, seq[3].weight, seq[3].bias)


def _load_tower(cfg: vt.TowerConfig, config, env: str) -> vt.VisionLanguageModel:
    path = getattr(config, "clip_path", None) or os.environ.get(env)
    if path:
        return vt.VisionLanguageModel.from_pretrained(path, cfg)
    warnings.warn(f"no checkpoint given for the vision tower (set config.clip_path or ${env}); using random init")
    return vt.VisionLanguageModel(cfg)


class _Wrapper(nn.Module):
    feat_dim: int

    def _finish(self, model, config, feat_dim):
        self.project_clip = _projector(feat_dim, config.clip_dim)
        self.project_t5 = _projector(feat_dim, config.t5_dim)
        self.model = model
        self.config = config

    def class_token(self, images, _norm=None):
        out = self.model.vision_model(images, _norm=_norm)
        if self.model.config.kind == "clip":
            return vt.project(self.model, out.pooler_output)
        return out.pooler_output

    def project(self, class_token):
        """-> (projection_clip, projection_t5[:, None, :]): the two trainable MLP projectors (CLIP_bank.py:36-39)."""
        return run_projector(self.project_clip, class_token), run_projector(self.project_t5, class_token[:, None, :])

    def forward(self, images, _norm=None):
        class_token = self.class_token(images, _norm)
        projection_clip, projection_t5 = self.project(class_token)
        return class_token, projection_clip, projection_t5


class OpenAICLIP(_Wrapper):  # CLIP_bank.py:8-40
    def __init__(self, config):
        super().__init__()
        if config.clip_image_size not in (224, 336):
            raise ValueError("OpenAICLIP: clip_image_size must be 224 or 336")
        cfg = vt.openai_vit_l14(config.clip_image_size)
        self._finish(_load_tower(cfg, config, f"GENHANCER_OPENAI_CLIP_{config.clip_image_size}"), config, 768)


class SigLIP(_Wrapper):  # CLIP_bank.py:43-73
    def __init__(self, config):
        super().__init__()
        if config.clip_image_size not in (224, 384):
            raise ValueError("SigLIP: clip_image_size must be 224 or 384")
        cfg = vt.siglip_so400m(config.clip_image_size)
        self._finish(_load_tower(cfg, config, f"GENHANCER_SIGLIP_{config.clip_image_size}"), config, 1152)


class MetaCLIP(_Wrapper):  # CLIP_bank.py:76-122
    def __init__(self, config):
        super().__init__()
        if config.clip_type == "large":
            cfg, feat = vt.openai_vit_l14(getattr(config, "clip_image_size", 224)), 768
        elif config.clip_type == "huge":
            cfg, feat = vt.metaclip_h14(getattr(config, "clip_image_size", 224)), 1024
        else:
            raise ValueError("MetaCLIP: clip_type must be 'large' or 'huge'")
        self._finish(_load_tower(cfg, config, f"GENHANCER_METACLIP_{config.clip_type.upper()}"), config, feat)
