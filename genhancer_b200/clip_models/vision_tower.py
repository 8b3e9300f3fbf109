"""CLIP / SigLIP vision towers on the sm_100a kernels, with HF-compatible module trees and state_dict keys.

The reference does not contain the tower: it calls HF ``transformers`` (pinned 4.43.3) --
``CLIPModel`` / ``SiglipModel`` ``.vision_model`` (modeling_clip.py:138-218,261-385,647-696;
modeling_siglip.py:116-187,252-362,586-654).  These classes reproduce that arithmetic:

  patch conv (im2col gather + tcgen05 GEMM) -> [cls] + pos -> [pre_layrnorm] -> N x {LN1 -> fused QKV GEMM ->
  flash attention straight off the QKV buffer -> out_proj GEMM (+residual) -> LN2 -> fc1 GEMM (+act) ->
  fc2 GEMM (+residual)} -> post_layernorm / MAP head.

Stage 1 runs the tower frozen (no activations kept).  Compute is bf16 with fp32 accumulation; parameters stay
fp32 in the module (as the reference's ``class_model.to(torch.float32)``, build_CLIP.py:9) and bf16 operand
copies are cached until the parameters change.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from types import SimpleNamespace

import torch
from torch import nn

from .. import kernels as K
from ..kernels import ACT_GELU_TANH, ACT_QUICK_GELU, BF16, F32


@dataclass
class TowerConfig:
    kind: str = "clip"          # "clip" (OpenAI / MetaCLIP) | "siglip"
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    image_size: int = 224
    patch_size: int = 14
    projection_dim: int = 768
    layer_norm_eps: float = 1e-5
    hidden_act: str = "quick_gelu"

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_tokens(self) -> int:
        return self.grid ** 2 + (1 if self.kind == "clip" else 0)


def openai_vit_l14(image_size: int) -> TowerConfig:
    return TowerConfig("clip", 1024, 24, 16, 4096, image_size, 14, 768, 1e-5, "quick_gelu")


def metaclip_h14(image_size: int = 224) -> TowerConfig:
    return TowerConfig("clip", 1280, 32, 16, 5120, image_size, 14, 1024, 1e-5, "quick_gelu")


def siglip_so400m(image_size: int) -> TowerConfig:
    return TowerConfig("siglip", 1152, 27, 16, 4304, image_size, 14, 1152, 1e-6, "gelu_pytorch_tanh")


def _pad_heads_out(t: torch.Tensor, n_heads: int, d: int, dp: int) -> torch.Tensor:
    """Rows (or entries) grouped as n_heads blocks of d -> blocks of dp with zero rows appended per head: the output
    features of a q/k/v projection laid out for the attention kernel's head_dim (72 / 80 -> 128 with zero lanes,
    SURVEY.md R3; exact: zero q/k lanes add nothing to the scores, zero v lanes give zero outputs)."""
    if d == dp:
        return t.detach()
    t = t.detach()
    shp = t.shape
    out = torch.zeros((n_heads, dp) + tuple(shp[1:]), dtype=t.dtype, device=t.device)
    out[:, :d] = t.reshape((n_heads, d) + tuple(shp[1:]))
    return out.reshape((n_heads * dp,) + tuple(shp[1:]))


def _pad_heads_in(w: torch.Tensor, n_heads: int, d: int, dp: int) -> torch.Tensor:
    """[N, n_heads*d] -> [N, n_heads*dp]: the input features of the out-projection, zero columns for the pad lanes."""
    if d == dp:
        return w.detach()
    w = w.detach()
    out = torch.zeros(w.shape[0], n_heads, dp, dtype=w.dtype, device=w.device)
    out[:, :, :d] = w.reshape(w.shape[0], n_heads, d)
    return out.reshape(w.shape[0], n_heads * dp)


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))


class _MLP(nn.Module):
    def __init__(self, d, m):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, m), nn.Linear(m, d)


class _Layer(nn.Module):
    def __init__(self, c: TowerConfig):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.self_attn = _Attn(c.hidden_size)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.mlp = _MLP(c.hidden_size, c.intermediate_size)


class _Encoder(nn.Module):
    def __init__(self, c: TowerConfig):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class _Embeddings(nn.Module):
    def __init__(self, c: TowerConfig):
        super().__init__()
        if c.kind == "clip":
            self.class_embedding = nn.Parameter(torch.randn(c.hidden_size))
        self.patch_embedding = nn.Conv2d(3, c.hidden_size, c.patch_size, c.patch_size, bias=(c.kind == "siglip"))
        self.position_embedding = nn.Embedding(c.num_tokens, c.hidden_size)


class _MapHead(nn.Module):  # SiglipMultiheadAttentionPoolingHead parameter tree
    def __init__(self, c: TowerConfig):
        super().__init__()
        self.probe = nn.Parameter(torch.randn(1, 1, c.hidden_size))
        self.attention = nn.MultiheadAttention(c.hidden_size, c.num_attention_heads, batch_first=True)
        self.layernorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.mlp = _MLP(c.hidden_size, c.intermediate_size)


class VisionTransformer(nn.Module):
    """`.vision_model` of the HF model: call -> object with .last_hidden_state [B,T,D] and .pooler_output [B,D]."""

    def __init__(self, c: TowerConfig):
        super().__init__()
        self.config = c
        self.embeddings = _Embeddings(c)
        if c.kind == "clip":
            self.pre_layrnorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)  # (sic) HF spelling
        self.encoder = _Encoder(c)
        self.post_layernorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        if c.kind == "siglip":
            self.head = _MapHead(c)
        self._init_weights()

    @property
    def head_dim_padded(self) -> int:
        d = self.config.hidden_size // self.config.num_attention_heads
        if d in (64, 128):
            return d
        if d < 128 and d % 8 == 0:
            return 128
        raise NotImplementedError(f"head_dim {d}: the sm_100a attention kernels take 64 or 128 (72 / 80 are zero-padded)")

    def _init_weights(self):  # HF _init_weights flavour (modeling_clip.py:403-459): small normal, unit LayerNorm
        c = self.config
        std = c.hidden_size ** -0.5
        with torch.no_grad():
            for n, p in self.named_parameters():
                if "norm" in n:
                    p.fill_(1.0) if n.endswith("weight") else p.zero_()
                elif n.endswith("bias"):
                    p.zero_()
                elif "class_embedding" in n or "probe" in n:
                    p.normal_(0, std)
                elif "position_embedding" in n or "patch_embedding" in n:
                    p.normal_(0, 0.02)
                else:
                    p.normal_(0, std * (2 * c.num_hidden_layers) ** -0.5 if ("out_proj" in n or "fc2" in n) else std)

    def forward(self, pixel_values, output_hidden_states: bool = False, _norm=None, **_):
        """pixel_values: normalised images [B,3,S,S] (any float dtype).  `_norm=(mean3, std3)` lets the fused
        training step hand over raw [0,1] images and fold transforms.Normalize into the im2col gather.
        The schedule itself (frozen forward, or forward + saved activations + LoRA backward in stage 2) lives in
        ``tower_engine``."""
        from . import tower_engine
        self.head_dim_padded  # raises NotImplementedError for head dims the attention kernels cannot take
        owner = _OWNER.get(self)
        model = owner() if owner is not None else None
        if model is None:
            model = _Standalone(self)
        lhs, pooled = tower_engine.run_tower(model, pixel_values, _norm)
        return SimpleNamespace(last_hidden_state=lhs, pooler_output=pooled, hidden_states=None)


class _Standalone:
    """A bare ``VisionTransformer`` seen through the interface the engine expects of a ``VisionLanguageModel``."""

    def __init__(self, vm):
        self.vision_model = vm
        eng = _ENGINES.get(vm)
        if eng is not None:
            self._engine = eng

    def parameters(self):
        return self.vision_model.parameters()

    def __setattr__(self, k, v):
        object.__setattr__(self, k, v)
        if k == "_engine":
            _ENGINES[self.vision_model] = v


_OWNER: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()    # vision_model -> weakref(VisionLanguageModel)
_ENGINES: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()  # standalone vision_model -> TowerEngine


class VisionLanguageModel(nn.Module):
    """Stands where HF ``CLIPModel`` / ``SiglipModel`` stands in the reference wrappers (``wrapper.model``):
    `.vision_model`, `.visual_projection`, `.text_projection` (the 336 scripts re-wrap both projection weights as
    contiguous Parameters, train_OpenAICLIP_video_stage1.py:197-200).  The text tower is not on the path."""

    def __init__(self, c: TowerConfig):
        super().__init__()
        self.config = c
        self.vision_model = VisionTransformer(c)
        _OWNER[self.vision_model] = weakref.ref(self)
        if c.kind == "clip":
            self.visual_projection = nn.Linear(c.hidden_size, c.projection_dim, bias=False)
            self.text_projection = nn.Linear(8, c.projection_dim, bias=False)  # placeholder: text tower unused
            with torch.no_grad():
                self.visual_projection.weight.normal_(0, c.hidden_size ** -0.5)

    def get_image_features(self, pixel_values):
        out = self.vision_model(pixel_values)
        return project(self, out.pooler_output)

    @classmethod
    def from_pretrained(cls, path: str, c: TowerConfig):
        """Load the vision-side tensors of an HF checkpoint directory (pytorch_model.bin / model.safetensors)."""
        import os
        m = cls(c)
        sd = None
        for fn in ("model.safetensors", "pytorch_model.bin"):
            fp = os.path.join(path, fn)
            if os.path.exists(fp):
                if fn.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(fp)
                else:
                    sd = torch.load(fp, map_location="cpu", weights_only=True)
                break
        if sd is None:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        own = m.state_dict()
        picked = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
        missing = [k for k in own if k not in picked and not k.startswith("text_projection")]
        if missing:
            raise RuntimeError(f"checkpoint {path} lacks vision tensors: {missing[:5]} ...")
        m.load_state_dict(picked, strict=False)
        # Everything that is not on the vision path (text tower, logit_scale / logit_bias, the real text_projection,
        # persistent position_ids buffers) rides along untouched, so that the stage-2 export
        # (lora.save_pretrained = merge_and_unload().save_pretrained of train_SigLIP_stage2_all.py:305-311) writes the
        # FULL CLIPModel / SiglipModel the evaluation scripts load, not a vision-only file.
        m._passthrough_state = {k: v for k, v in sd.items() if k not in picked}
        m._from_checkpoint = path
        cj = os.path.join(path, "config.json")
        if os.path.exists(cj):
            import json
            with open(cj) as f:
                m._hf_config = json.load(f)
        return m


def project(model: VisionLanguageModel, pooled: torch.Tensor) -> torch.Tensor:
    """visual_projection (no bias) on the tcgen05 GEMM; frozen in stage 1, LoRA-wrapped under 'all-linear'."""
    from . import tower_engine
    return tower_engine.run_projection(model, pooled)
