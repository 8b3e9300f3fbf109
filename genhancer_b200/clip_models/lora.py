"""LoRA on the vision tower (stage 2) -- what the reference gets from ``peft`` (0.14.0, not installed here):

    lora_config = LoraConfig(r, lora_alpha, target_modules=[...] | "all-linear", lora_dropout, bias)
    clip_vis.model = get_peft_model(clip_vis.model, lora_config)
    ... merge_and_unload().save_pretrained(dir, safe_serialization=False)

(/root/reference/Continuous/train_SigLIP_stage2_all.py:134-142,305-311;
 train_OpenAICLIP_use2frames_nextpredic_stage2_all.py:174-182).

Restated semantics (peft's published LoRA layer): for every targeted ``nn.Linear``
``y = W x + b + (lora_alpha / r) * B(A(dropout(x)))``; ``A`` [r, in] ~ kaiming_uniform(a=sqrt(5)), ``B`` [out, r] = 0;
the base model is frozen; ``bias="lora_only"`` leaves the biases of the wrapped layers trainable; merging is
``W += (lora_alpha / r) * B @ A``.  The arithmetic itself runs in ``tower_engine`` with the rank-r branch folded into
the base GEMM (a second operand pair: one extra MMA k-step), not as two extra GEMMs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn

SIGLIP_TARGETS = ("k_proj", "v_proj", "q_proj", "out_proj", "fc1", "fc2")


@dataclass
class LoraConfig:
    r: int = 16
    lora_alpha: int = 16
    target_modules: tuple | str = SIGLIP_TARGETS
    lora_dropout: float = 0.0
    bias: str = "none"  # "none" | "lora_only" | "all"

    @property
    def scaling(self) -> float:
        return self.lora_alpha / self.r


class LoraPair(nn.Module):
    def __init__(self, in_features: int, out_features: int, r: int):
        super().__init__()
        self.A = nn.Parameter(torch.empty(r, in_features))
        self.B = nn.Parameter(torch.zeros(out_features, r))
        nn.init.kaiming_uniform_(self.A, a=math.sqrt(5))


def _targets(model, cfg: LoraConfig):
    """(dotted name, nn.Linear) of every wrapped layer of the VISION side (the text tower is not on the path)."""
    out = []
    all_linear = cfg.target_modules == "all-linear"
    for name, m in model.named_modules():
        if not isinstance(m, nn.Linear) or name.startswith("text_projection"):
            continue
        leaf = name.rsplit(".", 1)[-1]
        if name.startswith("vision_model.head.attention"):
            continue  # inside nn.MultiheadAttention, whose forward reads .weight/.bias directly: a pair there never acts
        if all_linear or leaf in cfg.target_modules:
            out.append((name, m))
    return out


def get_peft_model(model, cfg: LoraConfig):
    """Freeze ``model`` (a vision_tower.VisionLanguageModel) and attach LoRA pairs; returns the same object, like
    ``clip_vis.model = get_peft_model(clip_vis.model, lora_config)`` in the reference."""
    if cfg.r % 16 != 0:
        raise ValueError("the fused LoRA GEMM path takes r as a multiple of 16 (the reference uses r = 16)")
    model.requires_grad_(False)
    model.lora_config = cfg
    model.lora = nn.ModuleDict()
    dev = next(model.parameters()).device
    for name, lin in _targets(model, cfg):
        model.lora[name.replace(".", "/")] = LoraPair(lin.in_features, lin.out_features, cfg.r).to(dev)
        if cfg.bias in ("lora_only", "all") and lin.bias is not None:
            lin.bias.requires_grad_(True)
    head = getattr(model.vision_model, "head", None)
    if head is not None and cfg.bias in ("lora_only", "all") and (cfg.target_modules == "all-linear" or "out_proj" in cfg.target_modules):
        head.attention.out_proj.bias.requires_grad_(True)  # peft wraps that module too, so its bias trains
    if cfg.bias == "all":
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.requires_grad_(True)
    return model


def lora_pair(model, dotted: str):
    lo = getattr(model, "lora", None)
    if lo is None:
        return None
    key = dotted.replace(".", "/")
    return lo[key] if key in lo else None


def print_trainable_parameters(model) -> None:
    tr = sum(p.numel() for p in model.parameters() if p.requires_grad)
    al = sum(p.numel() for p in model.parameters())
    print(f"trainable params: {tr:,d} || all params: {al:,d} || trainable%: {100 * tr / max(al, 1):.4f}")


@torch.no_grad()
def merged_state_dict(model) -> dict:
    """state_dict of the base model with every LoRA pair merged in (``merge_and_unload``), HF key names."""
    sd = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.startswith("lora.")}
    cfg = getattr(model, "lora_config", None)
    if cfg is not None:
        for key, pair in model.lora.items():
            name = key.replace("/", ".")
            sd[f"{name}.weight"] = sd[f"{name}.weight"] + cfg.scaling * (pair.B.float() @ pair.A.float()).to(sd[f"{name}.weight"].dtype)
    extra = getattr(model, "_passthrough_state", None)   # text tower etc. of a loaded HF checkpoint, untouched
    if getattr(model, "_from_checkpoint", None) and extra is None:
        raise RuntimeError("the tower was loaded from an HF checkpoint but its non-vision tensors were dropped: the merged "
                           "export would lack the text tower (evaluation loads it with CLIPModel / SiglipModel.from_pretrained)")
    # `text_projection` of this module is a placeholder (the text tower is not on the path): only a loaded one is saved
    sd = {k: v for k, v in sd.items() if not k.startswith("text_projection")}
    if extra:
        sd.update({k: v.detach().clone() for k, v in extra.items()})
    return sd


def save_pretrained(model, path: str) -> None:
    """``merge_and_unload().save_pretrained(path, safe_serialization=False)``: pytorch_model.bin + config.json."""
    import json
    import os
    c = model.config
    if getattr(model, "_from_checkpoint", None) and getattr(model, "_hf_config", None) is None:
        raise RuntimeError(f"no config.json was found next to the checkpoint {model._from_checkpoint}: cannot write a "
                           "loadable HF directory (text_config would be missing)")
    os.makedirs(path, exist_ok=True)
    torch.save({k: v.cpu() for k, v in merged_state_dict(model).items()}, os.path.join(path, "pytorch_model.bin"))
    cfg = getattr(model, "_hf_config", None) or {
        "model_type": "clip" if c.kind == "clip" else "siglip", "projection_dim": c.projection_dim,
        "vision_config": {"hidden_size": c.hidden_size, "intermediate_size": c.intermediate_size,
                          "num_hidden_layers": c.num_hidden_layers, "num_attention_heads": c.num_attention_heads,
                          "image_size": c.image_size, "patch_size": c.patch_size, "layer_norm_eps": c.layer_norm_eps,
                          "hidden_act": c.hidden_act, "projection_dim": c.projection_dim}}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=2)
