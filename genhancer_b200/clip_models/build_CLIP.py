"""``load_clip_model_*`` -- drop-ins for /root/reference/Continuous/clip_models/build_CLIP.py:5-29."""
import torch

from .CLIP_bank import MetaCLIP, OpenAICLIP, SigLIP


def _load(cls, config, device):
    class_model = cls(config)
    class_model.to(device)
    class_model.to(torch.float32)  # parameters stay fp32 as in the reference; kernels use cached bf16 operands
    return class_model


def load_clip_model_OpenAICLIP(config, device):
    return _load(OpenAICLIP, config, device)


def load_clip_model_SigLIP(config, device):
    return _load(SigLIP, config, device)


def load_clip_model_MetaCLIP(config, device):
    return _load(MetaCLIP, config, device)
