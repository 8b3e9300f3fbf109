"""Flow-matching sampler -- drop-in for /root/reference/Continuous/src/flux/sampling.py (imported by every
``train_*.py``: ``from src.flux.sampling import denoise, get_noise, get_schedule, unpack``).

Same names, argument order and results: ``get_noise`` (:12-29), ``time_shift`` / ``get_lin_function`` /
``get_schedule`` (:66-94), ``denoise`` (:97-150, Euler steps with the true-CFG negative branch) and ``unpack``
(:234-242).  The DiT evaluations run on the fused sm_100a engine (``Flux.forward`` under ``no_grad`` keeps
nothing for backward); the Euler update and the guidance mix are one kernel (``gh_euler_cfg_step``).
``prepare`` (:32-63) needs the T5 / CLIP text ``HFEmbedder``s, which GenHancer never instantiates;
``denoise_controlnet`` (:152-231) drives a ControlNet that is not part of this repository either.
"""
from __future__ import annotations

import math
from typing import Callable

import torch
from torch import Tensor

from .. import kernels as K


def get_noise(num_samples: int, height: int, width: int, device: torch.device, dtype: torch.dtype, seed: int):
    return torch.randn(num_samples, 16, 2 * math.ceil(height / 16), 2 * math.ceil(width / 16), device=device, dtype=dtype,
                       generator=torch.Generator(device=device).manual_seed(seed))  # allow for packing


def prepare(t5, clip, img: Tensor, prompt):
    raise NotImplementedError("text conditioning (T5 / CLIP-text HFEmbedder) is never instantiated by GenHancer; the visual "
                              "conditioning comes from clip_models.sampling.prepare_clip")


def time_shift(mu: float, sigma: float, t: Tensor):
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_lin_function(x1: float = 256, y1: float = 0.5, x2: float = 4096, y2: float = 1.15) -> Callable[[float], float]:
    m = (y2 - y1) / (x2 - x1)
    b = y1 - m * x1
    return lambda x: m * x + b


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> list[float]:
    timesteps = torch.linspace(1, 0, num_steps + 1)  # extra step for zero
    if shift:  # favour high timesteps for higher-signal images: mu from a linear estimate between two points
        mu = get_lin_function(y1=base_shift, y2=max_shift)(image_seq_len)
        timesteps = time_shift(mu, 1.0, timesteps)
    return timesteps.tolist()


@torch.no_grad()
def denoise(model, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, vec: Tensor, neg_txt: Tensor,
            neg_txt_ids: Tensor, neg_vec: Tensor, timesteps: list[float], guidance: float = 4.0, true_gs=1,
            timestep_to_start_cfg=0, image_proj: Tensor = None, neg_image_proj: Tensor = None,
            ip_scale: Tensor | float = 1.0, neg_ip_scale: Tensor | float = 1.0):
    if image_proj is not None or neg_image_proj is not None:
        raise NotImplementedError("IP-adapter inputs are unused by every GenHancer script (dead processors, SURVEY.md Q11)")
    in_dtype = img.dtype
    x = img.to(torch.bfloat16).contiguous().clone()   # the engine computes in bf16 (the scripts run the DiT in bf16)
    guidance_vec = torch.full((x.shape[0],), guidance, device=x.device, dtype=x.dtype)
    i = 0
    for t_curr, t_prev in zip(timesteps[:-1], timesteps[1:]):
        t_vec = torch.full((x.shape[0],), t_curr, dtype=x.dtype, device=x.device)
        pred = model(img=x, img_ids=img_ids, txt=txt, txt_ids=txt_ids, y=vec, timesteps=t_vec, guidance=guidance_vec)
        neg_pred = None
        if i >= timestep_to_start_cfg:
            neg_pred = model(img=x, img_ids=img_ids, txt=neg_txt, txt_ids=neg_txt_ids, y=neg_vec, timesteps=t_vec,
                             guidance=guidance_vec)
        K.euler_cfg_step(x, pred.contiguous(), None if neg_pred is None else neg_pred.contiguous(), t_prev - t_curr,
                         float(true_gs))
        i += 1
    return x.to(in_dtype)


def unpack(x: Tensor, height: int, width: int) -> Tensor:
    """b (h w) (c ph pw) -> b c (h ph) (w pw), ph = pw = 2."""
    h, w = math.ceil(height / 16), math.ceil(width / 16)
    b, _, d = x.shape
    c = d // 4
    return x.view(b, h, w, c, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, c, h * 2, w * 2)
