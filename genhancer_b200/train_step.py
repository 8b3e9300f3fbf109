"""The stage-1 / stage-2 training step -- the hot path itself.

In the reference the step has no function boundary: it is ~25 lines inside the loop of every ``train_*.py``
(/root/reference/Continuous/train_SigLIP_stage1.py:238-275; video variants
train_OpenAICLIP_video_stage1.py:347-468).  Here it is one object so that the train scripts, ``bench.py`` and the
parity tests all run the same code:

    x_1  = patchify(vae.encode(NORMALIZE_VAE(img)))              # frozen AE, RNG draw #1 (randn_like(mean))
    cond = clip_vis(NORMALIZE_CLIP(img))                          # tower (+LoRA in stage 2) + projectors
    t    = sigmoid(randn(B) * scale_factor)                       # RNG draw #2
    x_0  = randn_like(x_1)                                        # RNG draw #3
    x_t  = (1 - t) x_1 + t x_0                                    # gh_fm_interp_fwd
    pred = dit(x_t, img_ids, txt, txt_ids, t, vec, guidance=4)    # fused DiT engine
    loss = mse(pred.float(), x_0 - x_1)                           # gh_fm_mse_loss_fwdbwd (loss + dpred in one pass)

The three RNG draws stay ``torch.randn`` calls of the reference's shapes/dtypes/order on the same device, so a
given seed reproduces the reference's t, x_0 and AE noise bit-for-bit.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor

from . import kernels as K
from .clip_models.sampling import make_img_ids
from .kernels import BF16, F32

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # train_MetaCLIP_stage1.py:54-55
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
SIGLIP_MEAN = (0.5, 0.5, 0.5)                             # train_SigLIP_stage1.py:54-55
SIGLIP_STD = (0.5, 0.5, 0.5)


# The frozen AE encoder and the tower read the same image and nothing of each other: the AE runs on a side stream,
# forked at the top of the step and joined before x_1 is first read.  Two independent chains of kernels overlap at
# their edges (an SM done with a conv's last wave starts on a ViT GEMM instead of idling until the kernel boundary)
# and the HBM-bound GroupNorm passes run beside tensor-bound GEMMs.  The torch.randn of the AE noise is still the first
# RNG call of the step in host order, so the draws stay those of the reference.  GH_AE_STREAM=0 switches it off.
OVERLAP_AE = os.environ.get("GH_AE_STREAM", "1") != "0"
_SIDE_STREAMS: dict = {}


def encode_on_side_stream(vae, img: Tensor, noise):
    """-> (x_1, join): ``vae.encode_patchified`` issued on the side stream; call ``join()`` on the consumer's stream
    before reading x_1."""
    if not (OVERLAP_AE and img.is_cuda):
        return vae.encode_patchified(img, 0.5, 0.5, noise=noise), (lambda: None)
    dev = img.device
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)          # the image (and, next step, everything that last read this step's buffers)
    with torch.cuda.stream(side):
        x_1 = vae.encode_patchified(img, 0.5, 0.5, noise=noise)
    return x_1, (lambda: torch.cuda.current_stream(dev).wait_stream(side))


class _FlowMatchLoss(torch.autograd.Function):
    """mean((pred.float() - (x_0 - x_1))^2) with d loss / d pred produced in the same pass
    (train_SigLIP_stage1.py:263 + its autograd)."""

    @staticmethod
    def forward(ctx, pred, x_0, x_1):
        loss, dpred = K.fm_mse_loss(pred, x_0, x_1, 1.0, want_grad=True)
        ctx.save_for_backward(dpred)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g.to(dpred.dtype), None, None


def flow_match_loss(pred: Tensor, x_0: Tensor, x_1: Tensor) -> Tensor:
    return _FlowMatchLoss.apply(pred.contiguous(), x_0, x_1)


def sample_t_x0(x_1: Tensor, scale_factor: float, t: Tensor | None = None, x_0: Tensor | None = None):
    """RNG draws #2 and #3 in the reference's order (train_SigLIP_stage1.py:248-249)."""
    bs = x_1.shape[0]
    if t is None:
        t = torch.sigmoid(torch.randn((bs,), device=x_1.device) * scale_factor)
    if x_0 is None:
        x_0 = torch.randn_like(x_1)
    return t, x_0


class Stage1ImageStep:
    """Image-mode step (global-token conditioning): train_{SigLIP,MetaCLIP}_stage{1,2_*}.py and the
    OpenAI-CLIP variants the reference's docs name."""

    def __init__(self, clip_vis, dit, vae, clip_mean=OPENAI_CLIP_MEAN, clip_std=OPENAI_CLIP_STD,
                 scale_factor: float = 1.0, guidance: float = 4.0):
        self.clip_vis, self.dit, self.vae = clip_vis, dit, vae
        self.clip_mean, self.clip_std = tuple(clip_mean), tuple(clip_std)
        self.scale_factor, self.guidance = scale_factor, guidance
        self._ids = {}

    def _static(self, B: int, h2: int, w2: int, n_txt: int, dev):
        key = (B, h2, w2, n_txt, str(dev))
        if key not in self._ids:  # the reference rebuilds these on the CPU and copies them every step
            self._ids[key] = (make_img_ids(B, h2, w2, dev).contiguous(), torch.zeros(B, n_txt, 3, device=dev),
                              torch.full((B,), self.guidance, device=dev, dtype=BF16))
        return self._ids[key]

    def __call__(self, img: Tensor, ae_noise: Tensor | None = None, t: Tensor | None = None,
                 x_0: Tensor | None = None, return_parts: bool = False, before_trainable=None):
        """img: [B,3,S,S] fp32 in [0,1] on the device.  Returns the scalar loss (autograd-connected).

        ``before_trainable()`` is called once, after everything that reads NO trainable parameter (the frozen AE
        encoder, and in stage 1 the frozen tower) and before the first kernel that does.  The data-parallel loop
        puts the PREVIOUS step's ``reducer.finish(); opt.step(); opt.zero_grad()`` there, so the tail of the
        gradient all-reduce hides under ~30 ms of frozen forward instead of being exposed after backward; the
        arithmetic is that of the sequential loop (weights are updated before they are next read)."""
        B = img.shape[0]
        dev = img.device
        x_1, join_ae = encode_on_side_stream(self.vae, img, ae_noise)                    # [B, L, 64] fp32
        tower_trains = any(p.requires_grad for p in self.clip_vis.model.parameters())
        if before_trainable is not None and tower_trains:
            before_trainable()
        cls = self.clip_vis.class_token(img, _norm=(self.clip_mean, self.clip_std))
        if before_trainable is not None and not tower_trains:
            before_trainable()
        vec, txt = self.clip_vis.project(cls)
        join_ae()
        h2 = w2 = int(round(x_1.shape[1] ** 0.5))
        img_ids, txt_ids, guidance = self._static(B, h2, w2, txt.shape[1], dev)
        t, x_0 = sample_t_x0(x_1, self.scale_factor, t, x_0)
        x_t = K.fm_interp(x_1, x_0, t)
        pred = self.dit(img=x_t, img_ids=img_ids, txt=txt.to(BF16), txt_ids=txt_ids, y=vec.to(BF16),
                        timesteps=t.to(BF16), guidance=guidance)
        loss = flow_match_loss(pred, x_0, x_1)
        if return_parts:
            return loss, dict(x_1=x_1, x_t=x_t, t=t, x_0=x_0, pred=pred, vec=vec, txt=txt)
        return loss
