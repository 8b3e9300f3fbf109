"""Video modes of the training step: per-patch conditioning through the ``VisualPromptAdapter``.

Drop-ins for the script-local pieces of /root/reference/Continuous/train_OpenAICLIP_*_stage{1,2_all}.py:
  * ``VisualPromptAdapter`` (train_OpenAICLIP_video_stage1.py:85-97; state_dict keys ``proj.{0,2,3}.*``, the content of
    ``checkpoint-visual-adapter-N.bin``) and ``SuperModel`` (:99-114; attributes ``clip_vis``, ``dit``, ``visual_adapter``),
  * ``create_spatio_temporal_ids`` (:128-151),
  * ``build_windows_with_mask`` (train_OpenAICLIP_sliding_windows_nextpredic_stage1.py:149-204),
  * the step body of the eight video scripts as ``VideoStep``: which frames condition, which
    frame is the target, and the RoPE time index of each.

Reference quirks handled as SURVEY.md Appendix C suggests: the discarded ``prepare_clip`` call that the scripts make
only to obtain ``img_ids`` (one wasted tower forward per step, Q2) is skipped -- the ids are built directly; the
hard-coded 24x24 grid of video_stage1 (Q4) is derived from the token count.
"""
from __future__ import annotations

import random as _random

import torch
from torch import Tensor, nn

from . import kernels as K
from . import ops
from .clip_models import vision_tower as vt
from .clip_models.sampling import make_img_ids
from .kernels import ACT_SILU, BF16
from .train_step import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, encode_on_side_stream, flow_match_loss, sample_t_x0


class VisualPromptAdapter(nn.Module):
    """Linear(in, 2 in) -> SiLU -> Linear(2 in, out) -> LayerNorm(out), applied to every patch token."""

    def __init__(self, in_dim: int = 1024, out_dim: int = 4096):
        super().__init__()
        self.proj = nn.Sequential(nn.Linear(in_dim, in_dim * 2), nn.SiLU(), nn.Linear(in_dim * 2, out_dim),
                                  nn.LayerNorm(out_dim))

    def forward(self, x: Tensor) -> Tensor:
        h = ops.linear(x, self.proj[0].weight, self.proj[0].bias, act=ACT_SILU)
        h = ops.linear(h, self.proj[2].weight, self.proj[2].bias)
        return ops.layer_norm(h, self.proj[3].weight, self.proj[3].bias, self.proj[3].eps)


class SuperModel(nn.Module):
    def __init__(self, clip_vis, dit, adapter_in_dim: int = 1024, adapter_out_dim: int = 4096):
        super().__init__()
        self.clip_vis = clip_vis
        self.dit = dit
        self.visual_adapter = VisualPromptAdapter(in_dim=adapter_in_dim, out_dim=adapter_out_dim)

    def get_clip_vis(self):
        return self.clip_vis

    def get_dit(self):
        return self.dit


def create_spatio_temporal_ids(h: int, w: int, time_step, device) -> Tensor:
    """[h*w, 3] integer ids (time, row, col)."""
    gh, gw = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    fh, fw = gh.flatten(), gw.flatten()
    return torch.stack([torch.full_like(fh, fill_value=time_step), fh, fw], dim=1)


def build_windows_with_mask(frames: Tensor, frame_mask: Tensor, window_cond: int = 3, window_stride: int = 1,
                            max_windows_per_video: int | None = 8, rng=None):
    """frames [B,T,C,H,W] (padded), frame_mask [B,T] -> (cond_0 .. cond_{window_cond-1}, target, avg_nw, bs_eff) with
    every tensor [bs_eff,C,H,W], or None when no video yields a window.  One device gather instead of the
    reference's per-frame ``torch.stack`` loops; window choice (incl. ``random.sample`` when a video has more than
    ``max_windows_per_video`` windows) is the reference's."""
    if frames.ndim != 5:
        raise AssertionError(f"expect [B,T,C,H,W], got {tuple(frames.shape)}")
    if frame_mask.ndim != 2:
        raise AssertionError(f"expect [B,T], got {tuple(frame_mask.shape)}")
    B, T = frames.shape[:2]
    lengths = frame_mask.sum(dim=1).tolist()  # one host sync for the whole batch (the reference syncs per video)
    vid, start, nw = [], [], []
    for i in range(B):
        Ti = int(lengths[i])
        if Ti < window_cond + 1:
            continue
        starts = list(range(0, Ti - window_cond, window_stride))
        if not starts:
            continue
        if max_windows_per_video is not None and max_windows_per_video > 0 and len(starts) > max_windows_per_video:
            starts = sorted((rng or _random).sample(starts, k=max_windows_per_video))
        nw.append(len(starts))
        vid += [i] * len(starts)
        start += starts
    if not vid:
        return None
    flat = frames.reshape(B * T, *frames.shape[2:])
    base = torch.tensor(vid, device=frames.device) * T + torch.tensor(start, device=frames.device)
    out = [flat.index_select(0, base + j) for j in range(window_cond + 1)]
    return (*out[:window_cond], out[window_cond], float(sum(nw)) / float(max(1, len(nw))), len(vid))


# cond frame keys -> RoPE time index; target key -> time index   
VIDEO_MODES = {
    "video": (("start_frame", 0), ("end_frame", 2)), "video_target": ("middle_frame", 1),
    "nextpredic": (("start_frame", 0),), "nextpredic_target": ("middle_frame", 1),
    "use2frames_nextpredic": (("start_frame", 0), ("middle_frame", 1)), "use2frames_nextpredic_target": ("end_frame", 2),
}


class VideoStep:
    """cond frames -> tower patch tokens -> adapter -> txt stream of the DiT; target frame -> AE latent.

    ``cond_times`` / ``target_time`` are the RoPE time indices, e.g. (0, 2) -> 1 for frame interpolation,
    (0, 1) -> 2 for use2frames next-frame prediction, (0, 1, 2) -> 3 for sliding windows."""

    def __init__(self, super_model: SuperModel, vae, cond_times=(0, 2), target_time=1, clip_mean=OPENAI_CLIP_MEAN,
                 clip_std=OPENAI_CLIP_STD, scale_factor: float = 1.0, guidance: float = 4.0, tower_grad: bool = False):
        self.m, self.vae = super_model, vae
        self.cond_times, self.target_time = tuple(cond_times), target_time
        self.clip_mean, self.clip_std = tuple(clip_mean), tuple(clip_std)
        self.scale_factor, self.guidance = scale_factor, guidance
        self.tower_grad = tower_grad
        self._ids = {}

    def _static(self, B, h2, w2, g, dev):
        key = (B, h2, w2, g, str(dev))
        if key not in self._ids:
            txt_ids = torch.cat([create_spatio_temporal_ids(g, g, t, dev) for t in self.cond_times], dim=0)
            self._ids[key] = (make_img_ids(B, h2, w2, dev, float(self.target_time)).contiguous(),
                              txt_ids.float()[None].expand(B, -1, -1).contiguous(),
                              torch.full((B,), self.guidance, device=dev, dtype=BF16))
        return self._ids[key]

    def __call__(self, cond_frames, target: Tensor, ae_noise=None, t=None, x_0=None, return_parts: bool = False,
                 before_trainable=None):
        """cond_frames: sequence of [B,3,S,S] fp32 frames (as the dataset yields them); target [B,3,S,S].
        ``before_trainable``: see Stage1ImageStep.__call__ (the deferred exchange + update of the previous step)."""
        if len(cond_frames) != len(self.cond_times):
            raise ValueError(f"expected {len(self.cond_times)} conditioning frames, got {len(cond_frames)}")
        B, dev = target.shape[0], target.device
        n = len(cond_frames)
        x_1, join_ae = encode_on_side_stream(self.vae, target, ae_noise)     # beside the tower pass (train_step.py)
        model = self.m.clip_vis.model
        if before_trainable is not None and self.tower_grad:
            before_trainable()
        frames = torch.cat(list(cond_frames), dim=0) if n > 1 else cond_frames[0]       # one tower pass for all frames
        ctx = torch.enable_grad() if self.tower_grad else torch.no_grad()
        with ctx:
            out = model.vision_model(frames, output_hidden_states=True, _norm=(self.clip_mean, self.clip_std))
            patches = out.last_hidden_state[:, 1:, :]                                     # raw last layer, no CLS
            P, D = patches.shape[1], patches.shape[2]
            vec = vt.project(model, out.pooler_output).view(n, B, -1).float().mean(dim=0)  # mean of the frames' vecs
            visual_context = patches.reshape(n, B, P, D).permute(1, 0, 2, 3).reshape(B, n * P, D)
        g = int(round(P ** 0.5))
        if g * g != P:
            raise AssertionError(f"patch tokens must form a square grid, got P={P}")
        if before_trainable is not None and not self.tower_grad:
            before_trainable()
        txt = self.m.visual_adapter(visual_context)
        join_ae()
        h2 = w2 = int(round(x_1.shape[1] ** 0.5))
        img_ids, txt_ids, guidance = self._static(B, h2, w2, g, dev)
        if txt.shape[1] != txt_ids.shape[1]:
            raise AssertionError(f"txt/txt_ids length mismatch: {txt.shape[1]} vs {txt_ids.shape[1]}")
        t, x_0 = sample_t_x0(x_1, self.scale_factor, t, x_0)
        x_t = K.fm_interp(x_1, x_0, t)
        pred = self.m.dit(img=x_t, img_ids=img_ids, txt=txt.to(BF16), txt_ids=txt_ids, y=vec.to(BF16),
                          timesteps=t.to(BF16), guidance=guidance)
        loss = flow_match_loss(pred, x_0, x_1)
        if return_parts:
            return loss, dict(x_1=x_1, x_t=x_t, t=t, x_0=x_0, pred=pred, vec=vec, txt=txt, txt_ids=txt_ids,
                              img_ids=img_ids)
        return loss
