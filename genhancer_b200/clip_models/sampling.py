"""``prepare_clip`` -- drop-in for /root/reference/Continuous/clip_models/sampling.py:9-42."""
from __future__ import annotations

import torch
from torch import Tensor


def make_img_ids(bs: int, h2: int, w2: int, device, t: float = 0.0) -> Tensor:
    """(t, row, col) ids [bs, h2*w2, 3], built on the device (the reference builds them on the CPU every step)."""
    ids = torch.zeros(h2, w2, 3, device=device)
    ids[..., 0] = t
    ids[..., 1] = torch.arange(h2, device=device)[:, None]
    ids[..., 2] = torch.arange(w2, device=device)[None, :]
    return ids.reshape(1, h2 * w2, 3).expand(bs, -1, -1)


def patchify(z: Tensor) -> Tensor:
    """'b c (h ph) (w pw) -> b (h w) (c ph pw)', ph = pw = 2."""
    B, C, H, W = z.shape
    return z.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)


def prepare_clip(clip, original_img: Tensor, img: Tensor) -> dict[str, Tensor]:
    bs, c, h, w = img.shape
    _, projection_clip, projection_t5 = clip(original_img)
    txt = projection_t5
    if txt.shape[0] == 1 and bs > 1:
        txt = txt.expand(bs, -1, -1)
    vec = projection_clip
    if vec.shape[0] == 1 and bs > 1:
        vec = vec.expand(bs, -1)
    dev = img.device
    return {
        "img": patchify(img),
        "img_ids": make_img_ids(bs, h // 2, w // 2, dev),
        "txt": txt.to(dev),
        "txt_ids": torch.zeros(bs, txt.shape[1], 3, device=dev),
        "vec": vec.to(dev),
    }
