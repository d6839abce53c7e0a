"""Restatement of the CLIP ``preprocess`` transform -- test oracle only.

The reference gets ``preprocess`` from ``clip.load`` (models/CLIP/extract_clip.py:47) and applies it per frame to
``Image.fromarray(frame)`` (extract_clip.py:107-113).  ``clip`` is third-party (openai/CLIP, un-vendored, unpinned);
its published ``clip.clip._transform(224)`` is
    Resize(224, BICUBIC) -> CenterCrop(224) -> convert("RGB") -> ToTensor() -> Normalize(mean, std)
restated here with Pillow + torch fp32 arithmetic in torchvision's operation order.  Note the reference never swaps
the decoder's BGR to RGB before this transform; neither does this code.
"""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)
SIZE = 224


def center_crop_offset(dim: int, crop: int) -> int:
    return int(round((dim - crop) / 2.0))          # torchvision CenterCrop (python banker's rounding)


def resized_geometry(h: int, w: int, size: int = SIZE):
    """torchvision Resize(int): short side -> size, long side int(size*long/short); identity if already there."""
    short, long = (w, h) if w <= h else (h, w)
    if short == size:
        return h, w
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def preprocess_frame(frame: np.ndarray) -> torch.Tensor:
    """frame: HxWx3 uint8 (as the decoder delivers it) -> (3,224,224) fp32."""
    img = Image.fromarray(frame)
    h, w = frame.shape[:2]
    oh, ow = resized_geometry(h, w)
    if (oh, ow) != (h, w):
        img = img.resize((ow, oh), Image.BICUBIC)
    top, left = center_crop_offset(oh, SIZE), center_crop_offset(ow, SIZE)
    img = img.crop((left, top, left + SIZE, top + SIZE)).convert("RGB")
    x = torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).to(torch.float32).div(255)
    mean = torch.tensor(MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(STD, dtype=torch.float32)[:, None, None]
    return x.sub_(mean).div_(std)


def preprocess_batch(frames) -> torch.Tensor:
    return torch.stack([preprocess_frame(f) for f in frames])
