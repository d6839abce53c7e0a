"""numpy restatement of Pillow's 8-bit ``Image.resize`` (antialiased separable resample) -- test oracle only.

Follows third-party Pillow ``src/libImaging/Resample.c`` (precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc / Vertical_8bpc, ImagingResample).  The reference reaches it through torchvision
``Resize`` inside the CLIP transform (models/CLIP/extract_clip.py:112) and directly in
models/i3d/transforms/transforms.py:121,125.  Reference env pinned Pillow 7.2.0 (conda_env_torch_zoo.yml:44);
pinned here against the installed Pillow byte-for-byte (tests/test_oracle_resample.py, tests/golden/).
"""
from __future__ import annotations

import math

import numpy as np

BILINEAR, BICUBIC = 2, 3          # PIL.Image.BILINEAR / BICUBIC
PRECISION_BITS = 32 - 8 - 2       # 22


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def coefficients(in_size: int, out_size: int, filt: int):
    """-> (ksize, bounds[out,2] (first index, count), coefs[out,ksize] int32 fixed point)."""
    f, support = {BICUBIC: (_bicubic, 2.0), BILINEAR: (_bilinear, 1.0)}[filt]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [f((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coefs[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, coefs


def _pass(img: np.ndarray, axis: int, out_size: int, filt: int) -> np.ndarray:
    in_size = img.shape[axis]
    _, bounds, coefs = coefficients(in_size, out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        k = coefs[xx, :n].astype(np.int64).reshape((n,) + (1,) * (src.ndim - 1))
        acc = (1 << (PRECISION_BITS - 1)) + (src[x0:x0 + n] * k).sum(axis=0)
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize(img: np.ndarray, out_h: int, out_w: int, filt: int) -> np.ndarray:
    """img: (H,W,C) uint8.  Horizontal pass first (rounded to uint8), then vertical; a pass whose axis size is
    unchanged is skipped (ImagingResample need_horizontal / need_vertical)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    if out_w != img.shape[1]:
        img = _pass(img, 1, out_w, filt)
    if out_h != img.shape[0]:
        img = _pass(img, 0, out_h, filt)
    return np.ascontiguousarray(img)


def short_side_geometry(h: int, w: int, size: int, to_smaller_edge: bool = True):
    """models/i3d/transforms/transforms.py:114-125 ``resize`` with an int size -> (out_h, out_w).
    (== torchvision Resize(int) when to_smaller_edge is True.)"""
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if (w < h) == bool(to_smaller_edge):
        return int(size * h / w), size
    return size, int(size * w / h)
