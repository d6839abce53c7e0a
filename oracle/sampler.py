"""Restatement of the reference frame sampler index arithmetic -- test oracle only.

Follows utils/utils.py:297-333 ``extract_frames`` (``uni_N`` / ``fix_N``) and main.py:49-53 (contiguous
``torch.chunk`` shard of the video indices over ``--device_ids``).
"""
from __future__ import annotations

import numpy as np


def sample_indices(method: str, frame_cnt: int, fps: float) -> np.ndarray:
    ext = method.split('_')[0]
    params = method.split('_')[1:]
    if ext == "fix":
        samples_num = int(frame_cnt / fps * int(params[0]))        # utils/utils.py:315
    elif ext == "uni":
        samples_num = int(params[0])                               # utils/utils.py:323
    else:
        raise NotImplementedError(f'{ext} are not supported')      # utils/utils.py:333
    return np.linspace(1, frame_cnt - 2, samples_num).astype(int)  # utils/utils.py:317,326


def timestamps_ms(indices: np.ndarray, fps: float):
    mspf = 0.001 / fps                                             # utils/utils.py:312 (sic)
    return [i * mspf for i in indices]


def shard(n_items: int, n_parts: int):
    """[(begin, end)] per part: torch.chunk(arange(n), k) with k = min(n_parts, n) (main.py:51-53)."""
    k = min(n_parts, n_items)
    out = []
    cs = -(-n_items // k) if k > 0 else 0
    for p in range(n_parts):
        b = min(p * cs, n_items) if k > 0 else 0
        e = min(b + cs, n_items) if k > 0 else 0
        out.append((b, e))
    return out
