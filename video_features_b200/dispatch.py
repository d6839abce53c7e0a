"""``--device_ids`` dispatch (reference: main.py:11-55): the list of video indices is cut into contiguous chunks
exactly as ``torch.chunk`` does, one chunk per device.  The reference runs one Python *thread* per GPU inside one
process (GIL-bound); here every device gets its own *process* joined in a torch.distributed group (NCCL on GPUs),
and -- when the caller wants the features back -- a single ``all_gather`` returns every rank's feature blocks in
list order.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops


def shard_indices(n_items: int, n_parts: int, part: int) -> range:
    b, e = ops.shard_range(n_items, n_parts, part)
    return range(b, e)


def gather_feature_blocks(blocks: Sequence[torch.Tensor], width: int, device: torch.device) -> List[torch.Tensor]:
    """All-gather a per-rank list of (T_i, width) float32 blocks.  Two collectives in total: one for the row counts
    of every video, one for the rows (padded to the largest rank).  Returns the blocks of ALL ranks in rank order
    (== list order, because shards are contiguous)."""
    world = dist.get_world_size()
    counts = torch.tensor([b.shape[0] for b in blocks], dtype=torch.int64, device=device)
    n_local = torch.tensor([counts.numel(), int(counts.sum()) if counts.numel() else 0], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    max_videos = max(int(s[0]) for s in sizes)
    max_rows = max(int(s[1]) for s in sizes)
    cnt_pad = torch.zeros(max(max_videos, 1), dtype=torch.int64, device=device)
    cnt_pad[:counts.numel()] = counts
    all_cnt = torch.empty(world * cnt_pad.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_cnt, cnt_pad)
    all_cnt = all_cnt.view(world, cnt_pad.numel())
    rows = torch.zeros((max(max_rows, 1), width), dtype=torch.float32, device=device)
    if counts.numel() and int(counts.sum()):
        rows[:int(counts.sum())] = torch.cat([b.to(device, torch.float32) for b in blocks])
    all_rows = torch.empty((world * rows.shape[0], width), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(all_rows, rows)                    # concatenation along dim 0
    all_rows = all_rows.view(world, rows.shape[0], width)
    out: List[torch.Tensor] = []
    for r in range(world):
        off = 0
        for v in range(int(sizes[r][0])):
            t = int(all_cnt[r, v])
            out.append(all_rows[r, off:off + t])
            off += t
    return out


def _worker(rank: int, world: int, device_ids: List[int], make_extractor: Callable, n_items: int, port: int,
            backend: str) -> None:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        torch.cuda.set_device(device_ids[rank])
        device = torch.device("cuda", device_ids[rank])
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    else:
        device = torch.device("cpu")
        dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        extractor = make_extractor()
        idx = shard_indices(n_items, world, rank)
        if len(idx) > 0:
            extractor(torch.tensor(list(idx), dtype=torch.long, device=device))
        dist.barrier()
        if hasattr(extractor, "progress"):
            extractor.progress.close()
    finally:
        dist.destroy_process_group()


def parallel_feature_extraction(make_extractor: Callable, n_items: int, device_ids: List[int],
                                backend: str = "nccl", port: Optional[int] = None) -> None:
    """One process per entry of ``device_ids``; process p handles the p-th ``torch.chunk`` of ``arange(n_items)``."""
    import torch.multiprocessing as mp
    ids = list(device_ids)[:max(n_items, 1)]            # main.py:51 -- device_ids[:len(indices)]
    world = len(ids)
    port = port or (29500 + os.getpid() % 2000)
    if world == 1:
        _worker(0, 1, ids, make_extractor, n_items, port, backend)
        return
    mp.spawn(_worker, args=(world, ids, make_extractor, n_items, port, backend), nprocs=world, join=True)
