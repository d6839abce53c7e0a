"""``--device_ids`` dispatch (reference: main.py:11-55): the list of video indices is cut into contiguous chunks
exactly as ``torch.chunk`` does, one chunk per device.  The reference runs one Python *thread* per GPU inside one
process (GIL-bound); here every device gets its own *process* joined in a torch.distributed group (NCCL on GPUs),
and -- when the caller wants the features back -- a single ``all_gather`` returns every rank's feature blocks in
list order (``gather_key``).
"""
from __future__ import annotations

import os
import socket
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops


def shard_indices(n_items: int, n_parts: int, part: int) -> range:
    b, e = ops.shard_range(n_items, n_parts, part)
    return range(b, e)


def free_port() -> int:
    """A TCP port nobody is listening on right now (two jobs on one host must not meet on a pid-derived port)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return int(s.getsockname()[1])


def gather_feature_blocks(blocks: Sequence[torch.Tensor], width: int, device: torch.device,
                          rows_on_device: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """All-gather a per-rank list of (T_i, width) float32 blocks.  One collective for the row counts of every video
    (sizes first, so ranks can pad), one for the rows (padded to the largest rank).  Returns the blocks of ALL ranks in
    rank order (== list order, because shards are contiguous).  Without a process group (one device) it is the
    identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [b.to(torch.float32) for b in blocks]            # nothing to exchange: the blocks stay where they are
    world = dist.get_world_size()
    counts = torch.tensor([b.shape[0] for b in blocks], dtype=torch.int64, device=device)
    n_local = torch.tensor([counts.numel(), int(counts.sum()) if counts.numel() else 0], dtype=torch.int64, device=device)
    sizes = torch.empty((world, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n_local[None])
    sizes = sizes.cpu()
    max_videos, max_rows = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    cnt_pad = torch.zeros(max(max_videos, 1), dtype=torch.int64, device=device)
    cnt_pad[:counts.numel()] = counts
    all_cnt = torch.empty(world * cnt_pad.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_cnt, cnt_pad)
    all_cnt = all_cnt.view(world, cnt_pad.numel()).cpu()
    rows = torch.zeros((max(max_rows, 1), width), dtype=torch.float32, device=device)
    if counts.numel() and int(sizes[dist.get_rank(), 1]):
        # one concatenation where the blocks live, then ONE copy to the device (not a copy per video)
        src = rows_on_device if rows_on_device is not None else [b.to(torch.float32) for b in blocks]
        rows[:int(sizes[dist.get_rank(), 1])].copy_(torch.cat(list(src)), non_blocking=True)
    all_rows = torch.empty((world * rows.shape[0], width), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(all_rows, rows)                    # concatenation along dim 0
    all_rows = all_rows.view(world, rows.shape[0], width)
    out: List[torch.Tensor] = []
    for r in range(world):
        cnt = all_cnt[r, :int(sizes[r, 0])].tolist()
        out.extend(all_rows[r, :sum(cnt)].split(cnt) if cnt else [])
    return out


def run_shard(extractor, n_items: int, rank: int, world: int, device: torch.device,
              gather_key: Optional[str] = None) -> Optional[List[torch.Tensor]]:
    """The body of one worker: run ``extractor`` on the rank's ``torch.chunk`` of ``arange(n_items)``; with
    ``gather_key`` the (T_i, width) feature blocks of every video are all-gathered and returned in list order."""
    idx = shard_indices(n_items, world, rank)
    if hasattr(extractor, "progress"):
        extractor.progress.total = len(idx)                        # each bar counts its own shard
        extractor.progress.refresh()
    if gather_key is not None and hasattr(extractor, "keep_features"):
        extractor.keep_features = True
    res = extractor(torch.tensor(list(idx), dtype=torch.long, device=device)) if len(idx) > 0 else []
    if gather_key is None:
        return None
    blocks = [torch.as_tensor(d[gather_key], dtype=torch.float32) for d in (res or [])]
    chunks = getattr(extractor, "device_chunks", None)
    rows_dev = None
    if chunks and sum(c.shape[0] for _, c in chunks) == sum(b.shape[0] for b in blocks):
        # the extractor kept the same rows on the GPU, one tensor per engine call: gather those (no host concatenation
        # of thousands of blocks, no H2D copy)
        rows_dev = [c for _, c in sorted(chunks, key=lambda pc: pc[0])]
    width = blocks[0].shape[1] if blocks else 0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        wmax = torch.tensor([width], dtype=torch.int64, device=device)
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)                # a rank with an empty shard learns the width
        width = int(wmax)
    gathered = gather_feature_blocks(blocks, width, device, rows_dev)
    if chunks:
        extractor.device_chunks = []                               # the per-call device copies are no longer needed
    return gathered


def _worker(rank: int, world: int, device_ids: List[int], make_extractor: Callable, n_items: int, port: int,
            backend: str, gather_key: Optional[str], on_gathered: Optional[Callable]) -> None:
    grouped = world > 1
    if backend == "nccl":
        torch.cuda.set_device(device_ids[rank])
        device = torch.device("cuda", device_ids[rank])
    else:
        device = torch.device("cpu")
    if grouped:                                                    # one device: no rendezvous, no port, no collective
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["MASTER_PORT"] = str(port)
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        extractor = make_extractor()
        gathered = run_shard(extractor, n_items, rank, world, device, gather_key)
        if grouped:
            dist.barrier()
        if hasattr(extractor, "progress"):
            extractor.progress.close()
        if gathered is not None and on_gathered is not None and rank == 0:
            on_gathered([g.cpu() for g in gathered])
    finally:
        if grouped:
            dist.destroy_process_group()


def parallel_feature_extraction(make_extractor: Callable, n_items: int, device_ids: List[int],
                                backend: str = "nccl", port: Optional[int] = None, gather_key: Optional[str] = None,
                                on_gathered: Optional[Callable] = None) -> None:
    """One process per entry of ``device_ids``; process p handles the p-th ``torch.chunk`` of ``arange(n_items)``.
    ``gather_key``/``on_gathered``: all-gather that feature of every video and hand the list (list order) to
    ``on_gathered`` on rank 0."""
    import torch.multiprocessing as mp
    ids = list(device_ids)[:max(n_items, 1)]            # main.py:51 -- device_ids[:len(indices)]
    world = len(ids)
    if world == 1:
        _worker(0, 1, ids, make_extractor, n_items, 0, backend, gather_key, on_gathered)
        return
    port = port or free_port()
    mp.spawn(_worker, args=(world, ids, make_extractor, n_items, port, backend, gather_key, on_gathered), nprocs=world,
             join=True)
