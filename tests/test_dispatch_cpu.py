"""--device_ids dispatch: shard arithmetic and the feature all-gather, on a world_size-2 gloo group (CPU)."""
import argparse
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_features_b200.dispatch import gather_feature_blocks, shard_indices
    n_videos = 5
    idx = list(shard_indices(n_videos, world, rank))
    # video v has v+1 rows, every element equals v
    blocks = [torch.full((v + 1, 8), float(v)) for v in idx]
    allb = gather_feature_blocks(blocks, 8, torch.device("cpu"))
    # the same rows handed over in a few large pieces (what ExtractCLIP keeps per engine call) give the same answer
    rows = torch.cat(blocks) if blocks else torch.zeros((0, 8))
    pieces = [rows[:2], rows[2:]] if rows.shape[0] > 2 else [rows]
    allb2 = gather_feature_blocks(blocks, 8, torch.device("cpu"), rows_on_device=pieces)
    assert len(allb) == len(allb2) and all(torch.equal(a, b) for a, b in zip(allb, allb2))
    q.put((rank, idx, [(b.shape[0], float(b[0, 0])) for b in allb]))
    dist.destroy_process_group()


def test_gather_feature_blocks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]          # torch.chunk(arange(5), 2)
    want = [(v + 1, float(v)) for v in range(5)]
    assert res[0][2] == want and res[1][2] == want                   # every rank holds all blocks, list order


class _Recorder:
    """Stands in for an extractor: records which indices each rank received."""

    def __init__(self, out_dir):
        self.out_dir = out_dir

    def __call__(self, indices):
        rank = dist.get_rank()
        with open(os.path.join(self.out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(",".join(str(int(i)) for i in indices))


def _make_recorder(out_dir):
    return _Recorder(out_dir)


def test_parallel_feature_extraction_gloo(tmp_path):
    import functools
    from video_features_b200.dispatch import parallel_feature_extraction
    parallel_feature_extraction(functools.partial(_make_recorder, str(tmp_path)), 10, [0, 1, 2, 3], backend="gloo",
                                port=29900 + os.getpid() % 90)
    got = [open(tmp_path / f"rank{r}.txt").read() for r in range(4)]
    assert got == ["0,1,2", "3,4,5", "6,7,8", "9"]                  # == torch.arange(10).chunk(4)


def test_main_cli_parser_matches_reference_flags():
    import main
    p = main.make_parser()
    a = p.parse_args(["--feature_type", "CLIP-ViT-B/32", "--video_paths", "x.mp4", "--extract_method", "uni_12",
                      "--device_ids", "0", "1", "--on_extraction", "save_numpy", "--output_direct"])
    assert a.device_ids == [0, 1] and a.flow_type == "pwc" and a.batch_size == 1 and a.resize_to_smaller_edge is True
    assert a.tmp_path == "./tmp" and a.output_path == "./output" and a.cpu is False
    with pytest.raises(NotADirectoryError):
        main.build_extractor(argparse.Namespace(feature_type="nonsense"))


class _BlockMaker:
    """Stands in for ExtractCLIP under dispatch: video v yields a (v % 3 + 1, 4) block filled with v."""
    keep_features = False

    def __call__(self, indices):
        assert self.keep_features is True                      # dispatch asked for the features back
        return [{'feat': torch.full((int(i) % 3 + 1, 4), float(i)).numpy()} for i in indices]


def _make_block_maker():
    return _BlockMaker()


def _write_gathered(path, blocks):
    torch.save([b.clone() for b in blocks], path)


@pytest.mark.parametrize("world", [1, 3])
def test_parallel_feature_extraction_gathers_blocks_in_list_order(tmp_path, world):
    """gather_key: every rank's feature blocks come back through ONE all-gather, list order, on rank 0 -- including a
    world where the last rank's shard is shorter, and the single-device case (no process group at all)."""
    import functools
    from video_features_b200.dispatch import parallel_feature_extraction
    target = str(tmp_path / "g.pt")
    parallel_feature_extraction(_make_block_maker, 7, list(range(world)), backend="gloo", gather_key='feat',
                                on_gathered=functools.partial(_write_gathered, target))
    blocks = torch.load(target)
    assert [tuple(b.shape) for b in blocks] == [(v % 3 + 1, 4) for v in range(7)]
    assert all(float(b[0, 0]) == float(v) and float(b[-1, -1]) == float(v) for v, b in enumerate(blocks))


def test_free_port_is_bindable():
    import socket
    from video_features_b200.dispatch import free_port
    p = free_port()
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", p))
