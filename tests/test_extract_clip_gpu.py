"""ExtractCLIP (the reference-facing class) end to end on a synthetic video file: decode -> uni_N/fix_N sampler ->
fused transform + tower, against the oracle run on the very same decoded frames."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_video(path, n=40, h=120, w=160, fps=10.0):
    import cv2
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    assert vw.isOpened()
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for i in range(n):
        vw.write(np.roll(base, 3 * i, axis=1))
    vw.release()


def _args(paths, out, method="uni_6", **kw):
    d = dict(feature_type='CLIP-ViT-B/32', video_paths=paths, flow_paths=None, file_with_video_paths=None,
             video_dir=None, flow_dir=None, extraction_fps=None, extract_method=method, on_extraction='save_numpy',
             output_path=out, output_direct=True, tmp_path=os.path.join(out, 'tmp'))
    d.update(kw)
    return argparse.Namespace(**d)


def test_extract_clip_matches_oracle_and_writes_files(cuda_device, tmp_path, monkeypatch):
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "0")
    from oracle import clip_preprocess, clip_tower
    from video_features_b200 import synthetic_weights, utils
    from video_features_b200.extract.extract_clip import ExtractCLIP
    vid = str(tmp_path / "clip_a.mp4")
    _write_video(vid)
    out = str(tmp_path / "out")
    ex = ExtractCLIP(_args([vid], out))
    res = ex(torch.zeros([1], dtype=torch.long, device=cuda_device))
    assert res == []                                   # external_call False -> nothing returned (extract_clip.py:88)
    saved = np.load(os.path.join(out, "clip_a.npy"))
    assert saved.shape == (6, 512) and saved.dtype == np.float32
    frames, fps, ts = utils.extract_frames(vid, "uni_6")
    ref = clip_tower.encode_image(synthetic_weights.clip_vit_b32_state_dict(0), clip_preprocess.preprocess_batch(frames)).numpy()
    rel = np.linalg.norm(saved - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-3, rel.max()

    # external_call idiom of the reference README (README.md:53-56)
    ex2 = ExtractCLIP(_args([vid], out, method="fix_2"), external_call=True)
    d = ex2(torch.zeros([1], dtype=torch.long, device=cuda_device))[0]
    assert set(d) == {'CLIP-ViT-B/32', 'fps', 'timestamps_ms'}
    n = len(utils.extract_frames(vid, "fix_2")[0])
    assert d['CLIP-ViT-B/32'].shape == (n, 512) and d['timestamps_ms'].shape == (n,)


def test_extract_clip_failure_is_per_video(cuda_device, tmp_path, monkeypatch, capsys):
    """A broken video prints the reference's message and extraction continues (extract_clip.py:78-84)."""
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "0")
    from video_features_b200.extract.extract_clip import ExtractCLIP
    bad = str(tmp_path / "broken.mp4")
    open(bad, "wb").write(b"not a video")
    good = str(tmp_path / "ok.mp4")
    _write_video(good, n=12)
    out = str(tmp_path / "out")
    ex = ExtractCLIP(_args([bad, good], out, method="uni_3"))
    ex(torch.arange(2, device=cuda_device))
    assert "Extraction failed at" in capsys.readouterr().out
    assert os.path.exists(os.path.join(out, "ok.npy")) and not os.path.exists(os.path.join(out, "broken.npy"))


def test_extract_clip_refuses_cpu(tmp_path, monkeypatch, capsys):
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "0")
    from video_features_b200.extract.extract_clip import ExtractCLIP
    vid = str(tmp_path / "v.mp4")
    _write_video(vid, n=12)
    ex = ExtractCLIP(_args([vid], str(tmp_path / "o")), external_call=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ex(torch.zeros([1], dtype=torch.long))


# ------------------------------------------------------------------------------------------------ BASELINE config 1
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = os.path.join(GOLD, "v_GGSY1Qvo990.mp4")              # the reference's own sample (sample/v_GGSY1Qvo990.mp4)


def test_config1_main_py_on_the_sample_video(cuda_device, tmp_path):
    """BASELINE.json configs[0] == the reference README's command (README.md:33), through main.py as a user runs it:
    uni_12 on sample/v_GGSY1Qvo990.mp4, --on_extraction save_numpy --output_direct.  Frame indices and decoded frames
    are pinned to the reference sampler's (fixture made by scripts/make_golden.py from the reference's extract_frames);
    features are compared with the oracle on the very same frames.  Weights: synthetic (real CLIP weights are not
    available offline), the same seeded state dict on both sides."""
    import subprocess
    import sys
    import time
    from oracle import clip_preprocess, clip_tower
    from video_features_b200 import synthetic_weights, utils
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "out")
    env = dict(os.environ, VF_CLIP_SYNTHETIC="0")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(root, "main.py"), "--feature_type", "CLIP-ViT-B/32",
                        "--extract_method", "uni_12", "--video_paths", SAMPLE, "--on_extraction", "save_numpy",
                        "--output_direct", "--output_path", out, "--tmp_path", str(tmp_path / "tmp"), "--device_ids", "0"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=600)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.listdir(out) == ["v_GGSY1Qvo990.npy"]                          # utils/utils.py:83-87 naming
    got = np.load(os.path.join(out, "v_GGSY1Qvo990.npy"))
    assert got.shape == (12, 512) and got.dtype == np.float32
    gold = np.load(os.path.join(GOLD, "config1_sample_video.npz"))
    frames, fps, ts = utils.extract_frames(SAMPLE, "uni_12")
    assert int(gold["frame_cnt"]) == 355 and list(gold["indices"]) == [1, 33, 65, 97, 129, 161, 193, 225, 257, 289, 321, 353]
    assert abs(fps - float(gold["fps"])) < 1e-9 and np.allclose(ts, gold["timestamps_ms"], rtol=0, atol=1e-15)
    # the decoder on this box returns the frames the reference's sampler returned in the build container
    assert [int(f.astype(np.uint64).sum()) for f in frames] == [int(c) for c in gold["frame_checksums"]]
    t1 = time.perf_counter()
    ref = clip_tower.encode_image(synthetic_weights.clip_vit_b32_state_dict(0),
                                  clip_preprocess.preprocess_batch(np.stack(frames))).numpy()
    cpu_s = time.perf_counter() - t1
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    mx = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    print(f"config 1: main.py wall {wall:.2f} s (process start + weights + 12 frames); oracle transform+tower on the host "
          f"{cpu_s:.2f} s; rel-L2 {rel.max():.3e} max-abs {mx.max():.3e}")
    assert rel.max() <= 1e-3 and mx.max() <= 1e-3


def test_extract_clip_list_is_batched_and_matches_per_video_calls(cuda_device, tmp_path, monkeypatch):
    """A list of videos goes through the batched path (several videos per engine call); every video's features must be
    identical to the one-video-per-call path (batch-composition independence), files and per-video errors included."""
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "0")
    from video_features_b200.extract.extract_clip import ExtractCLIP
    vids = []
    for i in range(5):
        v = str(tmp_path / f"b{i}.mp4")
        _write_video(v, n=16 + 4 * i, h=120 if i < 4 else 96, w=160 if i < 4 else 128)     # last one: another geometry
        vids.append(v)
    bad = str(tmp_path / "broken.mp4")
    open(bad, "wb").write(b"not a video")
    vids.insert(2, bad)
    ex = ExtractCLIP(_args(vids, str(tmp_path / "o1"), method="uni_5"), external_call=True)
    ex.batch_frames = 12                                  # 2 videos per call
    got = ex(torch.arange(len(vids), device=cuda_device))
    ex1 = ExtractCLIP(_args(vids, str(tmp_path / "o2"), method="uni_5"), external_call=True)
    ex1.batch_frames = 0
    one = ex1(torch.arange(len(vids), device=cuda_device))
    assert len(got) == len(one) == 5
    for a, b in zip(got, one):
        assert a['CLIP-ViT-B/32'].shape == (5, 512) and np.array_equal(a['CLIP-ViT-B/32'], b['CLIP-ViT-B/32'])
        assert np.array_equal(a['timestamps_ms'], b['timestamps_ms'])


def test_extract_clip_async_sink_and_resume_on_gpu(cuda_device, tmp_path, monkeypatch):
    """VF_ASYNC_SINK=1 (features saved by a writer thread while the next engine call runs) writes the same files as the
    reference-style synchronous sink; VF_RESUME=1 then skips every video whose file exists (no engine work at all)."""
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "0")
    from video_features_b200.extract.extract_clip import ExtractCLIP
    vids = []
    for i in range(6):
        v = str(tmp_path / f"s{i}.mp4")
        _write_video(v, n=12 + 2 * i)
        vids.append(v)
    idx = torch.arange(len(vids), device=cuda_device)
    monkeypatch.delenv("VF_ASYNC_SINK", raising=False)
    monkeypatch.delenv("VF_RESUME", raising=False)
    ex = ExtractCLIP(_args(vids, str(tmp_path / "sync"), method="uni_4"))
    ex.batch_frames = 8                                   # 2 videos per engine call: several calls in flight
    assert ex(idx) == []
    monkeypatch.setenv("VF_ASYNC_SINK", "1")
    ex2 = ExtractCLIP(_args(vids, str(tmp_path / "async"), method="uni_4"))
    ex2.batch_frames = 8
    assert ex2(idx) == []
    for i in range(6):
        a = np.load(tmp_path / "sync" / f"s{i}.npy")
        b = np.load(tmp_path / "async" / f"s{i}.npy")
        assert a.shape == (4, 512) and np.array_equal(a, b)
    assert not [f for f in os.listdir(tmp_path / "async") if f.endswith(".tmp")]       # atomic writes left nothing behind
    # resume: nothing to do -> no engine launches, files untouched
    monkeypatch.setenv("VF_RESUME", "1")
    eng = ex2._engines[0]
    before = eng.launch_count
    stamp = os.path.getmtime(tmp_path / "async" / "s0.npy")
    assert ex2(idx) == []
    assert eng.launch_count == before and os.path.getmtime(tmp_path / "async" / "s0.npy") == stamp
    # one file removed -> exactly that video is extracted again
    os.remove(tmp_path / "async" / "s3.npy")
    assert ex2(idx) == []
    assert eng.launch_count > before
    assert np.array_equal(np.load(tmp_path / "async" / "s3.npy"), np.load(tmp_path / "sync" / "s3.npy"))
