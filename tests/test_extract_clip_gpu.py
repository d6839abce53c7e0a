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
    from video_features_b200 import utils
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
    ref = clip_tower.encode_image(clip_tower.synthetic_state_dict(0), clip_preprocess.preprocess_batch(frames)).numpy()
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
