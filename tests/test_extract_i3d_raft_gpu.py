"""ExtractI3D / ExtractRAFT (reference-facing classes) on a synthetic video, real checkpoints, against the oracle run
on the same decoded frames."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = all(os.path.exists(os.path.join(ROOT, "checkpoints", f)) for f in ("i3d_rgb.pt", "i3d_flow.pt", "raft-sintel.pth"))


def _write_video(path, n, h=120, w=160, fps=25.0):
    import cv2
    from oracle import raft_net
    fr = raft_net.synthetic_frames(n, h, w, seed=11, shift=(0.8, 0.5)).permute(0, 2, 3, 1).numpy().astype(np.uint8)
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    assert vw.isOpened()
    for f in fr:
        vw.write(f)
    vw.release()


def _ns(**kw):
    d = dict(feature_type='i3d', video_paths=None, flow_paths=None, file_with_video_paths=None, video_dir=None, flow_dir=None,
             extraction_fps=None, extract_method=None, on_extraction='save_numpy', output_path='./output',
             output_direct=False, tmp_path='./tmp', streams=None, flow_type='raft', stack_size=None, step_size=None,
             show_pred=False, keep_tmp_files=False, batch_size=1, resize_to_smaller_edge=True, side_size=None)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.skipif(not HAVE, reason="reference checkpoint copies not present (scripts/fetch_checkpoints.py)")
def test_extract_i3d_two_streams_vs_oracle(cuda_device, tmp_path):
    from PIL import Image
    from oracle import i3d_net, raft_net
    from video_features_b200 import utils
    from video_features_b200.extract.extract_i3d import ExtractI3D
    vid = str(tmp_path / "clip.mp4")
    _write_video(vid, 20)
    out = str(tmp_path / "out")
    ex = ExtractI3D(_ns(video_paths=[vid], output_path=out, tmp_path=str(tmp_path / "tmp"), stack_size=12, step_size=12),
                    external_call=True)
    res = ex(torch.zeros([1], dtype=torch.long, device=cuda_device))[0]
    assert set(res) == {'rgb', 'flow', 'fps', 'timestamps_ms'}
    # a 20-frame video is shorter than 65 frames: the reference resamples it to 65 (extract_i3d.py:250-255), so with
    # stack_size = step_size = 12 there are (65-1)//12 = 5 stacks
    assert res['rgb'].shape == (5, 1024) and res['flow'].shape == (5, 1024) and res['rgb'].dtype == np.float64
    res = {k: (v[:1] if k in ('rgb', 'flow') else v) for k, v in res.items()}       # compare the first stack
    # oracle on the same decoded frames: PIL bilinear resize to 256 -> rgb / RAFT+flow transforms -> I3D
    rd = utils.VideoReader(vid)
    ix = np.linspace(1, rd.frame_cnt - 1, 65).astype(int)[:13]
    frames = [rd.get_frame(int(i)) for i in ix]
    rs = torch.stack([torch.from_numpy(np.asarray(Image.fromarray(f).resize((341, 256), Image.BILINEAR)).copy())
                      for f in frames]).permute(0, 3, 1, 2).float().to(cuda_device)
    sd_rgb = {k: v.to(cuda_device) for k, v in torch.load(os.path.join(ROOT, "checkpoints", "i3d_rgb.pt")).items()}
    ref_rgb = i3d_net.forward_features(sd_rgb, i3d_net.rgb_transform(rs[:-1]))
    rel = float((torch.from_numpy(res['rgb']).to(cuda_device) - ref_rgb).norm() / ref_rgb.norm())
    print("ExtractI3D rgb vs oracle:", rel)
    assert rel < 1e-3
    sd_raft = {k: v.to(cuda_device) for k, v in torch.load(os.path.join(ROOT, "checkpoints", "raft-sintel.pth")).items()}
    xp = raft_net.pad(rs)
    flow = raft_net.forward(sd_raft, xp[:-1], xp[1:], 20)                # padded, never unpadded (extract_i3d.py:172)
    sd_flow = {k: v.to(cuda_device) for k, v in torch.load(os.path.join(ROOT, "checkpoints", "i3d_flow.pt")).items()}
    ref_flow = i3d_net.forward_features(sd_flow, i3d_net.flow_transform(flow))
    rel = float((torch.from_numpy(res['flow']).to(cuda_device) - ref_flow).norm() / ref_flow.norm())
    print("ExtractI3D flow (RAFT -> I3D) vs oracle:", rel)
    # The flow stream passes through the reference's 8-bit quantiser `round(128 + 255/40 f)` (transforms.py:43-51), a
    # discontinuity: on this clip (tiny motion, 3 quantisation levels in use) a flow perturbation of sigma = 1e-4 px --
    # 1.6e-4 of the flow itself -- already moves the oracle's OWN 1024-d feature by 3e-3, and 3e-4 px by 1.6e-2
    # (scripts: DESIGN.md §2).  Parity of this branch is therefore asserted per stage: RAFT flow (test_raft_gpu.py,
    # rel-L2 <= 1e-3) and quantiser + I3D on identical flow (test_i3d_gpu.py::test_i3d_fused_stream_transforms, 2e-4).
    # End to end only gross agreement can be asked for (measured 9.4e-3 with the split-fp16 RAFT, 4e-2 before it):
    assert rel < 0.05


@pytest.mark.skipif(not HAVE, reason="reference checkpoint copies not present (scripts/fetch_checkpoints.py)")
def test_extract_raft_writes_flow(cuda_device, tmp_path):
    from oracle import raft_net
    from video_features_b200.extract.extract_raft import ExtractRAFT
    vid = str(tmp_path / "clip2.mp4")
    _write_video(vid, 6, h=128, w=160)
    out = str(tmp_path / "out")
    ex = ExtractRAFT(_ns(feature_type='raft', video_paths=[vid], output_path=out, tmp_path=str(tmp_path / "tmp"), batch_size=2))
    assert ex(torch.zeros([1], dtype=torch.long, device=cuda_device)) is None
    flow = np.load(os.path.join(out, "raft", "clip2_raft.npy"))
    assert flow.shape == (5, 2, 128, 160) and flow.dtype == np.float64
    import cv2
    cap = cv2.VideoCapture(vid)
    fr = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        fr.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB))
    x = torch.from_numpy(np.stack(fr)).permute(0, 3, 1, 2).float().to(cuda_device)
    sd = {k: v.to(cuda_device) for k, v in torch.load(os.path.join(ROOT, "checkpoints", "raft-sintel.pth")).items()}
    ref = raft_net.forward(sd, x[:-1], x[1:], 20).cpu().numpy()
    rel = np.linalg.norm(flow - ref) / np.linalg.norm(ref)
    print("ExtractRAFT vs oracle:", rel)
    # Decoded (block-compressed) 128x160 frames are the hard case for reduced-precision RAFT: rounding only the conv
    # INPUTS to fp16 gives 7.6e-3 here in a CPU emulation, fp16 weights alone 6e-4 with a 4.6e-3 max-abs outlier.  With
    # every operand carried as a split-fp16 pair the engine measures 5.7e-5 (the fp32 oracle's own thread-count noise
    # on these frames is 2.5e-5).
    assert rel < 5e-4
