"""ViT-B/16 tower throughput (device-resident frames) for a few chunk sizes, with the per-category device time split
(development aid; the judged CLIP numbers come from bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_features_b200 import synthetic_weights
from video_features_b200.clip_engine import ClipEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1008
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [63]
sd = synthetic_weights.clip_vit_b16_state_dict(0)
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
for chunk in chunks:
    eng = ClipEngine(sd, 0, chunk)
    for _ in range(3):
        eng.encode_frames_u8(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.encode_frames_u8(frames)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    eng.profile(True)
    eng.encode_frames_u8(frames); torch.cuda.synchronize()
    gms, gl, gf = eng.profile_read()
    cats = eng.profile_categories()
    eng.profile(False)
    print(f"chunk {chunk}: {n / ms * 1e3:.0f} frames/s  ({ms:.2f} ms / {n} frames); eager split ms {cats}; "
          f"gemm {gf / gms / 1e9:.0f} TFLOP/s over {gl} launches", flush=True)
    eng.close()
