"""Ad-hoc timing of the CLIP path (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clip_tower
from video_features_b200.clip_engine import ClipEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256]
sd = clip_tower.synthetic_state_dict(0)
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
    fps = n / ms * 1e3
    print(f"chunk {chunk}: {ms:.3f} ms / {n} frames -> {fps:.0f} frames/s, {fps*clip_tower.FLOP_PER_FRAME/1e12:.1f} TFLOP/s")
    eng.close()
