"""One warm-up + one measured encode of N frames, for ncu launch lists (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_features_b200 import synthetic_weights
from video_features_b200.clip_engine import ClipEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 120
sd = synthetic_weights.clip_vit_b32_state_dict(0)
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
eng = ClipEngine(sd, 0, chunk)
eng.encode_frames_u8(frames); torch.cuda.synchronize()
eng.encode_frames_u8(frames); torch.cuda.synchronize()
