"""Time of the fused QKV + attention kernel on a 250-frame chunk (CUDA events, 50 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_features_b200 import synthetic_weights
from video_features_b200.clip_engine import ClipEngine
eng = ClipEngine(synthetic_weights.clip_vit_b32_state_dict(0), device=0)
x = torch.randn(250 * 50, 768, device="cuda").half()
for _ in range(5):
    eng.block_attention(3, x, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    eng.block_attention(3, x, True)
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('VF_TAG', '')} fused QKV+attention, 250 frames: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch")
