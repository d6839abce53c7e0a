"""Two Pillow-exact resizes of 256 frames 240x320 -> 224x298 (bicubic: config 1's geometry), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_features_b200  # noqa: F401
from video_features_b200._lib import VF_FILTER_BICUBIC
x = torch.randint(0, 256, (256, 240, 320, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.ops.vfeat.resize_u8(x, 224, 298, VF_FILTER_BICUBIC)
torch.cuda.synchronize()
