"""ncu target: I3D rgb forwards of 8 stacks x 64 frames (vendored weights); profile the last one.
VF_NO_GRAPH=1 makes the launches visible to ncu as kernels instead of one graph launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_features_b200.i3d_engine import I3DEngine
stream = sys.argv[1] if len(sys.argv) > 1 else "rgb"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = torch.load(f"checkpoints/i3d_{stream}.pt", map_location="cpu")
eng = I3DEngine(sd, stream, 0, max_stacks=S, max_T=64)
if stream == "rgb":
    x = torch.randint(0, 256, (S, 64, 224, 224, 3), dtype=torch.uint8, device="cuda")
    fn = lambda: eng.forward_frames_u8(x)
else:
    x = (torch.rand(S, 2, 64, 224, 224, device="cuda") * 2 - 1)
    fn = lambda: eng(x)
once = os.environ.get('VF_ONCE') == '1'
for _ in range(1 if once else 3):
    y = fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(1 if once else 5):
    y = fn()
e1.record(); torch.cuda.synchronize()
print(f"i3d {stream}: {S * (1 if once else 5) / e0.elapsed_time(e1) * 1e3:.0f} stacks/s, launches {eng.launch_count}", float(y.abs().mean()))
