"""Ad-hoc timing of the I3D path (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import i3d_net
from video_features_b200.i3d_engine import I3DEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mod = sys.argv[2] if len(sys.argv) > 2 else "rgb"
cin = 3 if mod == "rgb" else 2
sd = i3d_net.synthetic_state_dict(mod, 0)
eng = I3DEngine(sd, mod, 0, max_stacks=n, max_T=64)
x = torch.rand(n, cin, 64, 224, 224, device="cuda") * 2 - 1
for _ in range(3): eng(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gf = 222.30 if mod == "rgb" else 204.68
print(f"{mod} n={n} T=64: {ms:.3f} ms -> {n/ms*1e3:.1f} stacks/s, {n*64/ms*1e3:.0f} frames/s, {n*gf/ms:.1f} TFLOP/s algorithmic (fast={os.environ.get('VF_I3D_FAST','0')})")
