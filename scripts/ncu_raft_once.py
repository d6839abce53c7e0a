"""ncu target: two RAFT calls (8 pairs of 270x480 frames, 20 iterations); profile the second one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import raft_net as R
from video_features_b200.raft_engine import RAFTEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9
sd = torch.load("checkpoints/raft-sintel.pth", map_location="cpu")
eng = RAFTEngine(sd, 0, max_frames=n, max_h=272, max_w=480)
x = R.synthetic_frames(n, 270, 480, seed=3).cuda()
for _ in range(2):
    y = eng.flow(x, iters=20, unpad=True)
torch.cuda.synchronize()
print("launches", eng.launch_count, float(y.abs().mean()))
