"""Development aid: determinism + small-geometry check of the RAFT engine (run under compute-sanitizer on the box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import raft_net as R
from video_features_b200.raft_engine import RAFTEngine
sd = torch.load("checkpoints/raft-sintel.pth", map_location="cpu")
h, w, n, it = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (96, 128, 3, 2))]
eng = RAFTEngine(sd, 0, max_frames=n, max_h=h, max_w=w)
x = R.synthetic_frames(n, h, w, seed=11, shift=(0.8, 0.5)).cuda()
y1 = eng.flow(x, iters=it, unpad=True).clone()
y2 = eng.flow(x, iters=it, unpad=True).clone()
print("deterministic:", bool(torch.equal(y1, y2)), float((y1 - y2).abs().max()))
if os.environ.get("NO_ORACLE") is None:
    xp = R.pad(x)
    sdg = {k: v.cuda() for k, v in sd.items()}
    ref = R.unpad(R.forward(sdg, xp[:-1], xp[1:], it), h, w)
    print("rel vs oracle:", float((y1 - ref).norm() / ref.norm()), "max|ref|", float(ref.abs().max()))
