"""Development aid: locate a geometry bug in the RAFT lookup output (shift search against the oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import raft_net as R
from video_features_b200.raft_engine import RAFTEngine
sd = torch.load("checkpoints/raft-sintel.pth", map_location="cpu")
sdg = {k: v.cuda() for k, v in R._strip(sd).items()}
eng = RAFTEngine(sd, 0, max_frames=5, max_h=272, max_w=480)
x = R.synthetic_frames(3, 128, 160, seed=128).cuda()
eng.flow(x, iters=1, unpad=False)
img = 2 * (x / 255.0) - 1.0
fmap = R.encoder(sdg, "fnet", img, "instance")
pyr = R.corr_pyramid(fmap[:-1].float(), fmap[1:].float())
ys, xs = torch.meshgrid(torch.arange(16), torch.arange(20), indexing="ij")
coords0 = torch.stack([xs, ys], 0).float()[None].repeat(2, 1, 1, 1).cuda()
look = R.corr_lookup(pyr, coords0)
got = eng.debug_read(4)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("nonzero fraction", float((got != 0).float().mean()), "got norm", float(got.norm()), "ref norm", float(look.norm()))
for b in range(2):
    print("pair", b, rel(got[b], look[b]), "vs other pair", rel(got[b], look[1 - b]))
best = []
for dy in range(-3, 4):
    for dx in range(-3, 4):
        r = torch.roll(look, (dy, dx), (2, 3))
        best.append((rel(got[:, :, 4:12, 4:16], r[:, :, 4:12, 4:16]), dy, dx))
print(sorted(best)[:3])
for c in (0, 40, 80, 81, 200, 323):
    print("chan", c, rel(got[:, c], look[:, c]))
print("got[0,:6,5,5]", got[0, :6, 5, 5].tolist(), "ref", look[0, :6, 5, 5].tolist())
