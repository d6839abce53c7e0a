"""Development aid: per-stage checksums / errors of the RAFT engine vs the oracle (run several processes in a row)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
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
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("fmap", rel(eng.debug_read(0), fmap))
rows = eng.debug_read(5)[:, :, 0, :]          # (n, P, ld)
P = 320
off = 0
for l, (hh, ww) in enumerate([(16, 20), (8, 10), (4, 5), (2, 2)]):
    got = rows[:, :, off:off + hh * ww].reshape(2 * P, 1, hh, ww)
    print(f"pyr level {l}", rel(got, pyr[l]), "nan" if not torch.isfinite(got).all() else "")
    off += hh * ww
ys, xs = torch.meshgrid(torch.arange(16), torch.arange(20), indexing="ij")
coords0 = torch.stack([xs, ys], 0).float()[None].repeat(2, 1, 1, 1).cuda()
look = R.corr_lookup(pyr, coords0)
got = eng.debug_read(4)
print("lookup", rel(got, look))
for l in range(4):
    print(f"  lookup level {l}", rel(got[:, l * 81:(l + 1) * 81], look[:, l * 81:(l + 1) * 81]))
got0 = rows[:, :, :P].reshape(2, P, P)
ref0 = pyr[0].reshape(2, P, P)
d = (got0 - ref0).abs()
bad = d > 0.05 * ref0.abs().max()
print("bad entries:", int(bad.sum()), "of", bad.numel())
if bad.any():
    for b in range(2):
        r = bad[b].sum(1).nonzero().flatten().tolist(); c = bad[b].sum(0).nonzero().flatten().tolist()
        print(f" pair {b}: bad rows {r[:12]}{'...' if len(r) > 12 else ''} (n={len(r)}); bad cols {c[:12]}{'...' if len(c) > 12 else ''} (n={len(c)})")
        if r:
            rr = r[0]
            print("   row", rr, "got", [round(float(v), 3) for v in got0[b, rr, :6]], "ref", [round(float(v), 3) for v in ref0[b, rr, :6]],
                  "| got cols 256..261", [round(float(v), 3) for v in got0[b, rr, 256:262]], "ref", [round(float(v), 3) for v in ref0[b, rr, 256:262]])
