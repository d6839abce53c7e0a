"""Copies the reference's vendored pretrained weights (data, not source) into ./checkpoints/ (git-ignored, but it
travels to the GPU box with the repo snapshot) so the parity tests can also run on the real weights."""
import os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/models"
dst = os.path.join(ROOT, "checkpoints")
os.makedirs(dst, exist_ok=True)
for rel in ("i3d/checkpoints/i3d_rgb.pt", "i3d/checkpoints/i3d_flow.pt", "raft/checkpoints/raft-sintel.pth"):
    src = os.path.join(REF, rel)
    if os.path.exists(src):
        shutil.copy2(src, os.path.join(dst, os.path.basename(rel)))
        print("copied", rel)
    else:
        print("missing", src, file=sys.stderr)
