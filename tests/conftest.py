import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    try:
        import torch
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    except Exception:
        pass


def pytest_sessionstart(session):
    """The real-checkpoint parity tests never skip: in the build container the reference's vendored weights are copied
    into ./checkpoints/ on demand (the GPU box receives that folder with the snapshot)."""
    need = [os.path.join(ROOT, "checkpoints", n) for n in ("i3d_rgb.pt", "i3d_flow.pt", "raft-sintel.pth")]
    if not all(os.path.exists(n) for n in need) and os.path.isdir("/root/reference/models"):
        import subprocess
        subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "fetch_checkpoints.py")])


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", 0)


def rel_l2(a, b):
    import torch
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
