"""Shared test helpers (weights files, module trees).  Test infrastructure only."""
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT_DIR = os.path.join(ROOT, "checkpoints")


def checkpoint(name: str) -> str:
    """Path of a reference checkpoint copy (checkpoints/ is git-ignored; scripts/fetch_checkpoints.py fills it from
    /root/reference in the build container and it travels to the GPU box).  A missing file FAILS the test: the
    real-checkpoint parity tests must never go green by skipping."""
    p = os.path.join(CKPT_DIR, name)
    if not os.path.exists(p):
        pytest.fail(f"{p} is missing: run `python scripts/fetch_checkpoints.py` (copies the reference's vendored "
                    "weights) before the -m gpu suite; these parity tests do not skip")
    return p


class _Node(torch.nn.Module):
    def forward(self, x):
        return x


def module_tree(state_dict) -> torch.nn.Module:
    """A module whose parameters carry exactly the dotted names of `state_dict` (for TorchScript archives shaped like
    the ones `clip.load` downloads)."""
    root = _Node()
    for key, value in state_dict.items():
        parts = key.split('.')
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, _Node())
            m = getattr(m, p)
        m.register_parameter(parts[-1], torch.nn.Parameter(value.clone(), requires_grad=False))
    return root
