"""Checkpoint loading of ExtractCLIP (reference: `clip.load` at models/CLIP/extract_clip.py:47,60 accepts a model name
-> downloaded TorchScript archive, or a path to an archive / state dict) and the rule that the product package never
touches the oracle.  CPU only."""
import os
import re

import pytest
import torch

from helpers import ROOT, module_tree
from video_features_b200 import synthetic_weights
from video_features_b200.extract import extract_clip


@pytest.fixture(scope="module")
def sd():
    return synthetic_weights.clip_vit_b32_state_dict(3)


def _same(a, b):
    return set(a) >= set(b) and all(torch.equal(a[k], b[k]) for k in b)


def test_jit_archive_round_trip(tmp_path, sd, monkeypatch):
    """What `clip.load("ViT-B/32")` caches in ~/.cache/clip is a TorchScript archive: its state_dict() must come back
    with openai's `visual.*` keys, bit for bit."""
    full = dict(sd)
    full["logit_scale"] = torch.tensor(4.6)                                   # text-tower / head keys are carried along
    full["transformer.resblocks.0.ln_1.weight"] = torch.ones(512)
    p = str(tmp_path / "ViT-B-32.pt")
    torch.jit.save(torch.jit.script(module_tree(full)), p)
    got = extract_clip.read_clip_checkpoint(p)
    assert _same(got, sd) and "logit_scale" in got
    monkeypatch.delenv("VF_CLIP_SYNTHETIC", raising=False)
    monkeypatch.setenv("VF_CLIP_CKPT", p)
    assert _same(extract_clip.load_clip_state_dict('CLIP-ViT-B/32'), sd)


def test_plain_nested_and_clip4clip_state_dicts(tmp_path, sd):
    p1, p2, p3 = (str(tmp_path / n) for n in ("plain.pt", "nested.pt", "c4c.pth"))
    torch.save(dict(sd), p1)
    torch.save({"state_dict": dict(sd), "epoch": 3}, p2)
    torch.save({"clip." + k: v for k, v in sd.items()} | {"sim_header.w": torch.zeros(2)}, p3)     # CLIP4Clip layout
    for p in (p1, p2, p3):
        assert _same(extract_clip.read_clip_checkpoint(p), sd), p


def test_missing_checkpoint_errors_like_the_reference(monkeypatch, tmp_path):
    monkeypatch.delenv("VF_CLIP_SYNTHETIC", raising=False)
    monkeypatch.setenv("VF_CLIP_CKPT", str(tmp_path / "nope.pt"))
    monkeypatch.setenv("HOME", str(tmp_path))
    with pytest.raises(ValueError):                                           # extract_clip.py:57-58
        extract_clip.load_clip_state_dict('CLIP4CLIP-ViT-B-32')
    with pytest.raises(FileNotFoundError, match="no network"):
        extract_clip.load_clip_state_dict('CLIP-ViT-B/32')


def test_synthetic_env_selects_seed_and_outliers(monkeypatch):
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "3")
    a = extract_clip.load_clip_state_dict('CLIP-ViT-B/32')
    monkeypatch.setenv("VF_CLIP_SYNTHETIC", "3:outliers")
    b = extract_clip.load_clip_state_dict('CLIP-ViT-B/32')
    assert _same(a, synthetic_weights.clip_vit_b32_state_dict(3))
    assert float(b["visual.ln_pre.bias"].abs().max()) > 50 and float(a["visual.ln_pre.bias"].abs().max()) < 1


def test_product_package_never_imports_the_oracle():
    """oracle/ is checker infrastructure: nothing under video_features_b200/ (nor main.py) may import or name it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|\boracle\s*\.|importlib[^\n]*oracle", re.M)
    bad = []
    targets = [os.path.join(ROOT, "main.py")]
    for d, _, files in os.walk(os.path.join(ROOT, "video_features_b200")):
        targets += [os.path.join(d, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h"))]
    for t in targets:
        if pat.search(open(t).read()):
            bad.append(os.path.relpath(t, ROOT))
    assert not bad, f"product files reference the oracle: {bad}"
