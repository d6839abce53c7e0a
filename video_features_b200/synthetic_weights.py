"""Seeded synthetic weights in the layouts the engines load -- for benchmarks and tests on machines without the real
checkpoints (the reference downloads ViT-B-32.pt at run time, models/CLIP/extract_clip.py:47; there is no network
here).  ``VF_CLIP_SYNTHETIC=<seed>[:outliers]`` makes ``ExtractCLIP`` use them instead of a checkpoint file.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

WIDTH, LAYERS, TOKENS, PATCH, MLP, EMBED = 768, 12, 50, 32, 3072, 512


def clip_vit_b16_state_dict(seed: int = 0, outliers: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """The same for the ViT-B/16 tower (16-pixel patches, 197 tokens)."""
    return clip_vit_b32_state_dict(seed, outliers, patch=16)


def clip_vit_b32_state_dict(seed: int = 0, outliers: bool = False, patch: int = PATCH) -> "OrderedDict[str, torch.Tensor]":
    """openai ``visual.*`` key layout, fp32.  Initialisation scales are the ones openai/CLIP's
    ``initialize_parameters`` uses (width**-0.5 etc.), with perturbed LayerNorm gains / biases so that every term of the
    forward matters.

    ``outliers=True`` adds what trained ViT-B/32 weights have and random ones lack: a handful of residual-stream
    channels that carry magnitudes of 50-200 through every block (set up by the positional embedding and fed by the
    projection biases), LayerNorm gains with a heavy tail (a few entries near 0.05, a few above 4), and a few large
    rows in the attention / MLP output projections.  This is the regime where fp16 storage of intermediate tensors is
    most at risk."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    scale = WIDTH ** -0.5
    attn_std = WIDTH ** -0.5
    proj_std = (WIDTH ** -0.5) * ((2 * LAYERS) ** -0.5)
    fc_std = (2 * WIDTH) ** -0.5
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["visual.class_embedding"] = rn(WIDTH, std=scale)
    sd["visual.positional_embedding"] = rn((224 // patch) ** 2 + 1, WIDTH, std=scale)
    sd["visual.proj"] = rn(WIDTH, EMBED, std=scale)
    sd["visual.conv1.weight"] = rn(WIDTH, 3, patch, patch, std=(3 * patch * patch) ** -0.5)
    for name in ("ln_pre", "ln_post"):
        sd[f"visual.{name}.weight"] = 1.0 + rn(WIDTH, std=0.1)
        sd[f"visual.{name}.bias"] = rn(WIDTH, std=0.05)
    for i in range(LAYERS):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "attn.in_proj_weight"] = rn(3 * WIDTH, WIDTH, std=attn_std)
        sd[p + "attn.in_proj_bias"] = rn(3 * WIDTH, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(WIDTH, WIDTH, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(WIDTH, std=0.02)
        sd[p + "ln_1.weight"] = 1.0 + rn(WIDTH, std=0.1)
        sd[p + "ln_1.bias"] = rn(WIDTH, std=0.05)
        sd[p + "mlp.c_fc.weight"] = rn(MLP, WIDTH, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(MLP, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(WIDTH, MLP, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(WIDTH, std=0.02)
        sd[p + "ln_2.weight"] = 1.0 + rn(WIDTH, std=0.1)
        sd[p + "ln_2.bias"] = rn(WIDTH, std=0.05)
    if outliers:
        dims = torch.randperm(WIDTH, generator=g)[:6]
        sign = torch.tensor([1.0, -1.0, 1.0, -1.0, 1.0, -1.0])
        # ln_pre rescales whatever the embedding holds, so the outlier channels are planted in ln_pre's affine part:
        # the residual stream leaves ln_pre with +-(60..150) in these channels
        sd["visual.ln_pre.bias"][dims] = sign * (60.0 + 90.0 * torch.rand(6, generator=g))
        sd["visual.ln_pre.weight"][dims] = 8.0
        for i in range(LAYERS):
            p = f"visual.transformer.resblocks.{i}."
            for ln in ("ln_1", "ln_2"):
                w = sd[p + ln + ".weight"]
                w.mul_(torch.exp(rn(WIDTH, std=0.35)))                       # heavy-tailed gains
                w[dims[:3]] = 0.05 + 0.05 * torch.rand(3, generator=g)       # trained nets damp their outlier channels ...
                w[dims[3:]] = 3.0 + 2.0 * torch.rand(3, generator=g)         # ... or read them loudly
                sd[p + ln + ".bias"][dims] = rn(6, std=0.5)
            # the projections keep feeding the outlier channels (bias) and have a few loud rows
            sd[p + "attn.out_proj.bias"][dims] = sign * (1.0 + 2.0 * torch.rand(6, generator=g))
            sd[p + "mlp.c_proj.bias"][dims] = sign * (2.0 + 4.0 * torch.rand(6, generator=g))
            loud = torch.randperm(WIDTH, generator=g)[:4]
            sd[p + "attn.out_proj.weight"][loud] *= 6.0
            sd[p + "mlp.c_proj.weight"][loud] *= 6.0
        sd["visual.ln_post.weight"][dims] = 0.05
    return sd


def parse_env(value: str):
    """``VF_CLIP_SYNTHETIC`` value -> (seed, outliers): "3", "3:outliers", "" (seed 0)."""
    head, _, tail = (value or "").partition(":")
    return int(head or 0), tail.strip().lower() == "outliers"
