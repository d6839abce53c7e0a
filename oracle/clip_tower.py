"""fp32 CPU restatement of the CLIP ViT image tower -- test oracle only.

The reference runs ``model.encode_image(frames)`` (models/CLIP/extract_clip.py:128)
on a model returned by ``clip.load("ViT-B/32")`` (extract_clip.py:47).  ``clip`` is
openai/CLIP (third-party, un-vendored, version unpinned by the reference; weights
fetched at run time) -- absent offline.  This file restates the published algorithm
of ``clip/model.py``: ``VisionTransformer.forward``, ``ResidualAttentionBlock``,
``LayerNorm`` (fp32 compute), ``QuickGELU`` and ``CLIP.encode_image``; the state-dict
keys are openai's (``visual.*``) so a user-supplied real checkpoint loads unchanged.

PARITY UNPINNED versus the reference itself (no reference test or golden vector
touches this boundary and neither the package nor its weights can be obtained here).
It IS pinned against an independent implementation of the same math: HF
``transformers.CLIPVisionModelWithProjection`` (tests/test_oracle_clip.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F

# ViT-B/32 hyper-parameters (clip/model.py: build_model for "ViT-B/32")
WIDTH = 768
LAYERS = 12
HEADS = 12
PATCH = 32
RES = 224
GRID = RES // PATCH            # 7
TOKENS = GRID * GRID + 1       # 50
MLP = 4 * WIDTH                # 3072
EMBED = 512
LN_EPS = 1e-5


def synthetic_state_dict(seed: int = 0, dtype=torch.float32, patch: int = PATCH) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights in openai's ``visual.*`` key layout.

    Scales follow clip/model.py (VisionTransformer.__init__: ``scale = width**-0.5``
    for class/positional embedding and proj; CLIP.initialize_parameters:
    attn_std = width**-0.5, proj_std = width**-0.5 * (2*layers)**-0.5,
    fc_std = (2*width)**-0.5) so activations have realistic magnitudes.  LayerNorm
    gains/biases and linear biases are perturbed so every term of the forward is
    exercised by parity tests.
    """
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    scale = WIDTH ** -0.5
    attn_std = WIDTH ** -0.5
    proj_std = (WIDTH ** -0.5) * ((2 * LAYERS) ** -0.5)
    fc_std = (2 * WIDTH) ** -0.5
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["visual.class_embedding"] = rn(WIDTH, std=scale)
    sd["visual.positional_embedding"] = rn((RES // patch) ** 2 + 1, WIDTH, std=scale)
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
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # clip/model.py LayerNorm: compute in fp32, cast back to the input dtype.
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), LN_EPS).to(x.dtype)


def _quick_gelu(x: torch.Tensor) -> torch.Tensor:
    # clip/model.py QuickGELU
    return x * torch.sigmoid(1.702 * x)


def _attention(x: torch.Tensor, w_in, b_in, w_out, b_out) -> torch.Tensor:
    """nn.MultiheadAttention(width, heads) self-attention, no mask, batch-first here.

    q is scaled by head_dim**-0.5 before q@k^T (torch MHA semantics, what
    ResidualAttentionBlock.attention calls with need_weights=False).
    """
    B, S, D = x.shape
    hd = D // HEADS
    qkv = F.linear(x, w_in, b_in)                      # (B,S,3D) = cat(q,k,v)
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, HEADS, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.view(B, S, HEADS, hd).transpose(1, 2)
    v = v.view(B, S, HEADS, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, D)
    return F.linear(o, w_out, b_out)


@torch.no_grad()
def encode_image(sd: Dict[str, torch.Tensor], frames: torch.Tensor, *, return_hidden: bool = False):
    """``CLIP.encode_image`` == ``VisionTransformer.forward``.

    frames: (B,3,224,224) float, already normalised (output of the CLIP transform).
    returns (B,512) in the weight dtype.  No L2 normalisation (the reference saves
    the raw projection, extract_clip.py:128-131).
    """
    w = sd["visual.conv1.weight"]
    x = frames.to(w.dtype)
    patch = w.shape[-1]                                          # 32 (ViT-B/32) or 16 (ViT-B/16): conv stride == kernel
    x = F.conv2d(x, w, None, stride=patch)                       # (B,768,7,7) / (B,768,14,14)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)   # (B,49|196,768) row-major grid
    cls = sd["visual.class_embedding"].to(x.dtype).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"].to(x.dtype)
    x = _ln(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    hidden = [x]
    for i in range(LAYERS):
        p = f"visual.transformer.resblocks.{i}."
        h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        x = x + _attention(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
                           sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        h = _quick_gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        hidden.append(x)
    x = _ln(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    out = x @ sd["visual.proj"]
    return (out, hidden) if return_hidden else out


def to_hf_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Remap openai ``visual.*`` keys to HF CLIPVisionModelWithProjection keys
     -- used only to pin this oracle against HF's code."""
    out = {}
    out["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    out["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    out["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    out["vision_model.pre_layrnorm.weight"] = sd["visual.ln_pre.weight"]
    out["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.bias"]
    out["vision_model.post_layernorm.weight"] = sd["visual.ln_post.weight"]
    out["vision_model.post_layernorm.bias"] = sd["visual.ln_post.bias"]
    out["visual_projection.weight"] = sd["visual.proj"].t().contiguous()
    for i in range(LAYERS):
        p = f"visual.transformer.resblocks.{i}."
        h = f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[p + "attn.in_proj_weight"].split(WIDTH, dim=0)
        bq, bk, bv = sd[p + "attn.in_proj_bias"].split(WIDTH, dim=0)
        for n, wt, bs in (("q", wq, bq), ("k", wk, bk), ("v", wv, bv)):
            out[h + f"self_attn.{n}_proj.weight"] = wt
            out[h + f"self_attn.{n}_proj.bias"] = bs
        out[h + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        out[h + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        out[h + "layer_norm1.weight"] = sd[p + "ln_1.weight"]
        out[h + "layer_norm1.bias"] = sd[p + "ln_1.bias"]
        out[h + "layer_norm2.weight"] = sd[p + "ln_2.weight"]
        out[h + "layer_norm2.bias"] = sd[p + "ln_2.bias"]
        out[h + "mlp.fc1.weight"] = sd[p + "mlp.c_fc.weight"]
        out[h + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.bias"]
        out[h + "mlp.fc2.weight"] = sd[p + "mlp.c_proj.weight"]
        out[h + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.bias"]
    return out


FLOP_PER_FRAME = 231_211_008 + 12 * 715_468_800 + 786_432   # SURVEY.md App. E (8.818 GFLOP)
