"""``ExtractCLIP`` -- drop-in for the reference's models/CLIP/extract_clip.py on the B200 engine.

Same constructor, attributes, ``forward(indices)`` / ``extract(...)`` signatures, dict keys and file naming.  What
changes underneath: ``clip.load`` is replaced by a ``ClipEngine`` (device weights + workspace), and ``preprocess`` +
``model.encode_image`` by ONE call into libvfeat.so that takes the decoder's raw uint8 frames and runs the
Pillow-exact bicubic resize, centre crop, normalisation and the ViT-B/32 tower on the GPU.

Differences a user can observe, all deliberate:
  * features are float32 on the GPU as well (the reference's GPU path returns float16 because ``clip.load`` keeps the
    model in half precision on CUDA; its ``--cpu`` path returns float32);
  * a CPU device is refused (no CPU fallback);
  * weights come from a local checkpoint (``$VF_CLIP_CKPT`` or ``~/.cache/clip/ViT-B-32.pt``), never the network.
"""
from __future__ import annotations

import os
import pathlib
import traceback
from typing import Dict

import numpy as np
import torch
from tqdm import tqdm

from ..clip_engine import ClipEngine
from ..utils import AsyncSink, action_on_extraction, already_extracted, extract_frames, form_list_from_user_input

_CKPT_NAMES = {'CLIP-ViT-B/32': 'ViT-B-32.pt', 'CLIP4CLIP-ViT-B-32': 'CLIP4CLIP-ViT-B-32.pth'}


def load_clip_state_dict(feature_type: str) -> Dict[str, torch.Tensor]:
    """Checkpoint in openai's format: a TorchScript archive (what ``clip.load`` downloads) or a plain state dict.
    ``VF_CLIP_SYNTHETIC=<seed>`` selects seeded synthetic weights (benchmarks / tests without the real file)."""
    if os.environ.get("VF_CLIP_SYNTHETIC") is not None:
        from oracle import clip_tower            # weight generator only
        return clip_tower.synthetic_state_dict(int(os.environ["VF_CLIP_SYNTHETIC"] or 0))
    name = _CKPT_NAMES[feature_type]
    cands = [os.environ.get("VF_CLIP_CKPT"), os.path.join(pathlib.Path(__file__).parent, 'checkpoints', name),
             os.path.expanduser(os.path.join("~/.cache/clip", name))]
    for p in cands:
        if p and os.path.exists(p):
            try:
                return torch.jit.load(p, map_location="cpu").state_dict()
            except RuntimeError:
                sd = torch.load(p, map_location="cpu")
                return sd.get("state_dict", sd)
    if feature_type == 'CLIP4CLIP-ViT-B-32':
        raise ValueError(cands[1])                 # extract_clip.py:57-58
    raise FileNotFoundError(f"CLIP checkpoint {name} not found (looked at {[c for c in cands if c]}); "
                            "there is no network access -- set VF_CLIP_CKPT")


class ExtractCLIP(torch.nn.Module):

    def __init__(self, args, external_call=False):
        super().__init__()
        for name in ('feature_type', 'extraction_fps', 'extract_method', 'on_extraction'):
            setattr(self, name, getattr(args, name))
        self.path_list = form_list_from_user_input(args)
        self.external_call = external_call
        if not external_call:
            # --output_direct writes <output_path>/<stem>.npy; otherwise a per-feature sub-folder (which cannot exist for
            # 'CLIP-ViT-B/32': the '/' in the key -- reference quirk)
            self.output_direct = args.output_direct
            self.output_path = args.output_path if self.output_direct is True else os.path.join(args.output_path, self.feature_type)
        self.progress = tqdm(total=len(self.path_list))
        self._engines: Dict[int, ClipEngine] = {}

    def _engine(self, device: torch.device) -> ClipEngine:
        if device.type != 'cuda':
            raise RuntimeError("the B200 engine has no CPU path: pass indices on a CUDA device "
                               "(the reference's --cpu flow is timed by bench.py --impl reference)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._engines:
            if self.feature_type not in _CKPT_NAMES:
                # the reference lists B/16 and the ResNet towers as well (extract_clip.py:46-64); only the
                # ViT-B/32 tower is built here (north_star)
                raise NotImplementedError(self.feature_type)
            self._engines[idx] = ClipEngine(load_clip_state_dict(self.feature_type), device=idx)
        return self._engines[idx]

    def forward(self, indices: torch.LongTensor):
        """indices {torch.LongTensor} -- indices to self.path_list; the device is taken from ``indices.device``."""
        device = indices.device
        model = self._engine(device)          # one engine per device, kept across calls
        collected = []
        # opt-in extras beyond the reference: VF_ASYNC_SINK=1 saves from a writer thread,
        # VF_RESUME=1 skips videos whose output files already exist
        saving = not self.external_call
        sink = AsyncSink() if saving and os.environ.get("VF_ASYNC_SINK") == "1" else None
        resume = saving and os.environ.get("VF_RESUME") == "1"
        try:
            for idx in indices:
                video = self.path_list[idx]
                try:                                # per-video catch-print-continue (extract_clip.py:71-84)
                    if resume and already_extracted([self.feature_type], video, self.output_path, self.on_extraction,
                                                    self.output_direct):
                        self.progress.update()
                        continue
                    feats = self.extract(device, model, None, video)
                    if self.external_call:
                        collected.append(feats)
                    elif sink is not None:
                        sink.submit(feats, video, self.output_path, self.on_extraction, self.output_direct)
                    else:
                        action_on_extraction(feats, video, self.output_path, self.on_extraction, self.output_direct)
                except KeyboardInterrupt:
                    raise
                except Exception as err:
                    print(err)
                    print(f'Extraction failed at: {video} with error (↑). Continuing extraction')
                    traceback.print_exc()
                self.progress.update()
        finally:
            if sink is not None:
                sink.close()
        return collected

    def extract(self, device: torch.device, model: ClipEngine, preprocess_func=None, video_path=None):
        """-> {feature_type: (T,512) float32, 'fps': (), 'timestamps_ms': (T,)}.  ``preprocess_func`` is accepted for
        signature compatibility; the transform is fused into the engine call."""
        decoded, fps, stamps = extract_frames(str(video_path), self.extract_method)
        decoded = [f for f in decoded if f is not None]
        if not decoded:
            raise RuntimeError(f"no frames decoded from {video_path}")
        batch = torch.from_numpy(np.stack(decoded))         # (T,H,W,3) uint8, decoder channel order untouched
        feats = model.encode_frames_u8_host(batch)          # H2D + transform + tower + D2H
        return {self.feature_type: feats.numpy(), 'fps': np.array(fps), 'timestamps_ms': np.array(stamps)}
