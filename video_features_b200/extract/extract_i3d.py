"""``ExtractI3D`` -- drop-in for the reference's models/i3d/extract_i3d.py on the B200 engine.

Same constructor / attributes / ``forward(indices)`` / ``extract(...)`` surface, dict keys ('rgb', 'flow', 'fps',
'timestamps_ms'), stack / step logic (stack_size+1 frames per stack, one overlap frame when step == stack) and
float64 ``(n_stacks, 1024)`` outputs.  Underneath, per stack of 65 frames:
  decoder frames (uint8, BGR kept as the reference does) -> GPU Pillow-exact bilinear resize to min side 256
  -> rgb stream : fused crop + 2x/255-1 + phase packing -> I3D (vf_i3d_forward_u8)
  -> flow stream: RAFT on the 64 consecutive pairs, padded and NOT unpadded (extract_i3d.py:172) -> fused crop + clamp
     + 8-bit quantisation + scaling -> I3D (vf_i3d_forward_flow)
``--flow_type pwc`` (the CLI default) and ``flow`` (pre-computed jpgs) are outside the rebuilt path:
they raise NotImplementedError at construction when the flow stream is requested.
"""
from __future__ import annotations

import os
import pathlib
import traceback
from typing import Dict

import numpy as np
import torch
from tqdm import tqdm

from .. import ops
from .._lib import VF_FILTER_BILINEAR
from ..i3d_engine import I3DEngine
from ..raft_engine import RAFTEngine
from ..utils import AsyncSink, already_extracted, VideoReader, action_on_extraction, form_list_from_user_input

PRE_CENTRAL_CROP_MIN_SIDE_SIZE = 256
CENTRAL_CROP_MIN_SIDE_SIZE = 224
DEFAULT_I3D_STEP_SIZE = 64
DEFAULT_I3D_STACK_SIZE = 64
_HERE = pathlib.Path(__file__).resolve().parent
_CKPT_DIRS = [os.environ.get("VF_CKPT_DIR"), str(_HERE / "checkpoints"), str(_HERE.parents[1] / "checkpoints")]
_CKPT = {'rgb': 'i3d_rgb.pt', 'flow': 'i3d_flow.pt', 'raft': 'raft-sintel.pth'}


def load_checkpoint(kind: str) -> Dict[str, torch.Tensor]:
    """The reference's vendored weights (models/i3d/checkpoints/i3d_{rgb,flow}.pt, models/raft/checkpoints/
    raft-sintel.pth), looked up in $VF_CKPT_DIR or ./checkpoints."""
    for d in _CKPT_DIRS:
        if d and os.path.exists(os.path.join(d, _CKPT[kind])):
            return torch.load(os.path.join(d, _CKPT[kind]), map_location="cpu")
    raise FileNotFoundError(f"{_CKPT[kind]} not found in {[d for d in _CKPT_DIRS if d]} (set VF_CKPT_DIR)")


class ExtractI3D(torch.nn.Module):

    def __init__(self, args, external_call=False):
        super(ExtractI3D, self).__init__()
        self.feature_type = args.feature_type
        self.streams = ['rgb', 'flow'] if args.streams is None else args.streams
        self.path_list = form_list_from_user_input(args)
        self.flow_type = args.flow_type
        self.min_side_size = PRE_CENTRAL_CROP_MIN_SIDE_SIZE
        self.central_crop_size = CENTRAL_CROP_MIN_SIDE_SIZE
        self.extraction_fps = args.extraction_fps
        self.step_size = args.step_size if args.step_size is not None else DEFAULT_I3D_STEP_SIZE
        self.stack_size = args.stack_size if args.stack_size is not None else DEFAULT_I3D_STACK_SIZE
        self.show_pred = args.show_pred
        self.keep_tmp_files = args.keep_tmp_files
        self.on_extraction = args.on_extraction
        self.tmp_path = os.path.join(args.tmp_path, self.feature_type)
        self.external_call = external_call
        if external_call is False:
            self.output_direct = args.output_direct
            self.output_path = args.output_path if self.output_direct is True else os.path.join(args.output_path, self.feature_type)
        if 'flow' in self.streams and self.flow_type != 'raft':
            raise NotImplementedError(f"flow_type '{self.flow_type}': only the RAFT flow branch is built on the B200 "
                                      "engine (pass --flow_type raft, or --streams rgb)")
        self.progress = tqdm(total=len(self.path_list))
        self._models: Dict[int, dict] = {}

    def _load(self, device: torch.device) -> dict:
        if device.type != 'cuda':
            raise RuntimeError("the B200 engine has no CPU path: pass indices on a CUDA device")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._models:
            m = {s: I3DEngine(load_checkpoint(s), s, idx, max_stacks=1, max_T=max(self.stack_size, 16))
                 for s in self.streams}
            self._models[idx] = {'i3d': m, 'raft': None}
        return self._models[idx]

    def forward(self, indices: torch.LongTensor):
        device = indices.device
        models = self._load(device)
        feats_list = []
        saving = self.external_call is False           # opt-in extras, see ExtractCLIP.forward
        sink = AsyncSink() if saving and os.environ.get("VF_ASYNC_SINK") == "1" else None
        resume = saving and os.environ.get("VF_RESUME") == "1"
        try:
            for idx in indices:
                try:
                    if resume and already_extracted(self.streams, self.path_list[idx], self.output_path, self.on_extraction):
                        self.progress.update()
                        continue
                    feats_dict = self.extract(device, None, models, self.path_list[idx])
                    if self.external_call is not False:
                        feats_list.append(feats_dict)
                    elif sink is not None:
                        sink.submit(feats_dict, self.path_list[idx], self.output_path, self.on_extraction)
                    else:
                        action_on_extraction(feats_dict, self.path_list[idx], self.output_path, self.on_extraction)
                except KeyboardInterrupt:
                    raise KeyboardInterrupt
                except Exception as e:
                    print(e)
                    print(f'Extraction failed at: {self.path_list[idx]}. Continuing extraction')
                    traceback.print_exc()
                self.progress.update()
        finally:
            if sink is not None:
                sink.close()
        return feats_list

    def _run_on_a_stack(self, feats_dict, frames_u8: torch.Tensor, models: dict, device: torch.device):
        """frames_u8: (stack_size+1, H, W, 3) uint8 host tensor, decoder order (extract_i3d.py:160-193)."""
        x = frames_u8.to(device, non_blocking=True)
        h, w = x.shape[1:3]
        oh, ow = ops.resize_geometry(h, w, self.min_side_size, True)
        if (oh, ow) != (h, w):
            x = torch.ops.vfeat.resize_u8(x, oh, ow, VF_FILTER_BILINEAR)          # ToPILImage -> ResizeImproved(256)
        for stream in self.streams:
            if stream == 'rgb':
                feats = models['i3d']['rgb'].forward_frames_u8(x[:-1][None])        # stack[:-1], crop/scale fused
            elif stream == 'flow':
                if models['raft'] is None:
                    models['raft'] = RAFTEngine(load_checkpoint('raft'), device.index or 0,
                                                max_frames=self.stack_size + 1, max_h=oh, max_w=ow)
                flow = models['raft'].flow(x, iters=20, unpad=False)                # padded flow, as the reference
                feats = models['i3d']['flow'].forward_flow(flow[None])
            else:
                raise NotImplementedError
            feats_dict[stream].extend(feats.cpu().tolist())

    def extract(self, device, flow_xtr_model, models, video_path=None):
        video = VideoReader(str(video_path))
        fps, frame_cnt = video.fps, video.frame_cnt
        mspf = 0.001 / fps                                   # (sic) extract_i3d.py:241
        if self.extraction_fps is not None:
            samples_num = int(frame_cnt / fps * self.extraction_fps)
            samples_ix = np.linspace(1, frame_cnt - 1, samples_num).astype(int)
        elif frame_cnt < DEFAULT_I3D_STACK_SIZE + 1:          # short video: resampled up to 65 frames
            samples_ix = np.linspace(1, frame_cnt - 1, DEFAULT_I3D_STACK_SIZE + 1).astype(int)
        else:
            samples_ix = np.arange(frame_cnt)
        frames, timestamps_ms = [], [i * mspf for i in samples_ix]
        for i in samples_ix:
            f = video.get_frame(int(i)) if int(i) < frame_cnt else None
            if f is not None:
                frames.append(f)
        feats_dict = {stream: [] for stream in self.streams}
        stack = []
        for rgb in frames:
            stack.append(torch.from_numpy(rgb))
            if len(stack) - 1 == self.stack_size:
                self._run_on_a_stack(feats_dict, torch.stack(stack), models, device)
                stack = stack[self.step_size:]
        feats_dict = {stream: np.array(feats) for stream, feats in feats_dict.items()}
        feats_dict['fps'] = np.array(fps)
        feats_dict['timestamps_ms'] = np.array(timestamps_ms)
        return feats_dict
