"""``ExtractRAFT`` -- drop-in for the reference's models/raft/extract_raft.py on the B200 engine.

Same constructor / attributes / ``forward`` (returns None) / ``extract`` surface; output key 'raft' is a float64
``(T-1, 2, H, W)`` array (the reference builds it with ``.tolist()``), saved under ``{output_path}/raft``.  Frames are
read sequentially with OpenCV, converted BGR->RGB (extract_raft.py:133 -- the stand-alone extractor does swap),
optionally resized (``--side_size``, Pillow-exact bilinear on the GPU), and processed in windows of batch_size+1 frames
with the last frame carried over; padding to /8 and unpadding happen inside the engine.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch
from tqdm import tqdm

from .. import ops
from .._lib import VF_FILTER_BILINEAR
from ..raft_engine import RAFTEngine
from ..utils import action_on_extraction, form_list_from_user_input
from .extract_i3d import load_checkpoint


class ExtractRAFT(torch.nn.Module):

    def __init__(self, args):
        super(ExtractRAFT, self).__init__()
        self.feature_type = args.feature_type
        self.path_list = form_list_from_user_input(args)
        self.batch_size = args.batch_size
        self.extraction_fps = args.extraction_fps
        self.resize_to_smaller_edge = args.resize_to_smaller_edge
        self.side_size = args.side_size
        self.show_pred = args.show_pred
        self.keep_tmp_files = args.keep_tmp_files
        self.on_extraction = args.on_extraction
        self.tmp_path = os.path.join(args.tmp_path, self.feature_type)
        self.output_path = os.path.join(args.output_path, self.feature_type)
        self.progress = tqdm(total=len(self.path_list))
        if self.extraction_fps is not None:
            raise NotImplementedError("extraction_fps re-encodes with ffmpeg (outside the rebuilt path, SURVEY.md §2)")
        self._engines: Dict[tuple, RAFTEngine] = {}

    def forward(self, indices: torch.LongTensor):
        device = indices.device
        if device.type != 'cuda':
            raise RuntimeError("the B200 engine has no CPU path: pass indices on a CUDA device")
        for idx in indices:
            try:
                feats_dict = self.extract(device, None, self.path_list[idx])
                action_on_extraction(feats_dict, self.path_list[idx], self.output_path, self.on_extraction)
            except KeyboardInterrupt:
                raise KeyboardInterrupt
            except Exception as e:
                print(e)
                print(f'Extraction failed at: {self.path_list[idx]} with error (↑). Continuing extraction')
            self.progress.update()

    def _engine(self, device: torch.device, h: int, w: int) -> RAFTEngine:
        key = (device.index or 0, h, w)
        if key not in self._engines:
            self._engines[key] = RAFTEngine(load_checkpoint('raft'), key[0], max_frames=self.batch_size + 1, max_h=h, max_w=w)
        return self._engines[key]

    def extract(self, device, model, video_path=None) -> Dict[str, np.ndarray]:
        import cv2
        cap = cv2.VideoCapture(video_path)
        fps = cap.get(cv2.CAP_PROP_FPS)
        timestamps_ms, batch, flow_frames = [], [], []
        first_frame = True

        def run(batch):
            x = torch.from_numpy(np.stack(batch)).to(device)                     # (B+1, H, W, 3) uint8 RGB
            if self.side_size is not None:
                oh, ow = ops.resize_geometry(x.shape[1], x.shape[2], self.side_size, self.resize_to_smaller_edge)
                if (oh, ow) != tuple(x.shape[1:3]):
                    x = torch.ops.vfeat.resize_u8(x, oh, ow, VF_FILTER_BILINEAR)
            eng = self._engine(device, x.shape[1], x.shape[2])
            flow = eng.flow(x, iters=20, unpad=True)
            flow_frames.extend(flow.cpu().tolist())

        while cap.isOpened():
            frame_exists, bgr = cap.read()
            if first_frame:
                first_frame = False
                if frame_exists is False:
                    continue
            if frame_exists:
                timestamps_ms.append(cap.get(cv2.CAP_PROP_POS_MSEC))
                batch.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
                if len(batch) - 1 == self.batch_size:
                    run(batch)
                    batch = [batch[-1]]
            else:
                if len(batch) > 1:
                    run(batch)
                cap.release()
                break
        return {self.feature_type: np.array(flow_frames), 'fps': np.array(fps), 'timestamps_ms': np.array(timestamps_ms)}
