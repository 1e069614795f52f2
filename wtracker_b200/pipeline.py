"""The whole per-frame hot path as one object: crop -> YOLOv8s -> decode/NMS -> tracking rows ->
ResMLP position prediction -> bbox error, every stage a kernel of libwtracker_b200.so.

``step_device`` is the device-resident path (frames already in HBM; what ``bench.py`` reports as
``value``); ``step_host`` is the same work entered with HOST buffers (camera views in pinned memory
in, result rows out — what ``bench.py`` reports as ``e2e`` and what the controllers use).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.neural.engine import ResMLPEngine
from wtracker_b200.neural.mlp import WormPredictor


@dataclass
class StepResult:
    boxes: torch.Tensor       # [n, max_det, 6] x1, y1, x2, y2 (view px), conf, anchor
    count: torch.Tensor       # [n] int32
    worm: torch.Tensor        # [n, 4] f64 xywh in frame px (NaN = none)
    mic: torch.Tensor         # [n, 4] f64 microscope box
    pred: torch.Tensor        # [n, 2] f32 ResMLP output (dx, dy of the worm `pred_frames[0]` frames ahead)
    pred_valid: torch.Tensor  # [n] uint8, 0 where the 7-box history had a gap
    bbox_error: torch.Tensor  # [n] f64


class HotPath:
    def __init__(self, state_dict: dict, predictor: WormPredictor, view: int = 640, imgsz: int = 640, batch: int = 64,
                 micro: int = 51, table_rows: int = 1 << 16, device: str = "cuda:0", conf: float = 0.1,
                 iou: float = 0.7, max_det: int = 1):
        self.lib = L.lib()
        self.device = torch.device(device)
        self.batch, self.view, self.micro = batch, view, micro
        self.det = DetectorEngine(state_dict, (view, view), imgsz, batch=batch, conf=conf, iou=iou, max_det=max_det,
                                  device=device)
        self.mlp = ResMLPEngine(predictor, device)
        self.offsets = torch.tensor(predictor.io_config.input_frames, dtype=torch.int32, device=self.device)
        self.k = int(self.offsets.numel())
        d = self.device
        self.table_rows = table_rows
        self.table = torch.full((table_rows, 4), float("nan"), dtype=torch.float64, device=d)   # worm xywh by frame
        self.mic_table = torch.zeros((table_rows, 4), dtype=torch.float64, device=d)
        self.mlp_x = torch.zeros((batch, 4 * self.k), dtype=torch.float32, device=d)
        self.mlp_valid = torch.zeros((batch,), dtype=torch.uint8, device=d)
        self.mlp_y = torch.zeros((batch, 2), dtype=torch.float32, device=d)
        self.err = torch.zeros((batch,), dtype=torch.float64, device=d)
        # host staging for the e2e path
        self.h_views = torch.empty((batch, view, view), dtype=torch.uint8).pin_memory()
        self.d_views = torch.empty((batch, view, view), dtype=torch.uint8, device=d)
        self.h_worm = torch.empty((batch, 4), dtype=torch.float64).pin_memory()
        self.h_boxes = torch.empty((batch, max_det, 6), dtype=torch.float32).pin_memory()
        self.h_count = torch.empty((batch,), dtype=torch.int32).pin_memory()
        self.h_pred = torch.empty((batch, 2), dtype=torch.float32).pin_memory()
        self.h_valid = torch.empty((batch,), dtype=torch.uint8).pin_memory()
        self.h_err = torch.empty((batch,), dtype=torch.float64).pin_memory()
        self._rows = torch.arange(table_rows, dtype=torch.int32, device=d)
        self._iota = self._rows[:batch]
        self._zeros = torch.zeros(batch, dtype=torch.int32, device=d)

    # ------------------------------------------------------------------ device-resident step
    def step_device(self, frames: torch.Tensor, frame_idx: torch.Tensor, crop_x: torch.Tensor, crop_y: torch.Tensor,
                    first_row: int, marks: list | None = None) -> StepResult:
        """frames u8 [F, H, W] on the device; frame_idx / crop_x / crop_y int32 [n]; the n results are
        written to rows [first_row, first_row + n) of the on-device tracking table."""
        n = int(frame_idx.numel())
        assert n <= self.batch and first_row + n <= self.table_rows
        s = torch.cuda.current_stream().cuda_stream
        boxes, count = self.det.detect_crops(frames, frame_idx, crop_x, crop_y, marks)
        worm = self.table[first_row: first_row + n]
        mic = self.mic_table[first_row: first_row + n]
        L.check(self.lib.wt_track_rows(boxes.data_ptr(), count.data_ptr(), self.det.max_det, crop_x.data_ptr(),
                                       crop_y.data_ptr(), self.view, self.view, self.micro, self.micro,
                                       worm.data_ptr(), mic.data_ptr(), n, s), "wt_track_rows")
        rows = self._rows[first_row: first_row + n]
        L.check(self.lib.wt_mlp_gather(self.table.data_ptr(), self.table_rows, rows.data_ptr(),
                                       self.offsets.data_ptr(), self.k, self.mlp_x.data_ptr(),
                                       self.mlp_valid.data_ptr(), n, s), "wt_mlp_gather")
        self.mlp.forward(self.mlp_x[:n], self.mlp_y[:n])
        L.check(self.lib.wt_bbox_error(worm.data_ptr(), mic.data_ptr(), self.err.data_ptr(), n, s), "wt_bbox_error")
        return StepResult(boxes, count, worm, mic, self.mlp_y[:n], self.mlp_valid[:n], self.err[:n])

    # ------------------------------------------------------------------ host-buffer step (public API)
    def step_host(self, views: np.ndarray | torch.Tensor, first_row: int = 0) -> dict[str, np.ndarray]:
        """views: u8 [n, view, view] HOST array of camera views.  Host->device copy of the views and
        device->host copies of every result are part of the call (no torch compute kernels involved).
        Returns host arrays: worm [n,4] f64 xywh (view px, NaN = none), boxes [n,max_det,6], count [n],
        pred [n,2], pred_valid [n], bbox_error [n]."""
        n = views.shape[0]
        assert n <= self.batch
        src = views if torch.is_tensor(views) else torch.from_numpy(views)
        if not src.is_pinned():
            self.h_views[:n].copy_(src)
            src = self.h_views[:n]
        self.d_views[:n].copy_(src, non_blocking=True)
        r = self.step_device(self.d_views, self._iota[:n], self._zeros[:n], self._zeros[:n], first_row)
        self.h_worm[:n].copy_(r.worm, non_blocking=True)
        self.h_boxes[:n].copy_(r.boxes, non_blocking=True)
        self.h_count[:n].copy_(r.count, non_blocking=True)
        self.h_pred[:n].copy_(r.pred, non_blocking=True)
        self.h_valid[:n].copy_(r.pred_valid, non_blocking=True)
        self.h_err[:n].copy_(r.bbox_error, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return dict(worm=self.h_worm[:n].numpy(), boxes=self.h_boxes[:n].numpy(), count=self.h_count[:n].numpy(),
                    pred=self.h_pred[:n].numpy(), pred_valid=self.h_valid[:n].numpy(), bbox_error=self.h_err[:n].numpy())

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.batch * self.view * self.view

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.batch * (4 * 8 + self.det.max_det * 6 * 4 + 4 + 2 * 4 + 1 + 8)
