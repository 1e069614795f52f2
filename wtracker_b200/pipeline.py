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
        boxes, count = self.det.detect_crops(frames, frame_idx, crop_x, crop_y, marks)
        return self._rows_mlp_error(boxes, count, crop_x, crop_y, first_row, n)

    def _rows_mlp_error(self, boxes: torch.Tensor, count: torch.Tensor, crop_x: torch.Tensor, crop_y: torch.Tensor,
                        first_row: int, n: int) -> StepResult:
        """Tracking rows -> ResMLP input gather -> ResMLP -> bbox error on the current stream."""
        s = torch.cuda.current_stream().cuda_stream
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

    # ------------------------------------------------------------------ pipelined host-buffer path
    def _make_slots(self):
        d, B, md = self.device, self.batch, self.det.max_det
        self._slots = []
        for _ in range(2):
            self._slots.append(dict(
                d_views=torch.empty((B, self.view, self.view), dtype=torch.uint8, device=d),
                h_worm=torch.empty((B, 4), dtype=torch.float64).pin_memory(),
                h_boxes=torch.empty((B, md, 6), dtype=torch.float32).pin_memory(),
                h_count=torch.empty((B,), dtype=torch.int32).pin_memory(),
                h_pred=torch.empty((B, 2), dtype=torch.float32).pin_memory(),
                h_valid=torch.empty((B,), dtype=torch.uint8).pin_memory(),
                h_err=torch.empty((B,), dtype=torch.float64).pin_memory(),
                ev_h2d=torch.cuda.Event(), ev_in_free=torch.cuda.Event(), ev_out=torch.cuda.Event(), n=0))
        self._s_copy = torch.cuda.Stream(device=d)
        self._s_post = torch.cuda.Stream(device=d)

    def run_host(self, batches, first_row: int = 0):
        """Pipelined form of ``step_host`` for a stream of batches (the public throughput API).

        ``batches`` yields u8 [n, view, view] HOST arrays (pinned torch tensors are copied straight from
        where they are; anything else goes through a pinned staging buffer first).  Three CUDA streams
        overlap the work of consecutive batches: host->device copy of batch i+1 || crop + YOLOv8s + decode/NMS
        of batch i || tracking rows + ResMLP + bbox error + device->host copy of batch i-1.  Every batch still
        pays its own H2D and D2H inside the call.  Yields one result dict per batch, in order, with the same
        keys as ``step_host``; the arrays are views of pinned buffers that stay valid until the generator is
        advanced twice more.  Results of batch i land in rows [first_row + i * batch, ...) of the tracking
        table (modulo the table size, batch-aligned)."""
        if not hasattr(self, "_slots"):
            self._make_slots()
        main = torch.cuda.current_stream(self.device)
        slots_rows = max(1, self.table_rows // self.batch)
        pending = None           # (slot, n) whose results have been enqueued but not yet handed out
        prev_out = None          # ev_out of the previous batch: guards the detector's single output buffers
        i = 0
        for views in batches:
            sl = self._slots[i & 1]
            n = int(views.shape[0])
            assert n <= self.batch
            src = views if torch.is_tensor(views) else torch.from_numpy(views)
            if not src.is_pinned():
                self.h_views[:n].copy_(src)
                src = self.h_views[:n]
            row0 = ((first_row // self.batch + i) % slots_rows) * self.batch
            # ---- copy stream: H2D once the crop kernel of batch i-2 has consumed this slot
            with torch.cuda.stream(self._s_copy):
                if i >= 2:
                    self._s_copy.wait_event(sl["ev_in_free"])
                sl["d_views"][:n].copy_(src, non_blocking=True)
                sl["ev_h2d"].record(self._s_copy)
            # ---- main stream: crop -> YOLOv8s -> decode/NMS
            main.wait_event(sl["ev_h2d"])
            det = self.det
            det.preprocess(sl["d_views"], self._iota[:n], self._zeros[:n], self._zeros[:n], n)
            sl["ev_in_free"].record(main)
            det.forward(n)
            if prev_out is not None:
                main.wait_event(prev_out)     # the previous batch's rows / D2H have read out_boxes, out_count
            det.postprocess(n)
            ev_det = torch.cuda.Event()
            ev_det.record(main)
            # ---- post stream: tracking rows -> ResMLP -> bbox error -> D2H
            with torch.cuda.stream(self._s_post):
                self._s_post.wait_event(ev_det)
                r = self._rows_mlp_error(det.out_boxes[:n], det.out_count[:n], self._zeros[:n], self._zeros[:n], row0, n)
                sl["h_worm"][:n].copy_(r.worm, non_blocking=True)
                sl["h_boxes"][:n].copy_(r.boxes, non_blocking=True)
                sl["h_count"][:n].copy_(r.count, non_blocking=True)
                sl["h_pred"][:n].copy_(r.pred, non_blocking=True)
                sl["h_valid"][:n].copy_(r.pred_valid, non_blocking=True)
                sl["h_err"][:n].copy_(r.bbox_error, non_blocking=True)
                sl["ev_out"].record(self._s_post)
            sl["n"] = n
            prev_out = sl["ev_out"]
            if pending is not None:
                yield self._collect(pending)
            pending = sl
            i += 1
        if pending is not None:
            yield self._collect(pending)
        main.wait_stream(self._s_post)

    @staticmethod
    def _collect(sl) -> dict[str, np.ndarray]:
        sl["ev_out"].synchronize()
        n = sl["n"]
        return dict(worm=sl["h_worm"][:n].numpy(), boxes=sl["h_boxes"][:n].numpy(), count=sl["h_count"][:n].numpy(),
                    pred=sl["h_pred"][:n].numpy(), pred_valid=sl["h_valid"][:n].numpy(),
                    bbox_error=sl["h_err"][:n].numpy())

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.batch * self.view * self.view

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.batch * (4 * 8 + self.det.max_det * 6 * 4 + 4 + 2 * 4 + 1 + 8)
