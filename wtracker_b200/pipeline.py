"""The whole per-frame hot path as one object: crop -> YOLOv8s -> decode/NMS -> tracking rows ->
ResMLP position prediction -> bbox error, every stage a kernel of libwtracker_b200.so.

``step_device`` is the device-resident path (frames already in HBM; what ``bench.py`` reports as
``value``); ``step_host`` / ``run_host`` are the same work entered with HOST buffers (camera views in,
result rows out — what ``bench.py`` reports as ``e2e``); ``run_frames`` is the ingest form: whole
camera FRAMES come from the host and the crops are taken on the device.
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


class ResultPack:
    """Every per-frame result of one batch in ONE byte buffer, so that a batch costs one device->host copy:
    worm f64 [B,4] | bbox_error f64 [B] | boxes f32 [B,max_det,6] | pred f32 [B,2] | count i32 [B] | valid u8 [B].
    ``dev`` lives on the GPU (the kernels write straight into its sections), ``host`` is its pinned mirror."""

    def __init__(self, batch: int, max_det: int, device: torch.device):
        sections = [("worm", torch.float64, (batch, 4)), ("err", torch.float64, (batch,)),
                    ("boxes", torch.float32, (batch, max_det, 6)), ("pred", torch.float32, (batch, 2)),
                    ("count", torch.int32, (batch,)), ("valid", torch.uint8, (batch,))]
        offs, total = [], 0
        for _, dt, shape in sections:
            total = (total + 15) & ~15
            offs.append(total)
            total += int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        self.nbytes = (total + 15) & ~15
        self.dev = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.host = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        for (name, dt, shape), off in zip(sections, offs):
            nb = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            setattr(self, name, self.dev[off: off + nb].view(dt).view(shape))
            setattr(self, "h_" + name, self.host[off: off + nb].view(dt).view(shape))

    def host_dict(self, n: int) -> dict[str, np.ndarray]:
        return dict(worm=self.h_worm[:n].numpy(), boxes=self.h_boxes[:n].numpy(), count=self.h_count[:n].numpy(),
                    pred=self.h_pred[:n].numpy(), pred_valid=self.h_valid[:n].numpy(), bbox_error=self.h_err[:n].numpy())


class HotPath:
    def __init__(self, state_dict: dict, predictor: WormPredictor, view: int = 640, imgsz: int = 640, batch: int = 64,
                 micro: int = 51, table_rows: int = 1 << 16, device: str = "cuda:0", conf: float = 0.1,
                 iou: float = 0.7, max_det: int = 1, fused_tail: bool = True):
        self.lib = L.lib()
        self.device = torch.device(device)
        self.batch, self.view, self.micro = batch, view, micro
        self.det = DetectorEngine(state_dict, (view, view), imgsz, batch=batch, conf=conf, iou=iou, max_det=max_det,
                                  device=device)
        self.mlp = ResMLPEngine(predictor, device)
        offs = [int(v) for v in predictor.io_config.input_frames]
        self.offsets = torch.tensor(offs, dtype=torch.int32, device=self.device)
        self.k = len(offs)
        self.fused_tail = fused_tail and self.k <= L.WT_TAIL_MAX_K
        d = self.device
        self.table_rows = table_rows
        self.table = torch.full((table_rows, 4), float("nan"), dtype=torch.float64, device=d)   # worm xywh by frame
        self.mic_table = torch.zeros((table_rows, 4), dtype=torch.float64, device=d)
        self.mlp_x = torch.zeros((batch, 4 * self.k), dtype=torch.float32, device=d)
        self.pack = ResultPack(batch, max_det, d)        # results of step_device / step_host
        self.d_views = torch.empty((batch, view, view), dtype=torch.uint8, device=d)
        self.h_views = torch.empty((batch, view, view), dtype=torch.uint8).pin_memory()
        self._rows = torch.arange(table_rows, dtype=torch.int32, device=d)
        self._iota = self._rows[:batch]
        self._zeros = torch.zeros(batch, dtype=torch.int32, device=d)
        self._tail = L.WtTailArgs()
        self._tail.max_det = max_det
        self._tail.cam_w = self._tail.cam_h = view
        self._tail.mic_w = self._tail.mic_h = micro
        self._tail.table, self._tail.mic_table = self.table.data_ptr(), self.mic_table.data_ptr()
        self._tail.table_rows = table_rows
        self._tail.k = self.k
        for j, o in enumerate(offs[: L.WT_TAIL_MAX_K]):
            self._tail.offsets[j] = o
        self._tail.mlp = self.mlp.desc
        self._tail.weights_t = self.mlp.weights_t.data_ptr()
        self._tail.x = self.mlp_x.data_ptr()
        self._bind(self.pack)

    def _bind(self, pack: ResultPack) -> None:
        """The detector and the tail kernel write their results into ``pack`` from now on."""
        self._pack = pack
        self.det.out_boxes, self.det.out_count = pack.boxes, pack.count

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
        """Tracking rows -> ResMLP input gather -> ResMLP -> bbox error on the current stream: ONE launch
        (``wt_hot_tail``), or the four separate entry points when ``fused_tail`` is off (tests compare the two)."""
        s = torch.cuda.current_stream().cuda_stream
        pk = self._pack
        worm = self.table[first_row: first_row + n]
        mic = self.mic_table[first_row: first_row + n]
        if self.fused_tail:
            t = self._tail
            t.boxes, t.count = boxes.data_ptr(), count.data_ptr()
            t.crop_x, t.crop_y = crop_x.data_ptr(), crop_y.data_ptr()
            t.first_row, t.n = first_row, n
            t.valid, t.y, t.err = pk.valid.data_ptr(), pk.pred.data_ptr(), pk.err.data_ptr()
            L.check(self.lib.wt_hot_tail(C.byref(t), s), "wt_hot_tail")
        else:
            L.check(self.lib.wt_track_rows(boxes.data_ptr(), count.data_ptr(), self.det.max_det, crop_x.data_ptr(),
                                           crop_y.data_ptr(), self.view, self.view, self.micro, self.micro,
                                           worm.data_ptr(), mic.data_ptr(), n, 0, s), "wt_track_rows")
            rows = self._rows[first_row: first_row + n]
            L.check(self.lib.wt_mlp_gather(self.table.data_ptr(), self.table_rows, rows.data_ptr(),
                                           self.offsets.data_ptr(), self.k, self.mlp_x.data_ptr(),
                                           pk.valid.data_ptr(), n, s), "wt_mlp_gather")
            self.mlp.forward(self.mlp_x[:n], pk.pred[:n])
            L.check(self.lib.wt_bbox_error(worm.data_ptr(), mic.data_ptr(), pk.err.data_ptr(), n, s), "wt_bbox_error")
        return StepResult(boxes, count, worm, mic, pk.pred[:n], pk.valid[:n], pk.err[:n])

    def _results_to_host(self, pack: ResultPack, worm: torch.Tensor, n: int) -> None:
        """One device->host copy of the whole result pack (the worm rows join it from the tracking table)."""
        pack.worm[:n].copy_(worm, non_blocking=True)
        pack.host.copy_(pack.dev, non_blocking=True)

    # ------------------------------------------------------------------ host-buffer step (public API)
    def step_host(self, views: np.ndarray | torch.Tensor, first_row: int = 0) -> dict[str, np.ndarray]:
        """views: u8 [n, view, view] HOST array of camera views.  Host->device copy of the views and
        the device->host copy of every result are part of the call (no torch compute kernels involved).
        Returns host arrays: worm [n,4] f64 xywh (view px, NaN = none), boxes [n,max_det,6], count [n],
        pred [n,2], pred_valid [n], bbox_error [n]."""
        n = views.shape[0]
        assert n <= self.batch
        self._bind(self.pack)
        src = views if torch.is_tensor(views) else torch.from_numpy(views)
        if not src.is_pinned():
            self.h_views[:n].copy_(src)      # synchronous host copy; the H2D below is ordered before the next call's
            src = self.h_views[:n]           # refill by the stream synchronize at the end of this call
        self.d_views[:n].copy_(src, non_blocking=True)
        r = self.step_device(self.d_views, self._iota[:n], self._zeros[:n], self._zeros[:n], first_row)
        self._results_to_host(self.pack, r.worm, n)
        torch.cuda.current_stream().synchronize()
        return self.pack.host_dict(n)

    # ------------------------------------------------------------------ pipelined host-buffer paths
    def _make_slots(self, frame_shape: tuple[int, int] | None = None):
        d, B, md = self.device, self.batch, self.det.max_det
        if not hasattr(self, "_slots"):
            self._slots = []
            for _ in range(2):
                self._slots.append(dict(
                    d_views=torch.empty((B, self.view, self.view), dtype=torch.uint8, device=d),
                    h_views=torch.empty((B, self.view, self.view), dtype=torch.uint8).pin_memory(),
                    pack=ResultPack(B, md, d), desc=torch.zeros((3, B), dtype=torch.int32, device=d),
                    h_desc=torch.zeros((3, B), dtype=torch.int32).pin_memory(),
                    ev_h2d=torch.cuda.Event(), ev_in_free=torch.cuda.Event(), ev_out=torch.cuda.Event(), n=0,
                    h2d_pending=False))
            self._s_copy = torch.cuda.Stream(device=d)
            self._s_post = torch.cuda.Stream(device=d)
        if frame_shape is not None and self._slots[0].get("frame_shape") != tuple(frame_shape):
            for sl in self._slots:
                sl["d_frames"] = torch.empty((B, *frame_shape), dtype=torch.uint8, device=d)
                sl["h_frames"] = torch.empty((B, *frame_shape), dtype=torch.uint8).pin_memory()
                sl["frame_shape"] = tuple(frame_shape)

    def _stage(self, sl: dict, src, n: int, key: str) -> torch.Tensor:
        """Pinned source for the H2D of this slot: the caller's tensor when it is pinned already, else the slot's own
        pinned staging buffer — refilled only after the H2D that last read it has completed."""
        src = src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src))
        if src.is_pinned():
            return src
        if sl["h2d_pending"]:
            sl["ev_h2d"].synchronize()
        stage = sl[key][:n]
        stage.copy_(src)
        return stage

    def _pipeline(self, items, first_row: int, from_frames: bool):
        main = torch.cuda.current_stream(self.device)
        slots_rows = max(1, self.table_rows // self.batch)
        pending = None           # slot whose results have been enqueued but not yet handed out
        i = 0
        try:
            for item in items:
                sl = self._slots[i & 1]
                if from_frames:
                    frames, cx, cy = item
                    n = int(frames.shape[0])
                else:
                    frames, cx, cy, n = item, None, None, int(item.shape[0])
                assert n <= self.batch
                src = self._stage(sl, frames, n, "h_frames" if from_frames else "h_views")
                if from_frames:
                    sl["h_desc"][1, :n] = torch.as_tensor(np.asarray(cx, dtype=np.int32))
                    sl["h_desc"][2, :n] = torch.as_tensor(np.asarray(cy, dtype=np.int32))
                row0 = ((first_row // self.batch + i) % slots_rows) * self.batch
                # ---- copy stream: H2D once the crop kernel of batch i-2 has consumed this slot
                with torch.cuda.stream(self._s_copy):
                    if i >= 2:
                        self._s_copy.wait_event(sl["ev_in_free"])
                        self._s_copy.wait_event(sl["ev_out"])     # (the tail kernel of batch i-2 read this slot's descriptors)
                    dst = sl["d_frames"] if from_frames else sl["d_views"]
                    dst[:n].copy_(src, non_blocking=True)
                    if from_frames:
                        sl["desc"].copy_(sl["h_desc"], non_blocking=True)
                    sl["ev_h2d"].record(self._s_copy)
                    sl["h2d_pending"] = True
                # ---- main stream: crop -> YOLOv8s -> decode/NMS (results into this slot's pack)
                main.wait_event(sl["ev_h2d"])
                if i >= 2:
                    main.wait_event(sl["ev_out"])     # the slot's pack has been copied out (batch i-2)
                det = self.det
                self._bind(sl["pack"])
                crop_x = sl["desc"][1, :n] if from_frames else self._zeros[:n]
                crop_y = sl["desc"][2, :n] if from_frames else self._zeros[:n]
                det.preprocess(dst, self._iota[:n], crop_x, crop_y, n)
                sl["ev_in_free"].record(main)
                det.forward(n)
                det.postprocess(n)
                ev_det = torch.cuda.Event()
                ev_det.record(main)
                # ---- post stream: tracking rows -> ResMLP -> bbox error (one launch) -> one D2H
                with torch.cuda.stream(self._s_post):
                    self._s_post.wait_event(ev_det)
                    r = self._rows_mlp_error(det.out_boxes[:n], det.out_count[:n], crop_x, crop_y, row0, n)
                    self._results_to_host(sl["pack"], r.worm, n)
                    sl["ev_out"].record(self._s_post)
                sl["n"] = n
                if pending is not None:
                    yield self._collect(pending)
                pending = sl
                i += 1
            if pending is not None:
                yield self._collect(pending)
        finally:
            main.wait_stream(self._s_post)
            main.wait_stream(self._s_copy)
            self._bind(self.pack)

    def run_host(self, batches, first_row: int = 0):
        """Pipelined form of ``step_host`` for a stream of batches (the public throughput API).

        ``batches`` yields u8 [n, view, view] HOST arrays (pinned torch tensors are copied straight from
        where they are; anything else goes through the slot's own pinned staging buffer).  Three CUDA streams
        overlap the work of consecutive batches: host->device copy of batch i+1 || crop + YOLOv8s + decode/NMS
        of batch i || tracking rows + ResMLP + bbox error + device->host copy of batch i-1.  Every batch still
        pays its own H2D and D2H inside the call.  Yields one result dict per batch, in order, with the same
        keys as ``step_host``; the arrays are views of pinned buffers that stay valid until the generator is
        advanced again.  Results of batch i land in rows [first_row + i * batch, ...) of the tracking
        table (modulo the table size, batch-aligned)."""
        self._make_slots()
        yield from self._pipeline(batches, first_row, from_frames=False)

    def run_frames(self, batches, first_row: int = 0):
        """Frame ingest (reference: FrameReader.__getitem__ utils/frame_reader.py:137-144 + ViewController.read
        view_controller.py:45-61): ``batches`` yields ``(frames u8 [n, H, W] HOST, crop_x [n], crop_y [n])`` —
        whole camera frames and the camera-view origins in frame coordinates.  The frames go host -> pinned
        ring -> device (2 MB per 1080p frame, double-buffered against the detector) and the views are taken by
        the crop kernel on the device with replicate borders, so no padded frame and no host-side crop exists.
        Yields the same dicts as ``run_host``; ``worm`` is in FRAME pixels."""
        it = iter(batches)
        try:
            first = next(it)
        except StopIteration:
            return
        self._make_slots(tuple(first[0].shape[1:]))

        def chain():
            yield first
            yield from it

        yield from self._pipeline(chain(), first_row, from_frames=True)

    @staticmethod
    def _collect(sl) -> dict[str, np.ndarray]:
        sl["ev_out"].synchronize()
        return sl["pack"].host_dict(sl["n"])

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.batch * self.view * self.view

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.pack.nbytes
