"""Runtime wrapper of the native detector: owns the torch tensors (weights, workspace, tables,
outputs), hands raw pointers to the C ABI and exposes crop -> detect as one call.

PyTorch is plumbing here (memory + streams); all arithmetic happens in libwtracker_b200.so.
"""

from __future__ import annotations


import ctypes as C

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.arch import YoloV8Arch
from wtracker_b200.detector.letterbox import Letterbox, letterbox_for, resize_tables
from wtracker_b200.detector.program import Program, blob_tensor, build_program, ops_as_ctypes
from wtracker_b200.detector.weights import infer_arch


def _ptr(t: torch.Tensor | None) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ingest_window(x0: int, y0: int, w: int, h: int, fw: int, fh: int) -> tuple[int, int, int, int]:
    """Frame ingest of one view (``DetectorEngine.detect_frames``): the (w, h) window of an (fw, fh) frame that travels to the
    GPU for the camera view with origin (x0, y0) — the view rectangle shifted to lie inside the frame — and the crop origin
    (cx, cy) inside that window.  The crop kernel clamps its addresses to the window, which replicates the window's edge
    where the view hangs over the frame border: exactly the frame's edge, because the window touches that border."""
    assert w <= fw and h <= fh
    wx, wy = min(max(x0, 0), fw - w), min(max(y0, 0), fh - h)
    return wx, wy, x0 - wx, y0 - wy


class DetectorEngine:
    """YOLOv8 detector for views of a fixed size ``view_hw`` letterboxed to ``imgsz``.

    detect_crops(): frames resident on the device + per-image (frame index, crop origin) ->
    rows [x1, y1, x2, y2, conf, anchor] per image (view pixel coordinates) and a count per image.
    """

    def __init__(self, state_dict: dict[str, torch.Tensor], view_hw: tuple[int, int], imgsz: int = 384,
                 batch: int = 16, conf: float = 0.1, iou: float = 0.7, max_det: int = 1, device: str = "cuda:0",
                 conv_impl: int = 0, arch: YoloV8Arch | None = None, fuse: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        self.device = torch.device(device)
        self.arch = arch or infer_arch(state_dict)
        self.lb: Letterbox = letterbox_for(view_hw, imgsz)
        self.batch = int(batch)
        self.conf, self.iou, self.max_det = float(conf), float(iou), int(max_det)
        self.program: Program = build_program(state_dict, self.arch, self.lb.dst_h, self.lb.dst_w,
                                              chain=conv_impl == 0 and fuse)   # fuse=False: one launch per conv (tests)

        with torch.cuda.device(self.device):
            self.weights = blob_tensor(self.program).to(self.device)
            self._bufs, self._ops = ops_as_ctypes(self.program)
            ws_bytes = self.lib.wt_engine_workspace_bytes(self._bufs, len(self.program.bufs), self.batch)
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            L.check(self.lib.wt_engine_create(self._bufs, len(self.program.bufs), self._ops, len(self.program.ops),
                                              self.batch, self.weights.data_ptr(), self.weights.numel(),
                                              self.workspace.data_ptr(), ws_bytes, conv_impl, C.byref(handle)),
                    "wt_engine_create")
            self.handle = handle
            # letterbox tables
            self._tables = {}
            if self.lb.resample:
                self._tables = {k: torch.from_numpy(v).to(self.device) for k, v in resize_tables(self.lb).items()}
            self._lb_c = L.WtLetterbox(self.lb.src_w, self.lb.src_h, self.lb.dst_w, self.lb.dst_h, self.lb.new_w,
                                       self.lb.new_h, self.lb.pad_left, self.lb.pad_top,
                                       _ptr(self._tables.get("xofs")), _ptr(self._tables.get("xcoef")),
                                       _ptr(self._tables.get("yofs")), _ptr(self._tables.get("ycoef")))
            # post-process state
            A = self.program.total_anchors
            # boxes f32 [batch][max_det][6] and counts i32 [batch] share one buffer of 32-bit words: one D2H per call
            nb = self.batch * self.max_det * 6
            self._out_words = torch.zeros((nb + self.batch,), dtype=torch.int32, device=self.device)
            self.out_boxes = self._out_words[:nb].view(torch.float32).view(self.batch, self.max_det, 6)
            self.out_count = self._out_words[nb:]
            self.scratch = torch.zeros(self.lib.wt_post_scratch_bytes(self.batch, A), dtype=torch.uint8,
                                       device=self.device)
            levels = (L.WtHeadLevel * 3)()
            for i, h in enumerate(self.program.head):
                levels[i] = L.WtHeadLevel(None, None, self.buffer_ptr(h["cls_logit"]), h["h"], h["w"], h["stride"],
                                          L.WT_DT_F32, 0, None, 0.0, self.buffer_ptr(h["box_feat"]),
                                          self.weights.data_ptr() + h["box_w_off"],
                                          self.weights.data_ptr() + h["box_b_off"], h["box_c"])
            self._levels = levels
            pad_x, pad_y = self.lb.scale_pad
            self._post = L.WtPostParams(self.conf, self.iou, self.max_det, self.lb.dst_w, self.lb.dst_h,
                                        self.lb.src_w, self.lb.src_h, self.lb.gain, pad_x, pad_y)

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self.lib.wt_engine_destroy(h)
            self.handle = None

    def buffer_ptr(self, buf_id: int) -> int:
        return self.lib.wt_engine_buffer(self.handle, buf_id)

    def buffer_tensor(self, name_or_id, n: int | None = None) -> torch.Tensor:
        """Copy of an activation buffer as a torch tensor [n, h, w, c] (debug / tests)."""
        bid = name_or_id if isinstance(name_or_id, int) else self.program.buf_id(name_or_id)
        h, w, c, dt = self.program.bufs[bid]
        dtype = {L.WT_DT_BF16: torch.bfloat16, L.WT_DT_F32: torch.float32, L.WT_DT_U8: torch.uint8}[dt]
        n = self.batch if n is None else n
        nbytes = n * h * w * c * torch.empty((), dtype=dtype).element_size()
        off = self.buffer_ptr(bid) - self.workspace.data_ptr()
        return self.workspace[off: off + nbytes].view(dtype).view(n, h, w, c).clone()

    @property
    def input_view(self) -> torch.Tensor:
        """The u8 network input buffer [batch, net_h, net_w] (a view into the workspace)."""
        h, w, _, _ = self.program.bufs[0]
        off = self.buffer_ptr(0) - self.workspace.data_ptr()
        return self.workspace[off: off + self.batch * h * w].view(self.batch, h, w)

    # ------------------------------------------------------------------ stages
    def preprocess(self, frames: torch.Tensor, frame_idx: torch.Tensor, crop_x: torch.Tensor, crop_y: torch.Tensor,
                   n: int, out_f32: torch.Tensor | None = None) -> None:
        """K1-K4 into the engine's input buffer (and optionally the fp32 NCHW reference tensor)."""
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous() and frames.dim() == 3
        assert frame_idx.dtype == crop_x.dtype == crop_y.dtype == torch.int32
        L.check(self.lib.wt_preprocess(frames.data_ptr(), frames.shape[0], frames.shape[1], frames.shape[2],
                                       frame_idx.data_ptr(), crop_x.data_ptr(), crop_y.data_ptr(), n,
                                       C.byref(self._lb_c), self.buffer_ptr(0), _ptr(out_f32), _stream()),
                "wt_preprocess")

    def forward(self, n: int, first_op: int = 0, last_op: int | None = None) -> None:
        """K5: runs the conv program on the first n images of the input buffer."""
        last = len(self.program.ops) if last_op is None else last_op
        L.check(self.lib.wt_engine_forward(self.handle, n, first_op, last, _stream()), "wt_engine_forward")

    def postprocess(self, n: int) -> None:
        """K6-K8 into out_boxes / out_count."""
        L.check(self.lib.wt_decode_nms(self._levels, 3, n, C.byref(self._post), self.out_boxes.data_ptr(),
                                       self.out_count.data_ptr(), self.scratch.data_ptr(), _stream()),
                "wt_decode_nms")

    # ------------------------------------------------------------------ whole path
    def detect_crops(self, frames: torch.Tensor, frame_idx: torch.Tensor, crop_x: torch.Tensor, crop_y: torch.Tensor,
                     marks: list | None = None):
        """Device-resident path: returns (boxes [n, max_det, 6], count [n]) device tensors (views into
        engine-owned outputs; valid until the next call).  ``marks`` (optional) receives a CUDA event
        recorded after each stage (pre, forward, post) for stage timing."""
        n = int(frame_idx.numel())
        assert n <= self.batch

        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)

        self.preprocess(frames, frame_idx, crop_x, crop_y, n)
        mark()
        self.forward(n)
        mark()
        self.postprocess(n)
        mark()
        return self.out_boxes[:n], self.out_count[:n]

    # CUDA graph of the three stages for small host batches (the per-cycle calls of the simulator: 1, 9 or 15 views).
    # Such a call is launch-bound — ~59 launches of a few microseconds each — so the stages are captured once per
    # (batch, source buffer) on the engine's static buffers and replayed; the programmatic-dependent-launch edges and the
    # fork / join of the engine's side stream are part of the capture.  ``graph_max_batch = 0`` switches it off.
    graph_max_batch = 16

    def _detect_crops_graphed(self, frames: torch.Tensor, frame_idx: torch.Tensor, crop_x: torch.Tensor, crop_y: torch.Tensor):
        n = int(frame_idx.numel())
        if n > self.graph_max_batch:
            return self.detect_crops(frames, frame_idx, crop_x, crop_y)
        key = (n, frames.data_ptr(), frame_idx.data_ptr(), crop_x.data_ptr(), crop_y.data_ptr(), self.out_boxes.data_ptr())
        graphs = self.__dict__.setdefault("_graphs", {})
        g = graphs.get(key)
        if g is None:                                # first call with these buffers: run eagerly (warms every kernel up) ...
            graphs[key] = False
            return self.detect_crops(frames, frame_idx, crop_x, crop_y)
        if g is False:                               # ... second call: capture
            torch.cuda.current_stream().synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.detect_crops(frames, frame_idx, crop_x, crop_y)
                graphs[key] = g
            except Exception as e:                   # same kernels, launched one by one
                import warnings

                warnings.warn(f"DetectorEngine: CUDA graph capture failed ({e}); small batches are launched eagerly")
                self.graph_max_batch = 0
                torch.cuda.synchronize()
                return self.detect_crops(frames, frame_idx, crop_x, crop_y)
        g.replay()
        return self.out_boxes[:n], self.out_count[:n]

    def _host_staging(self):
        """Persistent staging of the host path (allocated once): pinned view buffer, device view buffer, index
        vectors and a pinned mirror of the packed result words."""
        if not hasattr(self, "_h_views"):
            h, w = self.lb.src_h, self.lb.src_w
            self._h_views = torch.empty((self.batch, h, w), dtype=torch.uint8).pin_memory()
            self._h_views_np = self._h_views.numpy()
            self._d_views = torch.empty((self.batch, h, w), dtype=torch.uint8, device=self.device)
            self._iota = torch.arange(self.batch, dtype=torch.int32, device=self.device)
            self._zeros = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
            self._h_out = torch.empty_like(self._out_words, device="cpu").pin_memory()
            self._h_boxes = self._h_out[: self.batch * self.max_det * 6].view(torch.float32).view(self.batch, self.max_det, 6).numpy()
            self._h_count = self._h_out[self.batch * self.max_det * 6:].numpy()

    def detect_frames(self, frames: list[np.ndarray], crop_x: list[int], crop_y: list[int]) -> tuple[np.ndarray, np.ndarray]:
        """Host path with the crop on the device (frame ingest: reference utils/frame_reader.py:137-144 +
        view_controller.py:45-61): ``frames`` are whole (H, W) u8 grey frames, ``crop_x / crop_y`` the camera-view
        origins in frame coordinates (may hang over the border: replicate).  Same return as ``detect_views``.

        A frame is used for ONE view here, so only the view-sized WINDOW of the frame that contains the view's pixels
        travels: the window is the view rectangle shifted to lie inside the frame (a strided copy straight from the frame
        into the persistent pinned buffer, one H2D per chunk), and the crop kernel cuts the view out of it at the
        origin (x0 - window x, y0 - window y) — non-zero exactly where the view hangs over the border, where the kernel's
        clamp-addressing replicates the window's edge = the frame's edge.  16x fewer bytes than the whole frame at the
        reference's 360-pixel view.  Views larger than the frame take the whole-frame route (`_detect_whole_frames`)."""
        n = len(frames)
        fh, fw = frames[0].shape
        h, w = self.lb.src_h, self.lb.src_w
        if h > fh or w > fw:
            return self._detect_whole_frames(frames, crop_x, crop_y)
        boxes = np.zeros((n, self.max_det, 6), np.float32)
        counts = np.zeros((n,), np.int32)
        with torch.cuda.device(self.device):
            self._host_staging()
            if not hasattr(self, "_h_desc"):
                self._h_desc = torch.zeros((2, self.batch), dtype=torch.int32).pin_memory()
                self._d_desc = torch.zeros((2, self.batch), dtype=torch.int32, device=self.device)
            stream = torch.cuda.current_stream()
            own = self.out_boxes.data_ptr() == self._out_words.data_ptr()
            hd = self._h_desc.numpy()
            for s in range(0, n, self.batch):
                m = min(self.batch, n - s)
                for i in range(m):
                    f = frames[s + i]
                    assert f.shape == (fh, fw) and f.dtype == np.uint8, "frames must be grey u8 of one size"
                    x0, y0 = int(crop_x[s + i]), int(crop_y[s + i])
                    wx, wy, hd[0, i], hd[1, i] = ingest_window(x0, y0, w, h, fw, fh)
                    np.copyto(self._h_views_np[i], f[wy: wy + h, wx: wx + w])
                self._d_views[:m].copy_(self._h_views[:m], non_blocking=True)
                self._d_desc.copy_(self._h_desc, non_blocking=True)
                b, c = self._detect_crops_graphed(self._d_views, self._iota[:m], self._d_desc[0, :m], self._d_desc[1, :m])
                if own:
                    self._h_out.copy_(self._out_words, non_blocking=True)
                    stream.synchronize()
                    boxes[s: s + m] = self._h_boxes[:m]
                    counts[s: s + m] = self._h_count[:m]
                else:
                    boxes[s: s + m] = b.cpu().numpy()
                    counts[s: s + m] = c.cpu().numpy()
        return boxes, counts

    def _detect_whole_frames(self, frames: list[np.ndarray], crop_x: list[int], crop_y: list[int]) -> tuple[np.ndarray, np.ndarray]:
        """``detect_frames`` with the whole frames on the device (views larger than the frame; also what
        ``HotPath.run_frames`` does for streams of frames that are cropped more than once)."""
        n = len(frames)
        fh, fw = frames[0].shape
        boxes = np.zeros((n, self.max_det, 6), np.float32)
        counts = np.zeros((n,), np.int32)
        with torch.cuda.device(self.device):
            self._host_staging()
            if getattr(self, "_frame_shape", None) != (fh, fw):
                self._frame_shape = (fh, fw)
                self._h_frames = torch.empty((self.batch, fh, fw), dtype=torch.uint8).pin_memory()
                self._h_frames_np = self._h_frames.numpy()
                self._d_frames = torch.empty((self.batch, fh, fw), dtype=torch.uint8, device=self.device)
                self._h_fdesc = torch.zeros((2, self.batch), dtype=torch.int32).pin_memory()
                self._d_fdesc = torch.zeros((2, self.batch), dtype=torch.int32, device=self.device)
            stream = torch.cuda.current_stream()
            own = self.out_boxes.data_ptr() == self._out_words.data_ptr()
            for s in range(0, n, self.batch):
                m = min(self.batch, n - s)
                hd = self._h_fdesc.numpy()
                for i in range(m):
                    f = frames[s + i]
                    assert f.shape == (fh, fw) and f.dtype == np.uint8, "frames must be grey u8 of one size"
                    np.copyto(self._h_frames_np[i], f)
                    hd[0, i], hd[1, i] = crop_x[s + i], crop_y[s + i]
                self._d_frames[:m].copy_(self._h_frames[:m], non_blocking=True)
                self._d_fdesc.copy_(self._h_fdesc, non_blocking=True)
                b, c = self._detect_crops_graphed(self._d_frames, self._iota[:m], self._d_fdesc[0, :m], self._d_fdesc[1, :m])
                if own:
                    self._h_out.copy_(self._out_words, non_blocking=True)
                    stream.synchronize()
                    boxes[s: s + m] = self._h_boxes[:m]
                    counts[s: s + m] = self._h_count[:m]
                else:
                    boxes[s: s + m] = b.cpu().numpy()
                    counts[s: s + m] = c.cpu().numpy()
        return boxes, counts

    def detect_views(self, views: list[np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
        """Host path (what ``YoloController.predict`` calls): list of (h, w) u8 grey camera views -> (boxes
        [n, max_det, 6], count [n]) numpy.  Per chunk of ``batch`` views: the views are written straight into a
        persistent pinned buffer (they may be non-contiguous slices of a frame), ONE H2D, the three stages, ONE D2H
        of the packed result words, one stream synchronize.  Nothing is allocated per call."""
        n = len(views)
        h, w = self.lb.src_h, self.lb.src_w
        boxes = np.zeros((n, self.max_det, 6), np.float32)
        counts = np.zeros((n,), np.int32)
        with torch.cuda.device(self.device):
            self._host_staging()
            stream = torch.cuda.current_stream()
            own = self.out_boxes.data_ptr() == self._out_words.data_ptr()
            for s in range(0, n, self.batch):
                chunk = views[s: s + self.batch]
                m = len(chunk)
                for i, v in enumerate(chunk):
                    assert v.shape == (h, w), f"views must be {(h, w)}, got {tuple(v.shape)}"
                    np.copyto(self._h_views_np[i], v)
                self._d_views[:m].copy_(self._h_views[:m], non_blocking=True)
                b, c = self._detect_crops_graphed(self._d_views, self._iota[:m], self._zeros[:m], self._zeros[:m])
                if own:
                    self._h_out.copy_(self._out_words, non_blocking=True)
                    stream.synchronize()
                    boxes[s: s + m] = self._h_boxes[:m]
                    counts[s: s + m] = self._h_count[:m]
                else:       # outputs rebound by an owner (HotPath): two small copies
                    boxes[s: s + m] = b.cpu().numpy()
                    counts[s: s + m] = c.cpu().numpy()
        return boxes, counts
