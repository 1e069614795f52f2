"""K experiments stepped in LOCK-STEP through the simulator's timing loop (BASELINE configs[4]).

The reference runs one ``Simulator`` per experiment and, inside it, one batch-1 YOLO call per cycle
(wtracker/sim/simulator.py:140-194, sim_controllers/yolo_controller.py:95-109).  Experiments with the same
``TimingConfig`` reach every phase boundary at the same frame index, so here K of them advance together:

  * the per-experiment state of ``Simulator`` / ``ViewController`` / ``SineMotorController`` (platform position,
    pending motor steps and their rounding residuals, the ring of camera-view origins of the current cycle) is held
    as numpy arrays over K and updated with the same float64 operations in the same order, so every experiment's
    integer trace is the one its own ``Simulator`` would produce;
  * the frames stay resident on the device; a detection is a (frame index, crop x, crop y) descriptor, and the K
    descriptors of one cycle are ONE pass of the detector (one H2D of 3 K int32, one D2H of the packed results);
  * controllers are the batched forms of the reference's: ``BatchedYoloController`` (YoloController, optionally with
    the per-cycle ``_cycle_predict_all`` that ``LoggingController`` asks for) and ``BatchedMLPController``
    (MLPController over an on-device bbox table).

``Simulator + YoloController`` of this package and ``BatchedSimulator(K=1)`` give identical traces
(tests/test_gpu_batched.py).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.neural.engine import ResMLPEngine
from wtracker_b200.sim.config import TimingConfig


class BatchedDetector:
    """Detections for any number of (frame, crop origin) descriptors against device-resident frames, in chunks of
    the engine batch; descriptors go up in one H2D, packed results come back in one D2H."""

    def __init__(self, engine: DetectorEngine, frames: torch.Tensor, capacity: int):
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 3 and frames.is_contiguous()
        self.engine, self.frames, self.capacity = engine, frames, int(capacity)
        self.lib = L.lib()
        dev, md = engine.device, engine.max_det
        self.h_desc = torch.zeros((3, self.capacity), dtype=torch.int32).pin_memory()
        self.d_desc = torch.zeros((3, self.capacity), dtype=torch.int32, device=dev)
        # packed results: boxes f32 [cap][max_det][6] then counts i32 [cap], one buffer of 32-bit words
        nb = self.capacity * md * 6
        self.d_out = torch.zeros((nb + self.capacity,), dtype=torch.int32, device=dev)
        self.h_out = torch.zeros((nb + self.capacity,), dtype=torch.int32).pin_memory()
        self.d_boxes = self.d_out[:nb].view(torch.float32).view(self.capacity, md, 6)
        self.d_count = self.d_out[nb:]
        self.h_boxes = self.h_out[:nb].view(torch.float32).view(self.capacity, md, 6).numpy()
        self.h_count = self.h_out[nb:].numpy()
        self._desc_up = torch.cuda.Event()      # the H2D that last read h_desc
        self._desc_pending = False
        self.launched = 0

    def enqueue(self, frame_idx: np.ndarray, crop_x: np.ndarray, crop_y: np.ndarray) -> int:
        """Queues the detections on the current stream (no synchronisation); results land in d_boxes / d_count."""
        n = int(frame_idx.shape[0])
        assert n <= self.capacity
        if self._desc_pending:                  # an enqueue without a fetch (cycle logging) may still be reading h_desc
            self._desc_up.synchronize()
        hd = self.h_desc.numpy()
        hd[0, :n], hd[1, :n], hd[2, :n] = frame_idx, crop_x, crop_y
        eng = self.engine
        with torch.cuda.device(eng.device):
            self.d_desc.copy_(self.h_desc, non_blocking=True)
            self._desc_up.record()
            self._desc_pending = True
            for s in range(0, n, eng.batch):
                m = min(eng.batch, n - s)
                eng.out_boxes, eng.out_count = self.d_boxes[s: s + m], self.d_count[s: s + m]
                eng.detect_crops(self.frames, self.d_desc[0, s: s + m], self.d_desc[1, s: s + m], self.d_desc[2, s: s + m])
        self.launched += n
        return n

    def fetch(self, n: int) -> tuple[np.ndarray, np.ndarray]:
        """One D2H of the packed results + stream synchronize -> (boxes [n, max_det, 6], count [n]) host views."""
        with torch.cuda.device(self.engine.device):
            self.h_out.copy_(self.d_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self.h_boxes[:n], self.h_count[:n]

    def detect(self, frame_idx, crop_x, crop_y):
        return self.fetch(self.enqueue(frame_idx, crop_x, crop_y))


def best_xywh(boxes: np.ndarray, count: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """First box of every image as float32 (x, y, w, h) like ``YoloController.predict`` builds it
    (yolo_controller.py:85-90: np.array([x1, y1, x2 - x1, y2 - y1], dtype=float32)) and the found mask."""
    b = boxes[:, 0, :4].astype(np.float32, copy=False)
    xywh = np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1).astype(np.float32)
    return xywh, count > 0


class BatchedController:
    """Vectorised plug-in surface: the hooks of ``SimController`` the batched loop needs, over K experiments."""

    def on_sim_start(self, sim: "BatchedSimulator"):
        pass

    def on_cycle_end(self, sim: "BatchedSimulator"):
        pass

    def on_camera_frame(self, sim: "BatchedSimulator"):
        pass

    def provide_movement_vectors(self, sim: "BatchedSimulator") -> tuple[np.ndarray, np.ndarray]:
        raise NotImplementedError

    def on_sim_end(self, sim: "BatchedSimulator"):
        pass


class BatchedYoloController(BatchedController):
    """``YoloController`` for K lock-stepped experiments (yolo_controller.py:48-109).

    ``log_cycles=True`` adds what ``LoggingController(YoloController)`` triggers at every cycle end
    (logging_controller.py:187-200 -> ``_cycle_predict_all``, yolo_controller.py:108-109): the detector runs over
    all N buffered views of the finished cycle of every experiment and the absolute worm boxes go to the on-device
    table ``worm_table`` f64 [num_frames][K][4] (NaN = no detection; the last cycle is never logged, as in the
    reference).  ``mic_table`` receives the microscope boxes of the same frames."""

    def __init__(self, timing: TimingConfig, engine: DetectorEngine, frames: torch.Tensor, video_base: np.ndarray,
                 num_frames: int, log_cycles: bool = False, csv_zero_rows: bool = False):
        self.timing, self.engine = timing, engine
        self.csv_zero_rows = csv_zero_rows   # frames without a detection are logged as 0, 0, 0, 0 (what bboxes.csv holds)
        self.K = int(video_base.shape[0])
        self.N = timing.cycle_frame_num
        self.video_base = video_base.astype(np.int64)
        self.num_frames = int(num_frames)
        self.log_cycles = log_cycles
        self.det = BatchedDetector(engine, frames, self.K * (self.N if log_cycles else 1))
        self.cam_w, self.cam_h = timing.camera_size_px
        self.mic_w, self.mic_h = timing.micro_size_px
        assert (self.cam_h, self.cam_w) == (engine.lb.src_h, engine.lb.src_w), "engine view size != camera size"
        self.origins = np.zeros((self.N, self.K, 2), dtype=np.int64)   # camera-view origin per cycle step
        self.filled = 0                                                # views buffered in the current cycle
        self.worm_table = self.mic_table = None
        if log_cycles:
            dev = engine.device
            self.worm_table = torch.full((self.num_frames, self.K, 4), float("nan"), dtype=torch.float64, device=dev)
            self.mic_table = torch.zeros((self.num_frames, self.K, 4), dtype=torch.float64, device=dev)
        self.lib = L.lib()

    def on_sim_start(self, sim):
        self.filled = 0

    def on_camera_frame(self, sim):
        self.origins[sim.cycle_step] = sim.positions - np.array([self.cam_w // 2, self.cam_h // 2])
        self.filled = sim.cycle_step + 1

    def on_cycle_end(self, sim):
        """Called at the first frame of cycle c >= 1, before its camera frame: the buffer holds cycle c - 1."""
        if self.log_cycles:
            n_buf = self.filled
            first = (sim.cycle_number - 1) * self.N
            f = first + np.arange(n_buf, dtype=np.int64)
            fidx = (self.video_base[None, :] + f[:, None]).reshape(-1)                  # frame-major [n_buf][K]
            ox = self.origins[:n_buf, :, 0].reshape(-1)
            oy = self.origins[:n_buf, :, 1].reshape(-1)
            n = self.det.enqueue(fidx, ox, oy)
            rows = self.worm_table[first: first + n_buf].view(-1, 4)
            mic = self.mic_table[first: first + n_buf].view(-1, 4)
            with torch.cuda.device(self.engine.device):
                L.check(self.lib.wt_track_rows(self.det.d_boxes.data_ptr(), self.det.d_count.data_ptr(), self.engine.max_det,
                                               self.det.d_desc[1].data_ptr(), self.det.d_desc[2].data_ptr(), self.cam_w,
                                               self.cam_h, self.mic_w, self.mic_h, rows.data_ptr(), mic.data_ptr(), n,
                                               1 if self.csv_zero_rows else 0,
                                               torch.cuda.current_stream().cuda_stream), "wt_track_rows")
        self.filled = 0

    def provide_movement_vectors(self, sim):
        """deque[-pred_frame_num] of every experiment -> one detector pass -> round(box centre - view centre)."""
        t = self.timing
        step = sim.cycle_step - t.pred_frame_num + 1
        assert 0 <= step < self.filled
        frame = sim.frame_number - t.pred_frame_num + 1
        boxes, count = self.det.detect(self.video_base + frame, self.origins[step, :, 0], self.origins[step, :, 1])
        xywh, found = best_xywh(boxes, count)
        # float32 arithmetic exactly as the reference's scalars: bbox[0] + bbox[2] / 2, then minus camera_size / 2
        mid_x = xywh[:, 0] + xywh[:, 2] / np.float32(2)
        mid_y = xywh[:, 1] + xywh[:, 3] / np.float32(2)
        dx = np.rint(mid_x - np.float32(self.cam_w / 2)).astype(np.int64)
        dy = np.rint(mid_y - np.float32(self.cam_h / 2)).astype(np.int64)
        return np.where(found, dx, 0), np.where(found, dy, 0)


class BatchedMLPController(BatchedController):
    """``MLPController`` for K lock-stepped experiments (mlp_controllers.py:25-68) over an on-device bbox table
    ``worm_table`` f64 [num_frames][K][4] (absolute worm boxes; NaN = missing, e.g. the table a
    ``BatchedYoloController(log_cycles=True)`` pass produced): gather of the 7 input boxes + relativise
    (``wt_mlp_gather``) and the ResMLP (``wt_resmlp_forward``) run once per cycle over all K experiments."""

    def __init__(self, timing: TimingConfig, worm_table: torch.Tensor, predictor, max_speed: float = 0.9,
                 device: str | None = None):
        assert worm_table.is_cuda and worm_table.dtype == torch.float64 and worm_table.dim() == 3
        self.timing = timing
        self.table = worm_table.contiguous()
        self.num_frames, self.K = int(worm_table.shape[0]), int(worm_table.shape[1])
        dev = worm_table.device
        self.mlp = ResMLPEngine(predictor, str(dev) if device is None else device)
        io = predictor.io_config
        self.input_frames = np.asarray(io.input_frames, dtype=np.int64)
        self.k = int(self.input_frames.shape[0])
        px_per_frame = max_speed * (timing.px_per_mm / timing.frames_per_sec)
        self.max_dist = px_per_frame * io.pred_frames[0]
        self.cam_w, self.cam_h = timing.camera_size_px
        K = self.K
        # table row of (frame f, experiment e) = f * K + e: offsets scale by K, `frame` carries e
        self.d_offsets = torch.tensor(self.input_frames * K, dtype=torch.int32, device=dev)
        self.h_rows = torch.zeros(K, dtype=torch.int32).pin_memory()
        self.d_rows = torch.zeros(K, dtype=torch.int32, device=dev)
        self.d_x = torch.zeros((K, 4 * self.k), dtype=torch.float32, device=dev)
        # packed per-cycle result: first box f64 [K][4] | y f32 [K][2] | valid u8 [K]
        self.d_first = torch.zeros((K, 4), dtype=torch.float64, device=dev)
        self.d_y = torch.zeros((K, 2), dtype=torch.float32, device=dev)
        self.d_valid = torch.zeros((K,), dtype=torch.uint8, device=dev)
        self.h_first = torch.zeros((K, 4), dtype=torch.float64).pin_memory()
        self.h_y = torch.zeros((K, 2), dtype=torch.float32).pin_memory()
        self.h_valid = torch.zeros((K,), dtype=torch.uint8).pin_memory()
        self.lib = L.lib()
        self.launched = 0

    def provide_movement_vectors(self, sim):
        t = self.timing
        K = self.K
        base = sim.frame_number - t.pred_frame_num
        f0 = base + int(self.input_frames[0])
        # camera centre as BoxUtils.center(camera_position): origin + size / 2 in float64
        cam_cx = (sim.positions[:, 0] - self.cam_w // 2) + self.cam_w / 2
        cam_cy = (sim.positions[:, 1] - self.cam_h // 2) + self.cam_h / 2
        dev = self.table.device
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream().cuda_stream
            self.h_rows.numpy()[:] = base * K + np.arange(K)
            self.d_rows.copy_(self.h_rows, non_blocking=True)
            L.check(self.lib.wt_mlp_gather(self.table.data_ptr(), self.num_frames * K, self.d_rows.data_ptr(),
                                           self.d_offsets.data_ptr(), self.k, self.d_x.data_ptr(), self.d_valid.data_ptr(),
                                           K, s), "wt_mlp_gather")
            self.mlp.forward(self.d_x, self.d_y)
            if 0 <= f0 < self.num_frames:
                self.h_first.copy_(self.table[f0], non_blocking=True)
            self.h_y.copy_(self.d_y, non_blocking=True)
            self.h_valid.copy_(self.d_valid, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        self.launched += K
        valid = self.h_valid.numpy().astype(bool)
        first = self.h_first.numpy()
        # np.clip on the float32 network output, bounds cast to float32, as in the reference (:62); .item() -> float64
        pred = np.clip(self.h_y.numpy(), -self.max_dist, self.max_dist).astype(np.float64)
        with np.errstate(invalid="ignore"):
            rel_x, rel_y = first[:, 0] - cam_cx, first[:, 1] - cam_cy
            dx = np.where(valid, np.rint(pred[:, 0] + rel_x), 0.0)
            dy = np.where(valid, np.rint(pred[:, 1] + rel_y), 0.0)
        return dx.astype(np.int64), dy.astype(np.int64)


class BatchedSimulator:
    """``Simulator.run`` (simulator.py:140-194) over K experiments that share one ``TimingConfig``.

    positions: int64 [K, 2] platform centres (x, y); ``frame_hw``: frame size for the position clamp
    (view_controller.py:119-131).  ``run()`` returns ``pos_trace`` int64 [num_frames, K, 2] (the position at every
    ``on_camera_frame``) and ``vec_trace`` int64 [cycles, K, 2] (the movement vectors asked of the controller)."""

    def __init__(self, timing: TimingConfig, num_frames: int, init_positions: np.ndarray, frame_hw: tuple[int, int],
                 controller: BatchedController):
        self.timing = timing
        self.num_frames = int(num_frames)
        self.init_positions = np.asarray(init_positions, dtype=np.int64).reshape(-1, 2)
        self.K = self.init_positions.shape[0]
        self.frame_hw = tuple(frame_hw)
        self.controller = controller
        self.positions = self.init_positions.copy()
        self.frame_number = -1
        n_mov = timing.moving_frame_num
        # evaluated per step with scalar numpy calls, like the reference (motor_controllers.py:74-77): identical roundings
        self._frac = np.array([(np.cos((i * np.pi) / n_mov) - np.cos(((i + 1) * np.pi) / n_mov)) / 2 for i in range(n_mov)])
        self._queue = np.zeros((timing.moving_frame_num, self.K, 2), dtype=np.float64)

    @property
    def cycle_number(self) -> int:
        return self.frame_number // self.timing.cycle_frame_num

    @property
    def cycle_step(self) -> int:
        return self.frame_number % self.timing.cycle_frame_num

    def _set_positions(self, pos: np.ndarray):
        h, w = self.frame_hw
        self.positions = np.stack([np.clip(pos[:, 0], 0, w - 1), np.clip(pos[:, 1], 0, h - 1)], axis=1)

    def run(self) -> dict[str, np.ndarray]:
        t = self.timing
        N, n_img, n_mov = t.cycle_frame_num, t.imaging_frame_num, t.moving_frame_num
        ctrl = self.controller
        pos_trace = np.zeros((self.num_frames, self.K, 2), dtype=np.int64)
        vecs = []
        self._set_positions(self.init_positions.copy())
        self.frame_number = -1
        ctrl.on_sim_start(self)
        for idx in range(self.num_frames):
            self.frame_number = idx
            step = idx % N
            if step == 0 and idx >= N:
                ctrl.on_cycle_end(self)
            ctrl.on_camera_frame(self)
            pos_trace[idx] = self.positions
            if step == n_img:
                dx, dy = ctrl.provide_movement_vectors(self)
                vecs.append(np.stack([dx, dy], axis=1))
                # SineMotorController.register_move: frac_i * (dx, dy) in float64 (motor_controllers.py:70-79)
                d = np.stack([dx, dy], axis=1).astype(np.float64)
                self._queue[:] = self._frac[:, None, None] * d[None, :, :]
            if n_img <= step < n_img + n_mov:
                # SineMotorController.step: round-half-even, the residual is carried into the next step (:81-88)
                i = step - n_img
                f = self._queue[i]
                r = np.rint(f)
                if i + 1 < n_mov:
                    self._queue[i + 1] += f - r
                self._set_positions(self.positions + r.astype(np.int64))
        ctrl.on_sim_end(self)
        vec_trace = np.stack(vecs) if vecs else np.zeros((0, self.K, 2), dtype=np.int64)
        return dict(pos_trace=pos_trace, vec_trace=vec_trace)
