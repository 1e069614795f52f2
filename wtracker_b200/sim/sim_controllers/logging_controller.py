"""``LoggingController`` / ``LogConfig`` with the reference's interface
(wtracker/sim/sim_controllers/logging_controller.py:14-224): wraps any controller, forwards every hook, and at
each cycle end writes that cycle's rows of ``bboxes.csv`` (17 columns, wire format of every downstream analysis).

What changed underneath: the per-frame arithmetic of ``_log_cycle`` — absolute worm coordinates, the zeroing of
rows without a prediction, the integer crop of ``BoxUtils.discretize`` — is one CUDA kernel (``wt_log_rows``)
over the whole cycle, and ``log_table_device`` exposes the same kernel for whole tables that already live on the
GPU (the batched hot path logs many frames per launch).  The csv text is byte-identical to the reference's
(tests/golden/reference_bboxes_*.csv), including its quirks: rows without a prediction are logged as
``0.0,0.0,0.0,0.0`` (discretize zeroes them in place first, so the error-view branch never fires), the wrm columns
print in the dtype of the controller's prediction array, and the final cycle of a simulation is never logged.
"""

from __future__ import annotations

import os
import queue
import threading
from collections import deque
from copy import deepcopy
from dataclasses import dataclass, field

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.sim.simulator import SimController, Simulator
from wtracker_b200.utils.config_base import ConfigBase
from wtracker_b200.utils.log_utils import CSVLogger

LOG_COLUMNS = ["frame", "cycle", "phase", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w",
               "mic_h", "wrm_x", "wrm_y", "wrm_w", "wrm_h"]


def _join(*parts: str) -> str:
    return os.path.join(*parts).replace("\\", "/")


@dataclass
class LogConfig(ConfigBase):
    root_folder: str
    save_mic_view: bool = False
    save_cam_view: bool = False
    save_err_view: bool = True
    save_wrm_view: bool = False
    mic_folder_name: str = "micro"
    cam_folder_name: str = "camera"
    err_folder_name: str = "errors"
    wrm_folder_name: str = "worms"
    bbox_file_name: str = "bboxes.csv"
    mic_file_name: str = "mic_{:09d}.png"
    cam_file_name: str = "cam_{:09d}.png"
    wrm_file_name: str = "wrm_{:09d}.png"
    mic_file_path: str = field(init=False)
    cam_file_path: str = field(init=False)
    err_file_path: str = field(init=False)
    wrm_file_path: str = field(init=False)
    bbox_file_path: str = field(init=False)

    def __post_init__(self):
        self.mic_file_path = _join(self.root_folder, self.mic_folder_name, self.mic_file_name)
        self.cam_file_path = _join(self.root_folder, self.cam_folder_name, self.cam_file_name)
        self.err_file_path = _join(self.root_folder, self.err_folder_name, self.cam_file_name)
        self.wrm_file_path = _join(self.root_folder, self.wrm_folder_name, self.wrm_file_name)
        self.bbox_file_path = _join(self.root_folder, self.bbox_file_name)

    def create_dirs(self) -> None:
        for p in (self.bbox_file_path, self.mic_file_path, self.cam_file_path, self.err_file_path, self.wrm_file_path):
            os.makedirs(os.path.dirname(p), exist_ok=True)


class _Saver:
    """One worker thread writing images with cv2 (stands in for the reference's ImageSaver / FrameSaver,
    wtracker/utils/io_utils.py:11-107; image files are outside the accelerated path)."""

    def __init__(self, reader=None):
        self._reader = reader
        self._q: queue.Queue = queue.Queue(maxsize=100)
        self._thread = threading.Thread(target=self._work, daemon=True)

    def start(self):
        self._thread.start()

    def _work(self):
        import cv2 as cv

        while True:
            item = self._q.get()
            if item is None:
                return
            img, path = item
            if isinstance(img, tuple):                     # (frame index, crop xywh) of the reader
                idx, (x, y, w, h) = img
                img = self._reader[idx][y:y + h, x:x + w]
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            if not cv.imwrite(path, img):
                raise ValueError(f"Failed to save image {path}")

    def schedule_save(self, img, path: str):
        self._q.put((img, path))

    def close(self):
        if self._thread.is_alive():
            self._q.put(None)
            self._thread.join()


def log_table_device(worm_rel: torch.Tensor, cam_xywh: torch.Tensor, mic_xywh: torch.Tensor, plt_xy: torch.Tensor,
                     first_frame: int, cycle_frame_num: int, imaging_frame_num: int, bounds: tuple[int, int]):
    """Rows of bboxes.csv for n consecutive frames, all on the device: ``worm_rel`` f64 | f32 [n][4] camera-relative
    boxes (NaN = no prediction), i32 camera / microscope boxes [n][4], i32 platform positions [n][2] ->
    (table f64 [n][17] in LOG_COLUMNS order with phase 0 | 1, crop i32 [n][4], legal u8 [n])."""
    if not worm_rel.is_cuda:
        raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
    assert worm_rel.dtype in (torch.float64, torch.float32)
    n = worm_rel.shape[0]
    dev = worm_rel.device
    worm_rel = worm_rel.contiguous()
    cam_xywh, mic_xywh, plt_xy = (t.to(device=dev, dtype=torch.int32).contiguous() for t in (cam_xywh, mic_xywh, plt_xy))
    assert cam_xywh.shape == (n, 4) and mic_xywh.shape == (n, 4) and plt_xy.shape == (n, 2)
    table = torch.empty((n, 17), dtype=torch.float64, device=dev)
    crop = torch.empty((n, 4), dtype=torch.int32, device=dev)
    legal = torch.empty((n,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().wt_log_rows(worm_rel.data_ptr(), int(worm_rel.dtype == torch.float32), cam_xywh.data_ptr(),
                                    mic_xywh.data_ptr(), plt_xy.data_ptr(), n, int(first_frame), int(cycle_frame_num),
                                    int(imaging_frame_num), int(bounds[0]), int(bounds[1]), table.data_ptr(),
                                    crop.data_ptr(), legal.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "wt_log_rows")
    return table, crop, legal


def csv_rows(table: np.ndarray, worm_dtype) -> list[dict]:
    """The dicts ``CSVLogger`` prints for a host copy of the table: integer columns as Python ints, phase as the
    reference's strings, wrm columns as numpy scalars of the controller's dtype (their ``str`` is the csv text)."""
    rows = []
    wrm = table[:, 13:17].astype(worm_dtype)
    for r, w in zip(table, wrm):
        row = {k: int(v) for k, v in zip(LOG_COLUMNS[3:13], r[3:13])}
        row["cycle"] = int(r[1])
        row["frame"] = int(r[0])
        row["phase"] = "imaging" if r[2] == 0 else "moving"
        row["wrm_x"], row["wrm_y"], row["wrm_w"], row["wrm_h"] = w
        rows.append(row)
    return rows


class LoggingController(SimController):
    device = "cuda:0"

    def __init__(self, sim_controller: SimController, log_config: LogConfig):
        super().__init__(sim_controller.timing_config)
        self.sim_controller = sim_controller
        self.log_config = log_config
        n = self.timing_config.cycle_frame_num
        self._camera_frames = deque(maxlen=n)
        self._platform_positions = deque(maxlen=n)
        self._camera_bboxes = deque(maxlen=n)
        self._micro_bboxes = deque(maxlen=n)

    def on_sim_start(self, sim: Simulator):
        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        self.sim_controller.on_sim_start(sim)
        self._camera_frames.clear()
        self._platform_positions.clear()
        self._camera_bboxes.clear()
        self._micro_bboxes.clear()
        self.log_config.create_dirs()
        self._image_saver = _Saver()
        self._image_saver.start()
        # (the reference copies the reader for its FrameSaver unconditionally; an in-memory reader is hundreds of MB, and
        # the copy is only ever read when worm views are saved)
        self._frame_saver = None
        if self.log_config.save_wrm_view:
            self._frame_saver = _Saver(deepcopy(sim.view._frame_reader))
            self._frame_saver.start()
        self._stage = None
        self._bbox_logger = CSVLogger(self.log_config.bbox_file_path, col_names=list(LOG_COLUMNS))

    def on_cycle_start(self, sim: Simulator):
        self.sim_controller.on_cycle_start(sim)

    def on_camera_frame(self, sim: Simulator):
        self.sim_controller.on_camera_frame(sim)
        self._platform_positions.append(sim.position)
        self._camera_bboxes.append(sim.view.camera_position)
        self._micro_bboxes.append(sim.view.micro_position)
        if self.log_config.save_err_view:
            self._camera_frames.append(sim.camera_view())
        if self.log_config.save_cam_view:
            self._image_saver.schedule_save(sim.camera_view(), self.log_config.cam_file_path.format(sim.frame_number))
        if self.log_config.save_mic_view:
            self._image_saver.schedule_save(sim.view.micro_view(), self.log_config.mic_file_path.format(sim.frame_number))

    def _staging(self, n: int):
        """Persistent buffers of the per-cycle log call: ONE pinned input block (worm boxes | camera | microscope boxes |
        platform positions) with its device mirror, ONE device output block (table | crop | legal) with its pinned
        mirror — a cycle is one H2D, one kernel, one D2H and one stream synchronize, nothing allocated."""
        if self._stage is None or self._stage["n"] != n:
            dev = torch.device(self.device)
            al = lambda v: (v + 255) & ~255                    # every block 256-byte aligned (the kernel uses 16-byte accesses)
            o_cam = al(32 * n)
            o_mic = o_cam + al(16 * n)
            o_plt = o_mic + al(16 * n)
            in_bytes = o_plt + al(8 * n)
            o_crop = al(136 * n)
            o_legal = o_crop + al(16 * n)
            out_bytes = o_legal + al(n)
            h_in = torch.zeros(in_bytes, dtype=torch.uint8).pin_memory()
            h_out = torch.zeros(out_bytes, dtype=torch.uint8).pin_memory()
            d_in = torch.zeros(in_bytes, dtype=torch.uint8, device=dev)
            d_out = torch.zeros(out_bytes, dtype=torch.uint8, device=dev)
            hn, on = h_in.numpy(), h_out.numpy()
            self._stage = dict(
                n=n, h_in=h_in, h_out=h_out, d_in=d_in, d_out=d_out, o_cam=o_cam, o_mic=o_mic, o_plt=o_plt, o_crop=o_crop,
                o_legal=o_legal,
                worm64=hn[: 32 * n].view(np.float64).reshape(n, 4), worm32=hn[: 16 * n].view(np.float32).reshape(n, 4),
                cam=hn[o_cam: o_cam + 16 * n].view(np.int32).reshape(n, 4),
                mic=hn[o_mic: o_mic + 16 * n].view(np.int32).reshape(n, 4),
                plt=hn[o_plt: o_plt + 8 * n].view(np.int32).reshape(n, 2),
                table=on[: 136 * n].view(np.float64).reshape(n, 17),
                crop=on[o_crop: o_crop + 16 * n].view(np.int32).reshape(n, 4), legal=on[o_legal: o_legal + n])
        return self._stage

    def _log_cycle(self, sim: Simulator):
        cycle_number = sim.cycle_number - 1
        n_cyc = self.timing_config.cycle_frame_num
        frame_offset = cycle_number * n_cyc
        worm = np.asarray(self.sim_controller._cycle_predict_all(sim))
        dtype = worm.dtype if worm.dtype in (np.float32, np.float64) else np.dtype(np.float64)
        n = worm.shape[0]
        assert len(self._camera_bboxes) == len(self._micro_bboxes) == len(self._platform_positions) == n
        st = self._staging(n)
        (st["worm32"] if dtype == np.float32 else st["worm64"])[...] = worm
        st["cam"][...] = self._camera_bboxes
        st["mic"][...] = self._micro_bboxes
        st["plt"][...] = self._platform_positions
        with torch.cuda.device(torch.device(self.device)):
            stream = torch.cuda.current_stream()
            st["d_in"].copy_(st["h_in"], non_blocking=True)
            i0, o0 = st["d_in"].data_ptr(), st["d_out"].data_ptr()
            bounds = sim.experiment_config.orig_resolution
            L.check(L.lib().wt_log_rows(i0, int(dtype == np.float32), i0 + st["o_cam"], i0 + st["o_mic"], i0 + st["o_plt"], n,
                                        int(frame_offset), int(n_cyc), int(self.timing_config.imaging_frame_num),
                                        int(bounds[0]), int(bounds[1]), o0, o0 + st["o_crop"], o0 + st["o_legal"],
                                        stream.cuda_stream),
                    "wt_log_rows")
            st["h_out"].copy_(st["d_out"], non_blocking=True)
            stream.synchronize()
        table, crop, legal = st["table"], st["crop"], st["legal"].astype(bool)
        if self.log_config.save_wrm_view:
            for i in np.nonzero(legal)[0]:
                frame_number = frame_offset + int(i)
                self._frame_saver.schedule_save((frame_number, tuple(int(v) for v in crop[i])),
                                                self.log_config.wrm_file_path.format(frame_number))
        # (the reference's error-view branch tests the rows AFTER discretize zeroed them, so it never saves anything)
        self._bbox_logger.writerows(csv_rows(table, dtype))
        self._bbox_logger.flush()

    def on_cycle_end(self, sim: Simulator):
        self._log_cycle(sim)
        self.sim_controller.on_cycle_end(sim)
        self._camera_frames.clear()
        self._platform_positions.clear()
        self._camera_bboxes.clear()
        self._micro_bboxes.clear()

    def on_sim_end(self, sim: Simulator):
        self.sim_controller.on_sim_end(sim)
        self._image_saver.close()
        if self._frame_saver is not None:
            self._frame_saver.close()
        self._bbox_logger.close()

    def on_imaging_start(self, sim: Simulator):
        self.sim_controller.on_imaging_start(sim)

    def on_micro_frame(self, sim: Simulator):
        self.sim_controller.on_micro_frame(sim)

    def on_imaging_end(self, sim: Simulator):
        self.sim_controller.on_imaging_end(sim)

    def on_movement_start(self, sim: Simulator):
        self.sim_controller.on_movement_start(sim)

    def on_movement_end(self, sim: Simulator):
        self.sim_controller.on_movement_end(sim)

    def begin_movement_prediction(self, sim: Simulator) -> None:
        return self.sim_controller.begin_movement_prediction(sim)

    def provide_movement_vector(self, sim: Simulator) -> tuple[int, int]:
        return self.sim_controller.provide_movement_vector(sim)

    def _cycle_predict_all(self, sim: Simulator) -> np.ndarray:
        return self.sim_controller._cycle_predict_all(sim)
