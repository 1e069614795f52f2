"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the rows LoggingController._log_cycle writes to bboxes.csv
(reference: wtracker/sim/sim_controllers/logging_controller.py:145-185) and of BoxUtils.discretize
(wtracker/utils/bbox_utils.py:119-167).  Only tests/ may import this; the product path is wt_log_rows (CUDA).

Pinned: tests/golden/reference_bboxes_{f64,f32}.csv were written by the UNMODIFIED reference LoggingController
(tests/golden/make_golden_log.py); tests/test_log_cpu.py formats these rows and compares the text byte for byte.
"""
from __future__ import annotations

import numpy as np

LOG_COLUMNS = ["frame", "cycle", "phase", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w",
               "mic_h", "wrm_x", "wrm_y", "wrm_w", "wrm_h"]


def log_rows(worm_rel: np.ndarray, cam: np.ndarray, mic: np.ndarray, plt: np.ndarray, first_frame: int,
             cycle_frame_num: int, imaging_frame_num: int, bounds: tuple[int, int]):
    """(table [n][17] in the dtype of ``worm_rel`` for the wrm columns, crop i32 [n][4], legal bool [n])."""
    worm = np.array(worm_rel, copy=True)
    cam = np.asarray(cam, dtype=np.int64)
    worm[:, 0] += cam[:, 0]            # :153  (float32 arrays: computed in float64, cast back)
    worm[:, 1] += cam[:, 1]            # :154
    finite = np.isfinite(worm).all(axis=1)
    worm[~finite] = 0                  # bbox_utils.py:141 — in the caller's array, before the rows are written
    x1, y1 = worm[:, 0], worm[:, 1]
    x2, y2 = x1 + worm[:, 2], y1 + worm[:, 3]
    H, W = bounds
    ix1 = np.clip(np.floor(x1).astype(np.int32), 0, W)
    iy1 = np.clip(np.floor(y1).astype(np.int32), 0, H)
    ix2 = np.clip(np.ceil(x2).astype(np.int32), 0, W)
    iy2 = np.clip(np.ceil(y2).astype(np.int32), 0, H)
    legal = ((ix2 - ix1) > 0) & ((iy2 - iy1) > 0)
    crop = np.stack([ix1, iy1, ix2 - ix1, iy2 - iy1], 1).astype(np.int32)
    crop[~legal] = 0
    n = worm.shape[0]
    frame = first_frame + np.arange(n)
    table = np.zeros((n, 17), dtype=np.float64)
    table[:, 0] = frame
    table[:, 1] = frame // cycle_frame_num
    table[:, 2] = (frame % cycle_frame_num) >= imaging_frame_num
    table[:, 3:5] = plt
    table[:, 5:9] = cam
    table[:, 9:13] = mic
    table[:, 13:17] = worm
    return table, crop, legal
