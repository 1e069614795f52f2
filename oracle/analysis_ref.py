"""TEST INFRASTRUCTURE — CPU restatement (numpy) of DataAnalyzer.initialize (reference:
wtracker/eval/data_analyzer.py:54-107).  Only tests/ may import this; the product path is wt_analysis_columns (CUDA).

Pinned: tests/golden/reference_analysis.npz holds the DataFrame the UNMODIFIED reference produces from the golden
bboxes.csv (tests/golden/make_golden_analysis.py).
"""
from __future__ import annotations

import numpy as np

from oracle.metrics_ref import bbox_error

NAMES = ["frame", "cycle", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w", "mic_h", "wrm_x",
         "wrm_y", "wrm_w", "wrm_h", "time", "cycle_step", "wrm_center_x", "wrm_center_y", "mic_center_x", "mic_center_y",
         "wrm_speed_x", "wrm_speed_y", "wrm_speed", "worm_deviation_x", "worm_deviation_y", "worm_deviation", "bbox_error",
         "precise_error"]


def analysis_columns(table17: np.ndarray, period: int, cycle_frame_num: int) -> np.ndarray:
    t = np.asarray(table17, dtype=np.float64)
    n = len(t)
    frame = t[:, 0]
    wrm, mic = t[:, 13:17], t[:, 9:13]
    wcx, wcy = wrm[:, 0] + wrm[:, 2] / 2, wrm[:, 1] + wrm[:, 3] / 2                       # :78-82
    mcx, mcy = mic[:, 0] + mic[:, 2] / 2, mic[:, 1] + mic[:, 3] / 2

    def lag(a):                                                                            # Series.diff(n)
        d = np.full(n, np.nan)
        d[period:] = a[period:] - a[:-period]
        return d

    with np.errstate(all="ignore"):
        dt = lag(frame)
        sx, sy = lag(wcx) / dt, lag(wcy) / dt                                              # :86-88
        speed = np.sqrt(sx ** 2 + sy ** 2)
        dx, dy = wcx - mcx, wcy - mcy                                                      # :93-95
        dev = np.sqrt(dx ** 2 + dy ** 2)
        err = bbox_error(wrm.copy(), mic.copy())                                           # :100-104
    out = np.empty((n, 30))
    out[:, 0:2] = t[:, 0:2]
    out[:, 2:12] = t[:, 3:13]
    out[:, 12:16] = wrm
    out[:, 16] = frame
    out[:, 17] = frame.astype(np.int64) % cycle_frame_num
    for k, col in enumerate((wcx, wcy, mcx, mcy, sx, sy, speed, dx, dy, dev, err)):
        out[:, 18 + k] = col
    out[:, 29] = np.nan
    out[:, 12:16] = np.round(out[:, 12:16], 5)                                             # DataFrame.round(5), :69
    out[:, 18:29] = np.round(out[:, 18:29], 5)
    return out


def clean_keep(table30: np.ndarray, moving: np.ndarray, imaging_only: bool = False, bounds=None,
               trim_cycles: bool = False) -> np.ndarray:
    """Rows DataAnalyzer.clean keeps (reference: wtracker/eval/data_analyzer.py:121-159), as a boolean mask."""
    t = np.asarray(table30, dtype=np.float64)
    keep = np.ones(len(t), dtype=bool)
    if imaging_only:
        keep &= ~np.asarray(moving, dtype=bool)
    if bounds is not None:
        wrm, mic = t[:, 12:16], t[:, 8:12]
        has_pred = np.isfinite(wrm).all(axis=1)
        with np.errstate(invalid="ignore"):
            # `mask_wrm = has_pred; mask_wrm &= <x condition>` narrows has_pred in place (:140-142; the result is a Series,
            # so the y condition of :143 does not reach has_pred): `~has_pred` at :145 is "no prediction or x range failed"
            in_wx = has_pred & (wrm[:, 0] >= bounds[0]) & (wrm[:, 0] + wrm[:, 2] <= bounds[2])
            in_w = in_wx & (wrm[:, 1] >= bounds[1]) & (wrm[:, 1] + wrm[:, 3] <= bounds[3])
            in_m = ~in_wx & (mic[:, 0] >= bounds[0]) & (mic[:, 0] + mic[:, 2] <= bounds[2]) & (mic[:, 1] >= bounds[1]) & \
                (mic[:, 1] + mic[:, 3] <= bounds[3])
        keep &= in_w | in_m
    if trim_cycles and keep.any():
        cyc = t[:, 1]
        keep &= (cyc != 0) & (cyc != cyc[keep].max())                                     # :150-153
    return keep


def anomaly_bits(table30: np.ndarray, no_preds=True, min_bbox_error=np.inf, min_dist_error=np.inf, min_speed=np.inf,
                 min_size=np.inf) -> np.ndarray:
    """calc_anomalies' six masks (:349-361) packed as bits 0..5: speed, bbox error, distance, width, height, no prediction."""
    t = np.asarray(table30, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        flags = [t[:, 24] >= min_speed, t[:, 28] >= min_bbox_error, t[:, 27] >= min_dist_error, t[:, 14] >= min_size,
                 t[:, 15] >= min_size, (~np.isfinite(t[:, 12:16]).all(axis=1)) & bool(no_preds)]
    return sum((f.astype(np.uint8) << i) for i, f in enumerate(flags)).astype(np.uint8)
