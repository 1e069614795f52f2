"""ORACLE (test infrastructure, not product code) — numpy restatement of the per-step metrics:
ErrorCalculator.calculate_bbox_error (/root/reference/wtracker/eval/error_calculator.py:163-195) and
calculate_mse_error (:197-212, with BoxUtils.center, utils/bbox_utils.py:77-92).  float64.
Pinned by tests/golden/reference_golden.npz (made by the real reference)."""

from __future__ import annotations

import numpy as np


def bbox_error(worm: np.ndarray, mic: np.ndarray) -> np.ndarray:
    worm, mic = np.asarray(worm, dtype=np.float64), np.asarray(mic, dtype=np.float64)
    wl, wt, ww, wh = worm.T
    ml, mt, mw, mh = mic.T
    with np.errstate(all="ignore"):
        iw = np.maximum(0, np.minimum(wl + ww, ml + mw) - np.maximum(wl, ml))
        ih = np.maximum(0, np.minimum(wt + wh, mt + mh) - np.maximum(wt, mt))
        total = ww * wh
        err = 1.0 - (iw * ih) / total
    err[total == 0] = 0.0
    return err


def mse_error(worm: np.ndarray, mic: np.ndarray) -> np.ndarray:
    worm, mic = np.asarray(worm, dtype=np.float64), np.asarray(mic, dtype=np.float64)
    wc = np.stack([worm[:, 0] + worm[:, 2] / 2, worm[:, 1] + worm[:, 3] / 2], 1)
    mc = np.stack([mic[:, 0] + mic[:, 2] / 2, mic[:, 1] + mic[:, 3] / 2], 1)
    return np.mean((wc - mc) ** 2, axis=1)
