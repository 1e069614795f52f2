"""TEST INFRASTRUCTURE — CPU restatement (numpy) of ErrorCalculator.calculate_precise / calculate_segmentation
(reference: wtracker/eval/error_calculator.py:19-161) with the worm views taken as crops of the frames at the
discretized worm boxes.  Only tests/ may import this; the product path is wt_precise_error (CUDA).

Pinned: tests/golden/reference_precise.npz holds the outputs of the UNMODIFIED reference on seeded frames
(tests/golden/make_golden_precise.py).  Returns the per-row errors in the rows' OWN positions; `compact()` applies
the reference's indexing quirk (results of the legal rows are written to errors[0..n_legal), :131-159).
"""
from __future__ import annotations

import numpy as np


def _discretize(b: np.ndarray, H: int, W: int):
    b = np.array(b, dtype=np.float64, copy=True)
    b[~np.isfinite(b).all(axis=1)] = 0
    x1 = np.clip(np.floor(b[:, 0]).astype(np.int32), 0, W)
    y1 = np.clip(np.floor(b[:, 1]).astype(np.int32), 0, H)
    x2 = np.clip(np.ceil(b[:, 0] + b[:, 2]).astype(np.int32), 0, W)
    y2 = np.clip(np.ceil(b[:, 1] + b[:, 3]).astype(np.int32), 0, H)
    legal = ((x2 - x1) > 0) & ((y2 - y1) > 0)
    out = np.stack([x1, y1, x2 - x1, y2 - y1], 1)
    out[~legal] = 0
    return out, legal


def precise_error(frames: np.ndarray, frame_idx: np.ndarray, background: np.ndarray, worm: np.ndarray, mic: np.ndarray,
                  diff_thresh: float = 10) -> tuple[np.ndarray, np.ndarray]:
    H, W = background.shape[:2]
    wb, legal = _discretize(worm, H, W)
    mb, _ = _discretize(mic, H, W)
    err = np.full(len(wb), np.nan)
    for i in np.nonzero(legal)[0]:
        x, y, w, h = wb[i]
        view = frames[frame_idx[i]][y:y + h, x:x + w]
        diff = np.abs(view.astype(np.int32) - background[y:y + h, x:x + w].astype(np.int32)).astype(np.uint8)
        mask = diff > diff_thresh
        il, it = max(x, mb[i, 0]), max(y, mb[i, 1])
        ir, ib = min(x + w, mb[i, 0] + mb[i, 2]), min(y + h, mb[i, 1] + mb[i, 3])
        iw, ih = max(0, ir - il), max(0, ib - it)
        mm = np.zeros_like(mask)
        mm[it - y:it - y + ih, il - x:il - x + iw] = True
        total = mask.sum()
        err[i] = 0.0 if total == 0 else 1.0 - np.logical_and(mask, mm).sum() / total
    return err, legal


def compact(err: np.ndarray, legal: np.ndarray) -> np.ndarray:
    """The array the reference returns: NaN at illegal rows, zeros elsewhere, then the results of the legal rows
    written to positions 0 .. n_legal-1 (its loop indexes `errors` with the index into the FILTERED arrays)."""
    out = np.zeros(len(err))
    out[~legal] = np.nan
    out[:int(legal.sum())] = err[legal]
    return out
