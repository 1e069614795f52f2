"""calculate_precise on the GPU: the drop-in ErrorCalculator.calculate_precise (worm views through a reader, as the
reference is called) reproduces the UNMODIFIED reference's array exactly, and the device form (frames in HBM) equals
the oracle row for row, also on a large hostile table."""
import os

import numpy as np
import pytest
import torch

from oracle import precise_ref
from wtracker_b200.eval.error_calculator import ErrorCalculator
from wtracker_b200.utils.bbox_utils import BoxFormat, BoxUtils

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_precise.npz")


class CropReader:
    def __init__(self, crops):
        self.crops = crops

    def __getitem__(self, k):
        return self.crops[k]


@pytest.mark.parametrize("thr", [10, 20.5])
def test_drop_in_call_matches_reference(thr):
    g = np.load(GOLD)
    frames, fn, worm = g["frames"], g["frame_nums"], g["worm"]
    H, W = g["background"].shape
    disc, _ = BoxUtils.discretize(worm.copy(), (H, W), BoxFormat.XYWH)
    crops = [frames[fn[k]][disc[k, 1]:disc[k, 1] + disc[k, 3], disc[k, 0]:disc[k, 0] + disc[k, 2]] for k in range(len(fn))]
    got = ErrorCalculator.calculate_precise(g["background"], worm.copy(), g["mic"].copy(), np.arange(len(fn)),
                                            CropReader(crops), diff_thresh=thr)
    want = g[f"err_thr{thr}"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])


@pytest.mark.parametrize("n", [1, 160, 20000])
def test_device_form_equals_oracle(n):
    g = np.load(GOLD)
    rng = np.random.default_rng(n)
    frames, bg = g["frames"], g["background"]
    F, H, W = frames.shape
    fidx = rng.integers(0, F, n).astype(np.int32)
    worm = np.stack([rng.uniform(-20, W, n), rng.uniform(-20, H, n), rng.uniform(-2, 40, n), rng.uniform(-2, 40, n)], 1)
    mic = np.stack([worm[:, 0] + rng.normal(0, 10, n), worm[:, 1] + rng.normal(0, 10, n), np.full(n, 14.0), np.full(n, 14.0)], 1)
    worm[rng.uniform(size=n) < 0.05] = np.nan
    mic[rng.uniform(size=n) < 0.03] = np.nan
    want, legal = precise_ref.precise_error(frames, fidx, bg, worm, mic, 12)
    got = ErrorCalculator.calculate_precise_device(torch.from_numpy(frames).cuda(), torch.from_numpy(fidx),
                                                   torch.from_numpy(bg).cuda(), torch.from_numpy(worm),
                                                   torch.from_numpy(mic), 12).cpu().numpy()
    assert np.array_equal(np.isnan(got), ~legal)
    assert np.array_equal(got[legal], want[legal])
