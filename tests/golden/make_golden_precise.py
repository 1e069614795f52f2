"""Generates tests/golden/reference_precise.npz by running the UNMODIFIED reference
ErrorCalculator.calculate_precise (/root/reference/wtracker/eval/error_calculator.py:63-161) on seeded synthetic
frames.  Run in the build container only:  python tests/golden/make_golden_precise.py

The worm views handed to the reference are the crops of the frames at the discretized worm boxes — what
LoggingController saves as wrm_*.png (logging_controller.py:169-173, io_utils.py:47-56) and the analysis reads back.
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, "/root/reference")
for name in ("tkinter", "tkinter.filedialog", "seaborn"):
    _m = sys.modules.setdefault(name, types.ModuleType(name))
    _m.Tk = object
sys.modules["tkinter"].filedialog = sys.modules["tkinter.filedialog"]
_pkg = types.ModuleType("wtracker.eval")
_pkg.__path__ = ["/root/reference/wtracker/eval"]
sys.modules["wtracker.eval"] = _pkg

from wtracker.eval.error_calculator import ErrorCalculator  # noqa: E402
from wtracker.utils.bbox_utils import BoxFormat, BoxUtils  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(77)
H, W, F, N = 180, 320, 8, 160
background = np.clip(200 + rng.normal(0, 3, (H, W)), 0, 255).astype(np.uint8)
frames = np.repeat(background[None], F, 0).copy()
centres = np.stack([rng.uniform(30, W - 30, F), rng.uniform(30, H - 30, F)], 1)
for f in range(F):                                   # a dark blob (the worm head) per frame + sensor noise
    yy, xx = np.mgrid[0:H, 0:W]
    blob = ((xx - centres[f, 0]) / 7.0) ** 2 + ((yy - centres[f, 1]) / 4.0) ** 2 < 1.0
    frames[f][blob] = 60
    frames[f] = np.clip(frames[f].astype(int) + rng.integers(-14, 15, (H, W)), 0, 255).astype(np.uint8)

frame_nums = rng.integers(0, F, N)
worm = np.stack([centres[frame_nums, 0] - 9 + rng.normal(0, 2, N), centres[frame_nums, 1] - 7 + rng.normal(0, 2, N),
                 18 + rng.normal(0, 1.5, N), 14 + rng.normal(0, 1.5, N)], 1)
mic = np.stack([worm[:, 0] + rng.normal(0, 7, N), worm[:, 1] + rng.normal(0, 5, N), np.full(N, 14.0), np.full(N, 14.0)], 1)
worm[::23] = np.nan                                   # no prediction
worm[5::31, 2] = -1.0                                 # empty after rounding / clipping
worm[7::41, 0] = -40.0                                # entirely left of the frame
worm[3::29, :2] = np.round(worm[3::29, :2])           # exact integer corners
mic[11::37, 0] += 60                                  # microscope far away: everything outside
mic[13::43] = np.nan
worm[150] = np.nan                                    # an illegal row BEHIND the last compacted result stays NaN


class CropReader:
    """worm_reader: the view of row k is the frame cropped at its discretized worm box (indexed by row)."""

    def __init__(self, crops):
        self.crops = crops

    def __getitem__(self, k):
        return self.crops[k]


disc, legal = BoxUtils.discretize(worm.copy(), (H, W), BoxFormat.XYWH)
crops = [frames[frame_nums[k]][disc[k, 1]:disc[k, 1] + disc[k, 3], disc[k, 0]:disc[k, 0] + disc[k, 2]] for k in range(N)]
out = {}
for thr in (10, 20.5):
    with np.errstate(all="ignore"):
        out[f"err_thr{thr}"] = ErrorCalculator.calculate_precise(background, worm.copy(), mic.copy(), np.arange(N),
                                                                CropReader(crops), diff_thresh=thr)
np.savez_compressed(os.path.join(HERE, "reference_precise.npz"), background=background, frames=frames,
                    frame_nums=frame_nums, worm=worm, mic=mic, legal=legal, **out)
print({k: (v.shape, np.isnan(v).sum(), float(np.nanmean(v))) for k, v in out.items()}, "legal", legal.sum())
