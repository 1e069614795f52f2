"""Generates tests/golden/reference_analysis.npz by running the UNMODIFIED reference DataAnalyzer.initialize
(/root/reference/wtracker/eval/data_analyzer.py:54-107) on the committed golden log tests/golden/reference_bboxes_f64.csv
(itself written by the unmodified LoggingController).  Run in the build container only:

    python tests/golden/make_golden_analysis.py
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

from wtracker.eval.data_analyzer import DataAnalyzer  # noqa: E402
from wtracker.sim.config import ExperimentConfig, TimingConfig  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
exp = ExperimentConfig("g", 400, 60, (1080, 1920), 90, (960, 540))
t = TimingConfig(exp, 100, 40, 50, (4.0, 4.0), (0.32, 0.32))
out = {}
for period in (10, 3):
    an = DataAnalyzer.load(t, os.path.join(HERE, "reference_bboxes_f64.csv"))
    with np.errstate(all="ignore"):
        an.initialize(period=period)
    df = an.data
    out[f"columns_p{period}"] = np.array(list(df.columns))
    num = df.drop(columns=["phase"])
    out[f"names_p{period}"] = np.array(list(num.columns))
    out[f"values_p{period}"] = num.to_numpy(dtype=np.float64)
np.savez_compressed(os.path.join(HERE, "reference_analysis.npz"), **out)
print({k: v.shape for k, v in out.items()}, list(out["names_p10"]))

# ---- DataAnalyzer.clean / calc_anomalies (data_analyzer.py:121-159, 326-374) on the analysed table; some rows get NaN
# worm boxes first (a log never holds NaN, but tables assembled in memory do) so that both branches of `has_pred` run
FLAGS = ["speed_anomaly", "bbox_error_anomaly", "dist_error_anomaly", "width_anomaly", "height_anomaly", "no_pred_anomaly"]
NAN_ROWS = np.array([5, 6, 77, 140, 141, 142, 300])
CLEAN_CASES = {
    "trim": dict(trim_cycles=True),
    "imaging": dict(imaging_only=True),
    "bounds": dict(bounds=(330.0, 200.0, 640.0, 470.0)),
    "all": dict(trim_cycles=True, imaging_only=True, bounds=(320.5, 190.25, 660.0, 480.0)),
}
ANOMALY_CASES = {
    "a": dict(no_preds=True, min_bbox_error=0.95, min_dist_error=40.0, min_speed=4.0, min_size=14.9),
    "b": dict(no_preds=False, min_bbox_error=0.9),
    "c": dict(no_preds=True, min_speed=2.5, remove_anomalies=True),
}
masks = {"nan_rows": NAN_ROWS}


def fresh():
    an = DataAnalyzer.load(t, os.path.join(HERE, "reference_bboxes_f64.csv"))
    with np.errstate(all="ignore"):
        an.initialize(period=10)
    an.data.loc[NAN_ROWS, ["wrm_x", "wrm_y", "wrm_w", "wrm_h"]] = np.nan
    return an


for name, kw in CLEAN_CASES.items():
    an = fresh()
    an.clean(**kw)
    masks[f"clean_{name}"] = an.data["frame"].to_numpy(dtype=np.int64)
for name, kw in ANOMALY_CASES.items():
    an = fresh()
    an.clean(imaging_only=(name == "b"))
    res = an.calc_anomalies(**kw)
    masks[f"anom_{name}_frames"] = res["frame"].to_numpy(dtype=np.int64)
    masks[f"anom_{name}_flags"] = res[FLAGS].to_numpy(dtype=bool)
    masks[f"anom_{name}_left"] = an.data["frame"].to_numpy(dtype=np.int64)
np.savez_compressed(os.path.join(HERE, "reference_masks.npz"), **masks)
print({k: v.shape for k, v in masks.items()})
