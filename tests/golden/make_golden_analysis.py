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
