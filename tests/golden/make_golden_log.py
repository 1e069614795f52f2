"""Generates tests/golden/reference_bboxes_{f64,f32}.csv by running the UNMODIFIED reference LoggingController
(/root/reference/wtracker/sim/sim_controllers/logging_controller.py) around its CsvController on the seeded
track of make_golden.py.  Run in the build container only:

    python tests/golden/make_golden_log.py

Two variants: the controller's `_cycle_predict_all` returns float64 (CsvController) or float32 (what
YoloController.predict returns when every frame has a detection, yolo_controller.py:85-90) — the CSV text
differs because str(np.float32) is shorter.
"""
import os
import shutil
import sys
import tempfile
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
_u = types.ModuleType("ultralytics")
_u.YOLO = object
sys.modules.setdefault("ultralytics", _u)

import pandas as pd  # noqa: E402
from wtracker.sim.config import ExperimentConfig, TimingConfig  # noqa: E402
from wtracker.sim.sim_controllers.csv_controller import CsvController  # noqa: E402
from wtracker.sim.sim_controllers.logging_controller import LogConfig, LoggingController  # noqa: E402
from wtracker.sim.simulator import Simulator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
tab = np.load(os.path.join(HERE, "reference_golden.npz"))["trace_csv_table"][:400]


class F32Controller(CsvController):
    """CsvController whose cycle predictions are float32 with no missing rows (the YoloController dtype quirk)."""

    def _cycle_predict_all(self, sim):
        return np.nan_to_num(super()._cycle_predict_all(sim), nan=7.25).astype(np.float32)


tmp = tempfile.mkdtemp()
csv_path = os.path.join(tmp, "track.csv")
pd.DataFrame(tab, columns=["wrm_x", "wrm_y", "wrm_w", "wrm_h"]).to_csv(csv_path, index=False)
for tag, cls in (("f64", CsvController), ("f32", F32Controller)):
    exp = ExperimentConfig("g", 400, 60, (1080, 1920), 90, (960, 540))
    t = TimingConfig(exp, 100, 40, 50, (4.0, 4.0), (0.32, 0.32))
    root = os.path.join(tmp, tag)
    log = LogConfig(root, save_mic_view=False, save_cam_view=False, save_err_view=False, save_wrm_view=False)
    Simulator(t, exp, LoggingController(cls(t, csv_path), log)).run()
    shutil.copy(os.path.join(root, "bboxes.csv"), os.path.join(HERE, f"reference_bboxes_{tag}.csv"))
    print(tag, sum(1 for _ in open(os.path.join(HERE, f"reference_bboxes_{tag}.csv"))), "lines")
