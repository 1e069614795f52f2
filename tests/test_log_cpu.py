"""bboxes.csv wire format on the CPU: the oracle restatement of LoggingController._log_cycle
(oracle/log_ref.py) driven by the host Simulator + CsvController, formatted through csv_rows + CSVLogger, must
reproduce the csv the UNMODIFIED reference wrote (tests/golden/reference_bboxes_*.csv, make_golden_log.py) byte for
byte — float64 and float32 prediction arrays, missed detections logged as zeros, last cycle never logged."""
import os

import numpy as np
import pytest

from oracle import log_ref
from wtracker_b200.sim import ExperimentConfig, Simulator, TimingConfig
from wtracker_b200.sim.simulator import SimController
from wtracker_b200.sim.sim_controllers.csv_controller import CsvController
from wtracker_b200.sim.sim_controllers.logging_controller import LOG_COLUMNS, LogConfig, csv_rows
from wtracker_b200.utils.log_utils import CSVLogger

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class F32Controller(CsvController):
    def _cycle_predict_all(self, sim):
        return np.nan_to_num(super()._cycle_predict_all(sim), nan=7.25).astype(np.float32)


class OracleLogger(SimController):
    """LoggingController's bookkeeping with the numpy oracle in place of the CUDA kernel."""

    def __init__(self, inner, path):
        super().__init__(inner.timing_config)
        self.inner, self.path = inner, path
        self.plt, self.cam, self.mic = [], [], []

    def on_sim_start(self, sim):
        self.inner.on_sim_start(sim)
        self.log = CSVLogger(self.path, list(LOG_COLUMNS))

    def on_camera_frame(self, sim):
        self.inner.on_camera_frame(sim)
        self.plt.append(sim.position)
        self.cam.append(sim.view.camera_position)
        self.mic.append(sim.view.micro_position)

    def on_cycle_end(self, sim):
        worm = self.inner._cycle_predict_all(sim)
        n = self.timing_config.cycle_frame_num
        table, crop, legal = log_ref.log_rows(worm, np.array(self.cam), np.array(self.mic), np.array(self.plt),
                                              (sim.cycle_number - 1) * n, n, self.timing_config.imaging_frame_num,
                                              sim.experiment_config.orig_resolution)
        self.log.writerows(csv_rows(table, worm.dtype))
        self.inner.on_cycle_end(sim)
        self.plt, self.cam, self.mic = [], [], []

    def on_sim_end(self, sim):
        self.inner.on_sim_end(sim)
        self.log.close()

    def begin_movement_prediction(self, sim):
        return self.inner.begin_movement_prediction(sim)

    def provide_movement_vector(self, sim):
        return self.inner.provide_movement_vector(sim)

    def _cycle_predict_all(self, sim):
        return self.inner._cycle_predict_all(sim)


def sim_setup():
    """The track goes through the same csv round trip as in make_golden_log.py (pandas' default float parser is not
    round-trip exact, so the controller must read the table exactly as the reference's CsvController did)."""
    import tempfile

    import pandas as pd

    tab = np.load(os.path.join(GOLD, "reference_golden.npz"))["trace_csv_table"][:400]
    path = os.path.join(tempfile.mkdtemp(), "track.csv")
    pd.DataFrame(tab, columns=["wrm_x", "wrm_y", "wrm_w", "wrm_h"]).to_csv(path, index=False)
    tab = path
    exp = ExperimentConfig("g", 400, 60, (1080, 1920), 90, (960, 540))
    t = TimingConfig(exp, 100, 40, 50, (4.0, 4.0), (0.32, 0.32))
    return tab, exp, t


@pytest.mark.parametrize("tag,cls", [("f64", CsvController), ("f32", F32Controller)])
def test_oracle_rows_reproduce_the_reference_csv(tmp_path, tag, cls):
    tab, exp, t = sim_setup()
    out = str(tmp_path / "bboxes.csv")
    Simulator(t, exp, OracleLogger(cls(t, tab), out)).run()
    want = open(os.path.join(GOLD, f"reference_bboxes_{tag}.csv"), newline="").read()
    got = open(out, newline="").read()
    assert got == want


def test_log_config_paths_and_logger(tmp_path):
    cfg = LogConfig(str(tmp_path / "run"))
    assert cfg.bbox_file_path.endswith("run/bboxes.csv") and cfg.err_file_path.endswith("run/errors/cam_{:09d}.png")
    assert cfg.wrm_file_path.format(12).endswith("worms/wrm_000000012.png")
    cfg.create_dirs()
    assert os.path.isdir(tmp_path / "run" / "micro") and os.path.isdir(tmp_path / "run" / "worms")
    with CSVLogger(str(tmp_path / "x.csv"), ["a", "b"]) as lg:
        lg.write({"b": np.float32(1.5), "a": 3})
        lg.write((4, "moving"))
        lg.writerows([(5, 6), (7, 8)])
    assert open(tmp_path / "x.csv", newline="").read() == "a,b\r\n3,1.5\r\n4,moving\r\n5,6\r\n7,8\r\n"


def test_discretize_part_matches_boxutils_golden():
    g = np.load(os.path.join(GOLD, "reference_golden.npz"))
    b = g["disc_in"].copy()
    n = len(b)
    zero4, zero2 = np.zeros((n, 4), np.int64), np.zeros((n, 2), np.int64)
    table, crop, legal = log_ref.log_rows(b, zero4, zero4, zero2, 0, 9, 6, (1080, 1920))
    assert np.array_equal(crop, g["disc_out"]) and np.array_equal(legal, g["disc_legal"])
    assert np.array_equal(table[:, 13:17], np.nan_to_num(np.where(np.isfinite(b).all(1)[:, None], b, 0.0)))
