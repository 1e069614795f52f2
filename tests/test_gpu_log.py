"""bboxes.csv wire format on the GPU: the drop-in LoggingController (wt_log_rows underneath) writes the same bytes as
the UNMODIFIED reference (tests/golden/reference_bboxes_*.csv), and the kernel equals the oracle on hostile tables."""
import os

import numpy as np
import pytest
import torch

from oracle import log_ref
from test_log_cpu import F32Controller, GOLD, sim_setup
from wtracker_b200.sim import Simulator
from wtracker_b200.sim.sim_controllers import CsvController, LogConfig, LoggingController
from wtracker_b200.sim.sim_controllers.logging_controller import log_table_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,cls", [("f64", CsvController), ("f32", F32Controller)])
def test_logging_controller_writes_the_reference_csv(tmp_path, tag, cls):
    tab, exp, t = sim_setup()
    cfg = LogConfig(str(tmp_path / tag), save_err_view=False)
    Simulator(t, exp, LoggingController(cls(t, tab), cfg)).run()
    want = open(os.path.join(GOLD, f"reference_bboxes_{tag}.csv"), newline="").read()
    assert open(cfg.bbox_file_path, newline="").read() == want


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 9, 4097, 1 << 20])
def test_kernel_equals_oracle(dtype, n):
    rng = np.random.default_rng(n)
    worm = np.stack([rng.uniform(-400, 400, n), rng.uniform(-400, 400, n), rng.uniform(-3, 40, n), rng.uniform(-3, 40, n)], 1)
    worm[rng.uniform(size=n) < 0.05] = np.nan
    worm[rng.uniform(size=n) < 0.02, 2] = np.inf
    m = rng.uniform(size=n) < 0.05
    worm[m] = np.round(worm[m])                     # boxes on exact integer coordinates (floor == ceil)
    worm = worm.astype(dtype)
    cam = np.stack([rng.integers(-200, 1900, n), rng.integers(-200, 1000, n), np.full(n, 360), np.full(n, 360)], 1)
    mic = np.stack([cam[:, 0] + 166, cam[:, 1] + 166, np.full(n, 29), np.full(n, 29)], 1)
    plt = cam[:, :2] + 180
    want_t, want_c, want_l = log_ref.log_rows(worm, cam, mic, plt, 123, 9, 6, (1080, 1920))
    t, c, l = log_table_device(torch.from_numpy(worm).cuda(), torch.from_numpy(cam), torch.from_numpy(mic),
                               torch.from_numpy(plt), 123, 9, 6, (1080, 1920))
    assert np.array_equal(t.cpu().numpy(), want_t)
    assert np.array_equal(c.cpu().numpy(), want_c)
    assert np.array_equal(l.cpu().numpy().astype(bool), want_l)
