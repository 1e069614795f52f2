"""DataAnalyzer.initialize on the GPU: the drop-in class reproduces the UNMODIFIED reference's DataFrame (values and
column order), and the kernel equals the oracle on a large log with missing detections."""
import os

import numpy as np
import pytest
import torch

from oracle import analysis_ref
from test_analysis_cpu import GOLD, log_table17
from wtracker_b200.eval.data_analyzer import ANALYSIS_COLUMNS, DataAnalyzer, analysis_table_device
from wtracker_b200.sim import ExperimentConfig, TimingConfig

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("period", [10, 3])
def test_drop_in_initialize_matches_reference(period):
    g = np.load(os.path.join(GOLD, "reference_analysis.npz"))
    exp = ExperimentConfig("g", 400, 60, (1080, 1920), 90, (960, 540))
    t = TimingConfig(exp, 100, 40, 50, (4.0, 4.0), (0.32, 0.32))
    an = DataAnalyzer.load(t, os.path.join(GOLD, "reference_bboxes_f64.csv"))
    an.initialize(period=period)
    assert list(an.data.columns) == list(g[f"columns_p{period}"])
    got = an.data.drop(columns=["phase"]).to_numpy(dtype=np.float64)
    assert np.array_equal(got, g[f"values_p{period}"], equal_nan=True)
    assert an.data["frame"].dtype == np.int64 and an.data["phase"].iloc[0] == "imaging"


@pytest.mark.parametrize("n", [1, 11, 100_000])
def test_kernel_equals_oracle(n):
    rng = np.random.default_rng(n)
    base = log_table17()
    t = base[rng.integers(0, len(base), n)].copy()
    t[:, 0] = np.arange(n) * rng.integers(1, 3)            # frame numbers (possibly with gaps)
    t[:, 13:17] += rng.normal(0, 3, (n, 4))
    t[rng.uniform(size=n) < 0.03, 13:17] = 0.0             # missed detections as the log stores them
    t[rng.uniform(size=n) < 0.01, 13:17] = np.nan
    want = analysis_ref.analysis_columns(t, 10, 9)
    got = analysis_table_device(torch.from_numpy(t).cuda(), 10, 9).cpu().numpy()
    assert got.shape == (n, len(ANALYSIS_COLUMNS))
    assert np.array_equal(got, want, equal_nan=True)
