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


def _analyzer():
    from test_analysis_cpu import analysed_table_with_gaps

    exp = ExperimentConfig("g", 400, 60, (1080, 1920), 90, (960, 540))
    t = TimingConfig(exp, 100, 40, 50, (4.0, 4.0), (0.32, 0.32))
    an = DataAnalyzer.load(t, os.path.join(GOLD, "reference_bboxes_f64.csv"))
    an.initialize(period=10)
    _, _, m = analysed_table_with_gaps()
    an.data.loc[m["nan_rows"], ["wrm_x", "wrm_y", "wrm_w", "wrm_h"]] = np.nan
    return an, m


def test_drop_in_clean_and_anomalies_match_reference():
    """The drop-in DataAnalyzer.clean / calc_anomalies select exactly the rows (and set exactly the flags) the unmodified
    reference does (tests/golden/reference_masks.npz)."""
    from test_analysis_cpu import ANOMALY_CASES, CLEAN_CASES
    from wtracker_b200.eval.data_analyzer import ANOMALY_FLAGS

    for name, kw in CLEAN_CASES.items():
        an, m = _analyzer()
        an.clean(**kw)
        assert np.array_equal(an.data["frame"].to_numpy(), m[f"clean_{name}"]), name
    for name, kw in ANOMALY_CASES.items():
        an, m = _analyzer()
        an.clean(imaging_only=(name == "b"))
        res = an.calc_anomalies(**kw, remove_anomalies=(name == "c"))
        assert np.array_equal(res["frame"].to_numpy(), m[f"anom_{name}_frames"]), name
        assert np.array_equal(res[ANOMALY_FLAGS].to_numpy(dtype=bool), m[f"anom_{name}_flags"]), name
        assert np.array_equal(an.data["frame"].to_numpy(), m[f"anom_{name}_left"]), name
    an, _ = _analyzer()
    an.remove_cycle([0, 3])
    assert not an.data["cycle"].isin([0, 3]).any()
    an.reset_changes()
    assert len(an.data) == 396


def test_mask_kernel_equals_oracle_on_a_large_table():
    from wtracker_b200 import _lib as L
    import ctypes as C

    rng = np.random.default_rng(5)
    n = 300_000
    g = np.load(os.path.join(GOLD, "reference_analysis.npz"))
    t = g["values_p10"][rng.integers(0, 396, n)].copy()
    t[:, 12:16] += rng.normal(0, 20, (n, 4))
    t[rng.uniform(size=n) < 0.05, 12 + rng.integers(0, 4)] = np.nan
    t[rng.uniform(size=n) < 0.02, 24] = np.nan
    moving = rng.uniform(size=n) < 0.3
    bounds = (330.0, 200.0, 640.0, 470.0)
    kw = dict(no_preds=True, min_bbox_error=0.7, min_dist_error=30.0, min_speed=3.0, min_size=20.0)
    d_t, d_m = torch.from_numpy(t).cuda(), torch.from_numpy(moving.astype(np.uint8)).cuda()
    keep, bits = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    L.check(L.lib().wt_analysis_masks(d_t.data_ptr(), d_m.data_ptr(), n, 1, (C.c_double * 4)(*bounds), 1, 0.7, 30.0, 3.0, 20.0,
                                      keep.data_ptr(), bits.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(keep.cpu().numpy().astype(bool), analysis_ref.clean_keep(t, moving, imaging_only=True, bounds=bounds))
    assert np.array_equal(bits.cpu().numpy(), analysis_ref.anomaly_bits(t, **kw))
