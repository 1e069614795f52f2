"""DataAnalyzer.initialize on the CPU: the oracle restatement against the DataFrame the
UNMODIFIED reference produced from the golden log (tests/golden/reference_analysis.npz, make_golden_analysis.py)."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import analysis_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def log_table17():
    df = pd.read_csv(os.path.join(GOLD, "reference_bboxes_f64.csv"))
    t = df.drop(columns=["phase"]).to_numpy(dtype=np.float64)
    return np.insert(t, 2, (df["phase"] == "moving").to_numpy(dtype=np.float64), axis=1)


@pytest.mark.parametrize("period", [10, 3])
def test_oracle_matches_reference(period):
    g = np.load(os.path.join(GOLD, "reference_analysis.npz"))
    assert list(g[f"names_p{period}"]) == analysis_ref.NAMES
    got = analysis_ref.analysis_columns(log_table17(), period, 9)
    assert np.array_equal(got, g[f"values_p{period}"], equal_nan=True)
    assert np.isnan(got[:period, 22:25]).all() and np.isfinite(got[period:, 22:25]).all()


CLEAN_CASES = {
    "trim": dict(trim_cycles=True),
    "imaging": dict(imaging_only=True),
    "bounds": dict(bounds=(330.0, 200.0, 640.0, 470.0)),
    "all": dict(trim_cycles=True, imaging_only=True, bounds=(320.5, 190.25, 660.0, 480.0)),
}
ANOMALY_CASES = {
    "a": dict(no_preds=True, min_bbox_error=0.95, min_dist_error=40.0, min_speed=4.0, min_size=14.9),
    "b": dict(no_preds=False, min_bbox_error=0.9),
    "c": dict(no_preds=True, min_speed=2.5),
}


def analysed_table_with_gaps():
    """The reference's analysed table (period 10) with the NaN rows make_golden_analysis.py injects before cleaning."""
    g = np.load(os.path.join(GOLD, "reference_analysis.npz"))
    m = np.load(os.path.join(GOLD, "reference_masks.npz"))
    t = g["values_p10"].copy()
    t[m["nan_rows"], 12:16] = np.nan
    moving = (pd.read_csv(os.path.join(GOLD, "reference_bboxes_f64.csv"))["phase"] == "moving").to_numpy()
    return t, moving, m


def test_clean_and_anomaly_oracle_matches_reference():
    """DataAnalyzer.clean / calc_anomalies of the UNMODIFIED reference (tests/golden/reference_masks.npz)."""
    t, moving, m = analysed_table_with_gaps()
    frames = t[:, 0].astype(np.int64)
    for name, kw in CLEAN_CASES.items():
        keep = analysis_ref.clean_keep(t, moving, **kw)
        assert np.array_equal(frames[keep], m[f"clean_{name}"]), name
        assert 0 < keep.sum() < len(t)
    for name, kw in ANOMALY_CASES.items():
        sel = ~moving if name == "b" else np.ones(len(t), dtype=bool)
        bits = analysis_ref.anomaly_bits(t[sel], **kw)
        hit = bits != 0
        assert np.array_equal(frames[sel][hit], m[f"anom_{name}_frames"]), name
        flags = np.stack([(bits[hit] >> i) & 1 for i in range(6)], 1).astype(bool)
        assert np.array_equal(flags, m[f"anom_{name}_flags"]), name
