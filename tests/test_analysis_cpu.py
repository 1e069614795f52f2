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
