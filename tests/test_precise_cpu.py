"""calculate_precise on the CPU: the oracle restatement against the outputs of the
UNMODIFIED reference (tests/golden/reference_precise.npz, make_golden_precise.py), indexing quirk included."""
import os

import numpy as np
import pytest

from oracle import precise_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_precise.npz")


@pytest.mark.parametrize("thr", [10, 20.5])
def test_oracle_matches_reference(thr):
    g = np.load(GOLD)
    err, legal = precise_ref.precise_error(g["frames"], g["frame_nums"], g["background"], g["worm"], g["mic"], thr)
    assert np.array_equal(legal, g["legal"])
    want = g[f"err_thr{thr}"]
    got = precise_ref.compact(err, legal)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).sum() == 1     # only the row behind n_legal
    assert np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    assert 0.0 < np.nanmean(want) < 1.0 and (want == 1.0).any() and (want == 0.0).any()
