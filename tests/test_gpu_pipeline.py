"""The whole hot path object (HotPath): the pipelined host API must give exactly what the one-batch-at-a-time
host API gives, batch by batch, and both must agree with the oracle on the kept anchor."""
import numpy as np
import pytest
import torch

from gpu_common import oracle_model, synthetic_sd, views_for
from oracle import yolov8_ref as Y

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hot_path():
    from wtracker_b200.neural.mlp import load_worm_predictor
    from wtracker_b200.paths import RESMLP_100
    from wtracker_b200.pipeline import HotPath

    return HotPath(synthetic_sd(), load_worm_predictor(RESMLP_100), view=360, imgsz=384, batch=4, micro=29,
                   table_rows=64)


def _batches(n_batches):
    out = []
    for b in range(n_batches):
        out.append(np.stack(views_for(360, 4, seed=0))[::-1 if b % 2 else 1].copy())
    out.append(out[0][:3].copy())          # ragged last batch
    return out


def test_run_host_equals_step_host(hot_path):
    hp = hot_path
    batches = _batches(5)
    want = []
    for i, v in enumerate(batches):
        hp.table.fill_(float("nan"))
        r = hp.step_host(v, first_row=0)
        want.append({k: a.copy() for k, a in r.items()})
    hp.table.fill_(float("nan"))
    got = []
    for r in hp.run_host(iter(batches), first_row=0):
        got.append({k: a.copy() for k, a in r.items()})
    torch.cuda.synchronize()
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ("boxes", "count", "worm", "bbox_error"):
            assert np.array_equal(g[k], w[k], equal_nan=True), k
    # pinned tensors go through without the staging copy and give the same rows
    pinned = [torch.from_numpy(v).pin_memory() for v in batches]
    hp.table.fill_(float("nan"))
    again = [{k: a.copy() for k, a in r.items()} for r in hp.run_host(iter(pinned), first_row=0)]
    for g, w in zip(again, want):
        assert np.array_equal(g["boxes"], w["boxes"]) and np.array_equal(g["count"], w["count"])


def test_run_host_kept_anchor_matches_oracle(hot_path):
    batches = _batches(1)[:1]
    res = list(hot_path.run_host(iter(batches)))[0]
    ref = Y.YoloOracle(oracle_model(), 384, max_det=1).detect(list(batches[0]))
    for i, (rows, idx) in enumerate(ref):
        assert res["count"][i] == rows.shape[0]
        if rows.shape[0]:
            assert int(res["boxes"][i, 0, 5]) == int(idx[0])
            assert np.abs(res["boxes"][i, 0, :4] - rows[0, :4].numpy()).max() < 0.5


def test_run_host_empty_iterator(hot_path):
    assert list(hot_path.run_host(iter([]))) == []
