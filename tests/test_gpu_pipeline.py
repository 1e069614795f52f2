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


def test_fused_tail_equals_the_four_separate_kernels(hot_path):
    """wt_hot_tail (one launch) against wt_track_rows + wt_mlp_gather + wt_resmlp_forward + wt_bbox_error: rows,
    network inputs, predictions, validity and errors must be bit-identical, with history both inside and before the batch."""
    hp = hot_path
    views = np.stack(views_for(360, 4, seed=0))
    hist = np.stack([np.linspace(100, 140, 40), np.linspace(80, 70, 40), np.full(40, 14.0), np.full(40, 13.0)], 1)
    outs = []
    for fused in (True, False):
        hp.fused_tail = fused
        hp.table.fill_(float("nan"))
        hp.table[:40] = torch.from_numpy(hist).to(hp.table.device)
        hp.table[17] = float("nan")                       # a gap in the history of some rows
        res = []
        for first in (40, 44, 48):                        # later batches read rows the earlier ones wrote
            r = hp.step_host(views, first_row=first)
            res.append({k: a.copy() for k, a in r.items()})
            res[-1]["x"] = hp.mlp_x.cpu().numpy().copy()
            res[-1]["mic"] = hp.mic_table[first: first + 4].cpu().numpy().copy()
        outs.append(res)
    hp.fused_tail = True
    assert any(r["pred_valid"].any() for r in outs[0]) and not all(r["pred_valid"].all() for r in outs[0])
    for a, b in zip(*outs):
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_run_host_with_distinct_unpinned_batches(hot_path):
    """Plain numpy batches go through the slots' own pinned staging buffers; a refill must wait for the H2D that last
    read the buffer (many distinct batches back to back, compared with the one-at-a-time path)."""
    hp = hot_path
    base = np.stack(views_for(360, 4, seed=0))
    batches = [np.roll(base, shift=(7 * i, -5 * i), axis=(1, 2)).copy() for i in range(14)]
    want = []
    for v in batches:
        hp.table.fill_(float("nan"))
        want.append({k: a.copy() for k, a in hp.step_host(v, first_row=0).items()})
    hp.table.fill_(float("nan"))
    got = [{k: a.copy() for k, a in r.items()} for r in hp.run_host(iter(batches), first_row=0)]
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        for k in ("boxes", "count", "bbox_error"):
            assert np.array_equal(g[k], w[k], equal_nan=True), (i, k)
    assert len({r["boxes"].tobytes() for r in want}) > 7, "the batches must differ for this test to mean anything"
    # abandoning the generator half way leaves the object usable
    it = hp.run_host(iter(batches))
    next(it)
    it.close()
    r = hp.step_host(batches[3])
    assert np.array_equal(r["boxes"], want[3]["boxes"])


def test_run_frames_takes_the_crops_on_the_device(hot_path):
    """Frame ingest: whole frames from the host + crop origins == the same views cropped on the host."""
    from gpu_common import sample_frames
    from wtracker_b200 import synth

    hp = hot_path
    frames, tr = sample_frames(6)
    batches_f, batches_v = [], []
    for b in range(3):
        idx = [(b + i) % 6 for i in range(4)]
        cx = np.array([int(tr[j, 0]) - 180 + 9 * i - 4 * b for i, j in enumerate(idx)], dtype=np.int32)
        cy = np.array([int(tr[j, 1]) - 180 - 6 * i + 3 * b for i, j in enumerate(idx)], dtype=np.int32)
        if b == 2:
            cx[0], cy[1] = -50, 900                       # views hanging over the frame border (replicate)
        batches_f.append((frames[idx], cx, cy))
        batches_v.append(np.stack([synth.camera_view(frames[j], (int(x) + 180, int(y) + 180), 360) for j, x, y in zip(idx, cx, cy)]))
    hp.table.fill_(float("nan"))
    want = [{k: a.copy() for k, a in r.items()} for r in hp.run_host(iter(batches_v))]
    hp.table.fill_(float("nan"))
    got = [{k: a.copy() for k, a in r.items()} for r in hp.run_frames(iter(batches_f))]
    for (f, cx, cy), g, w in zip(batches_f, got, want):
        assert np.array_equal(g["boxes"], w["boxes"]) and np.array_equal(g["count"], w["count"])
        shift = np.stack([cx, cy, 0 * cx, 0 * cx], 1).astype(np.float64)
        assert np.array_equal(g["worm"], w["worm"] + shift, equal_nan=True)
