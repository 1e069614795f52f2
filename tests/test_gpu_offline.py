"""Offline detection over a frame range (BASELINE configs[3]): the 32-byte rows written on the device must hold what
``YoloController.predict`` returns for the same views, in frame order, and ranges must tile the job."""
import numpy as np
import pytest
import torch

from gpu_common import sample_frames, synthetic_sd
from wtracker_b200 import synth

pytestmark = pytest.mark.gpu


def test_rows_equal_the_plugin_prediction():
    from wtracker_b200.detector.engine import DetectorEngine
    from wtracker_b200.offline import decode_rows, detect_range, run_offline

    frames, tr = sample_frames(6)
    d_frames = torch.from_numpy(frames).cuda()
    total = 23

    def schedule(first, n):
        f = np.arange(first, first + n)
        idx = (f % 6).astype(np.int32)
        cx = (tr[idx, 0].astype(np.int64) + (f * 37) % 61 - 30 - 180).astype(np.int32)
        cy = (tr[idx, 1].astype(np.int64) - (f * 53) % 47 + 20 - 180).astype(np.int32)
        return idx, cx, cy

    eng = DetectorEngine(synthetic_sd(), (360, 360), 384, batch=4, max_det=1)
    full, detect_ms, gather_ms = run_offline(eng, d_frames, schedule, total, 0, 1)
    rows = decode_rows(full)
    assert np.array_equal(rows["frame"], np.arange(total)) and detect_ms > 0
    # the same ranges computed piecewise (as ranks would) give the same rows
    from wtracker_b200.sharding import frame_range

    parts = [detect_range(eng, d_frames, schedule, *frame_range(total, r, 3)) for r in range(3)]
    assert torch.equal(torch.cat(parts), full)
    # against the host plugin path on numpy-cropped views
    idx, cx, cy = schedule(0, total)
    views = [np.ascontiguousarray(synth.camera_view(frames[i], (int(x) + 180, int(y) + 180), 360)) for i, x, y in zip(idx, cx, cy)]
    boxes, counts = DetectorEngine(synthetic_sd(), (360, 360), 384, batch=8, max_det=1).detect_views(views)
    assert np.array_equal(rows["valid"], counts > 0)
    v = rows["valid"]
    assert v.any()
    b = boxes[v, 0]
    want = np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)
    assert np.array_equal(rows["xywh"][v], want)
    assert np.array_equal(rows["conf"][v], b[:, 4]) and np.array_equal(rows["anchor"][v], b[:, 5].astype(np.int32))
    assert np.isnan(rows["xywh"][~v]).all() and (rows["anchor"][~v] == -1).all()
