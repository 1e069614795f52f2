"""Detector parity AT THE BENCHMARKED CONFIGURATION (BASELINE configs[1]: batch 64, 640 x 640, bf16): 256 distinct
synthetic camera views through a batch-64 engine against the fp32 oracle — every tapped feature map, the kept anchor,
box and confidence of every image — with the tie margin accounted for explicitly.

The kept index is the arg-max of the confidence (max_det = 1).  Confidences carry bf16 feature noise (stated tolerance
1e-2), so two anchors whose fp32 confidences are closer than that can swap.  Every image therefore falls into one of
three classes, all counted and printed: IDENTICAL index; SWAPPED inside the margin (the oracle scores the GPU's anchor
within 2e-2 of its own best); WRONG (anything else — the test fails).  The swapped fraction must stay below 10 %.
"""
import json
import os

import numpy as np
import pytest
import torch

from gpu_common import box_iou, sample_frames
from oracle import yolov8_ref as Y
from wtracker_b200 import synth

pytestmark = pytest.mark.gpu

BATCH, IMGSZ, N_VIEWS = 64, 640, 256
CONF_TOL, BOX_TOL, MARGIN = 1e-2, 0.5, 2e-2


def make_views():
    """256 distinct 640 x 640 views: 4 seeds x 8 frames x 8 crop offsets, among them crops that hang over the frame
    border (replicate padding) and crops with the worm near the view's edge."""
    views = []
    for seed in range(4):
        frames, tr = sample_frames(8, seed)
        for i in range(8):
            for j in range(8):
                dx, dy = (37 * j - 120) + 11 * i, (53 * j) % 160 - 80 - 7 * i
                pos = (int(tr[i, 0]) + dx, int(tr[i, 1]) + dy)
                if j == 7:
                    pos = (40 + 13 * i, int(tr[i, 1]))            # view hangs over the left frame border
                if j == 6 and i % 2:
                    pos = (int(tr[i, 0]), 1080 - 60 - 9 * i)       # ... over the bottom border
                views.append(np.ascontiguousarray(synth.camera_view(frames[i], pos, IMGSZ)))
    assert len(views) == N_VIEWS and len({v.tobytes() for v in views}) == N_VIEWS
    return views


def run_case(calibrated: bool, tag: str):
    from wtracker_b200.detector.engine import DetectorEngine
    from wtracker_b200.detector.weights import synthetic_state_dict

    sd = synthetic_state_dict(0, calibrated=calibrated, head_gain=1.0 if calibrated else 5.0)
    model = Y.build_model(sd)
    eng = DetectorEngine(sd, (IMGSZ, IMGSZ), IMGSZ, batch=BATCH, max_det=1)
    views = make_views()
    stats = dict(identical=0, swapped=0, wrong=0, none_both=0, count_mismatch=0, max_conf_err=0.0, max_box_err=0.0,
                 worst_margin=0.0, feature_err={})
    failures = []
    for b in range(N_VIEWS // BATCH):
        chunk = views[b * BATCH: (b + 1) * BATCH]
        boxes, counts = eng.detect_views(chunk)
        taps = {}
        with torch.no_grad():
            feats = model.features(Y.preprocess(chunk, IMGSZ), taps)
            pred = Y.decode_head(feats)
        # ---- every tapped map of this batch against the fp32 oracle
        for name, (bid, coff, c) in eng.program.taps.items():
            got = eng.buffer_tensor(bid, BATCH)[..., coff:coff + c].float().permute(0, 3, 1, 2).cpu()
            ref = taps[name]
            err = (got - ref).abs()
            rel_mean, rel_max = float(err.mean() / ref.std()), float(err.max() / ref.abs().max())
            prev = stats["feature_err"].get(name, (0.0, 0.0))
            stats["feature_err"][name] = (max(prev[0], rel_mean), max(prev[1], rel_max))
            if not (rel_mean < 0.012 and rel_max < 0.05):
                failures.append(f"batch {b} map {name}: mean err / std {rel_mean:.4f}, max err / max {rel_max:.4f}")
        # ---- detections
        res = Y.non_max_suppression(pred, 0.1, 0.7, 1)
        conf_all = pred[:, 4]
        for i, (rows, idx) in enumerate(res):
            img = b * BATCH + i
            top = float(conf_all[i].max())
            if rows.shape[0] == 0 or counts[i] == 0:
                if rows.shape[0] == 0 and counts[i] == 0:
                    stats["none_both"] += 1
                elif abs(top - 0.1) < CONF_TOL:
                    stats["count_mismatch"] += 1          # the best anchor sits on the confidence threshold
                else:
                    stats["wrong"] += 1
                    failures.append(f"image {img}: count {counts[i]} vs oracle {rows.shape[0]} (top conf {top:.4f})")
                continue
            got = boxes[i, 0]
            g_idx, o_idx = int(got[5]), int(idx[0])
            want = Y.scale_boxes((IMGSZ, IMGSZ), rows[:, :4], (IMGSZ, IMGSZ))[0].numpy()
            if g_idx == o_idx:
                stats["identical"] += 1
                ce = abs(float(got[4]) - float(rows[0, 4]))
                be = float(np.abs(got[:4] - want).max())
                stats["max_conf_err"] = max(stats["max_conf_err"], ce)
                stats["max_box_err"] = max(stats["max_box_err"], be)
                if ce >= CONF_TOL:
                    failures.append(f"image {img}: confidence off by {ce:.4f}")
                if not (be < BOX_TOL or box_iou(got[:4], want) >= 0.99):
                    failures.append(f"image {img}: box off by {be:.3f} px")
            else:
                margin = float(conf_all[i, o_idx] - conf_all[i, g_idx])
                stats["worst_margin"] = max(stats["worst_margin"], margin)
                if margin < MARGIN:
                    stats["swapped"] += 1
                else:
                    stats["wrong"] += 1
                    failures.append(f"image {img}: anchor {g_idx} vs oracle {o_idx}, oracle margin {margin:.4f}")
    judged = stats["identical"] + stats["swapped"] + stats["wrong"]
    stats["excluded_fraction"] = stats["swapped"] / max(judged, 1)
    stats["judged"] = judged
    line = f"[parity64 {tag}] " + json.dumps(stats)
    print(line)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", f"parity64_{tag}.json"), "w") as f:
        f.write(json.dumps(stats) + "\n" + "\n".join(failures) + "\n")
    return stats, failures


def test_batch64_640_parity_calibrated_weights():
    stats, failures = run_case(True, "calibrated")
    assert not failures, failures[:8]
    assert stats["judged"] >= 0.8 * N_VIEWS, "the worm must be detected wherever it is in view (48 views deliberately miss it)"
    assert stats["none_both"] >= 16, "views without the worm: both sides must agree that there is nothing"
    assert stats["excluded_fraction"] <= 0.10, f"{stats['swapped']} of {stats['judged']} kept indices swapped inside the margin"


def test_batch64_640_parity_uncalibrated_weights():
    """The same with purely random head convolutions (no fitted 1x1 heads; class weights x5 so that a fifth of the anchors
    passes conf 0.1).  Such a net decides by near-ties everywhere (class logits vary by ~0.2 over an image), so the kept
    index may legitimately swap inside the margin in many images — the fraction is printed, not bounded — but the
    tolerances that do not depend on a decision must hold on weights nobody shaped: every feature map, every confidence
    and box where the index agrees, and NO index outside the margin."""
    stats, failures = run_case(False, "uncalibrated")
    assert not failures, failures[:8]
    assert stats["judged"] >= 0.5 * N_VIEWS
    assert stats["wrong"] == 0


def test_oracle_against_real_ultralytics_when_installed():
    """prefer the real library where the box has it.  With ultralytics importable, the restatement must
    reproduce ``YOLO.predict`` (the call of yolo_controller.py:72-78) on the same weights and views; without it the
    detector stays 'parity unpinned' and this test is skipped."""
    ultralytics = pytest.importorskip("ultralytics")
    from ultralytics.nn.tasks import DetectionModel

    from wtracker_b200.detector.weights import synthetic_state_dict

    sd = synthetic_state_dict(0)
    net = DetectionModel("yolov8s.yaml", ch=3, nc=1, verbose=False)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "num_batches_tracked" not in k and "dfl" not in k], missing
    yolo = ultralytics.YOLO("yolov8s.yaml", task="detect", verbose=False)
    yolo.model = net.eval()
    views = make_views()[:8]
    oracle = Y.YoloOracle(Y.build_model(sd), IMGSZ, conf=0.1, iou=0.7, max_det=1)
    ref = oracle.predict(views)
    bgr = [np.repeat(v[:, :, None], 3, axis=2) for v in views]
    out = yolo.predict(bgr, imgsz=IMGSZ, conf=0.1, max_det=1, device="cpu", verbose=False)
    for i, r in enumerate(out):
        b = r.boxes.xyxy.cpu().numpy()
        if b.shape[0] == 0:
            assert np.isnan(ref[i]).all()
        else:
            want = np.array([b[0, 0], b[0, 1], b[0, 2] - b[0, 0], b[0, 3] - b[0, 1]])
            assert np.abs(ref[i] - want).max() < 1e-2, (i, ref[i], want)
