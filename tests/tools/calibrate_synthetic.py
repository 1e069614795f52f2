"""Builds the head constants for the seeded synthetic YOLOv8s weights and writes
models/synthetic_yolov8s_calib.json.

The trained ``models/yolov8s_trained.pt`` is not in the reference checkout, and a purely random
detector has a head that barely varies across anchors (bf16 rounding then dominates every
comparison, and nothing it "detects" relates to the worm).  So the two FINAL 1x1 convolutions of
the stride-8 level are fitted in closed form (ridge regression on the random backbone's features
over a few dozen synthetic worm views) to fire on the worm head and regress its 14x14 box:
  * class logit : class-balanced ridge fit of {head cell -> 1, far background -> 0} (the rest of the
                  worm is "do not care"), affinely mapped so head cells sit near conf 0.27 and the
                  highest background cell near conf 0.05
  * box logits  : per side a ridge fit of the DFL distance d (in stride units); the 16 bin logits
                  are -k*(i - d)^2 up to a per-side constant, i.e. LINEAR in d: 2*k*i*d - k*i^2
The stride-16/32 levels get zero class weights and a very negative bias (they never fire).
Everything is deterministic in the seed; uses the fp32 oracle model on the CPU.

    PYTHONPATH=. python tests/tools/calibrate_synthetic.py [seed ...]
"""
import json
import os
import math
import sys

import numpy as np
import torch

from oracle import yolov8_ref as O
from wtracker_b200 import synth
from wtracker_b200.detector.weights import CALIB_PATH, synthetic_state_dict



# a modest logit range keeps the bf16 feature noise (~0.6 % of the feature std after ~25 bf16-stored layers)
# below the 1e-2 confidence tolerance: head cells -> conf 0.21, the highest background cell -> conf 0.063
POS_LOGIT, NEG_LOGIT = -1.3, -2.7
DONT_CARE = int(os.environ.get("WT_DONT_CARE", "5"))   # cells around the head excluded from the fit
N_VIEWS, VIEW, RIDGE, KAPPA = int(os.environ.get('WT_NVIEWS', '80')), 640, float(os.environ.get('WT_RIDGE', '1.0')), 2.0


def calibrate(seed: int) -> dict:
    sd = synthetic_state_dict(seed, calibrated=False)
    model = O.build_model(sd)
    det = model.model[22]
    rng = np.random.default_rng(5000 + seed)
    track = synth.worm_track(4000, seed)
    feats_c, feats_b, heads = [], [], []
    for i in range(0, N_VIEWS, 4):
        views, hc = [], []
        for j in range(4):
            fi = int(rng.integers(0, 4000))
            while not (150 < track[fi, 0] < synth.FRAME_W - 150 and 150 < track[fi, 1] < synth.FRAME_H - 150):
                fi = int(rng.integers(0, 4000))     # keep the worm off the replicated frame border
            frame = synth.render_frame(fi, track, seed)
            off = rng.integers(-250, 251, 2)
            pos = (int(np.clip(track[fi, 0] + off[0], VIEW // 2, synth.FRAME_W - VIEW // 2)),
                   int(np.clip(track[fi, 1] + off[1], VIEW // 2, synth.FRAME_H - VIEW // 2)))   # view inside the frame
            views.append(np.ascontiguousarray(synth.camera_view(frame, pos, VIEW)))
            hc.append((track[fi, 0] - (pos[0] - VIEW // 2), track[fi, 1] - (pos[1] - VIEW // 2)))
        taps = {}
        with torch.no_grad():
            model.features(O.preprocess(views, VIEW), taps)
            fb = det.cv2[0][1](det.cv2[0][0](taps["x15"]))      # (4, 64, 80, 80)
            fc = det.cv3[0][1](det.cv3[0][0](taps["x15"]))      # (4, 128, 80, 80)
        feats_b.append(fb.permute(0, 2, 3, 1).double())
        feats_c.append(fc.permute(0, 2, 3, 1).double())
        heads += hc
    fb, fc = torch.cat(feats_b), torch.cat(feats_c)             # (N, 80, 80, C)
    n, gh, gw, _ = fc.shape

    # ---- class logit ------------------------------------------------------------------------
    # head cell -> 1, everything else (background, body, tail) -> 0; the 5x5 cells around the head
    # are "don't care" so the fit is not asked for a one-cell-sharp response
    target = torch.zeros(n, gh, gw, dtype=torch.float64)
    care = torch.ones(n, gh, gw, dtype=torch.bool)
    pos_cells = []
    for k, (hx, hy) in enumerate(heads):
        cx, cy = int(hx // 8), int(hy // 8)
        care[k, max(cy - DONT_CARE, 0): cy + DONT_CARE + 1, max(cx - DONT_CARE, 0): cx + DONT_CARE + 1] = False
        if 1 <= cx < gw - 1 and 1 <= cy < gh - 1:
            target[k, cy, cx] = 1.0
            care[k, cy, cx] = True
            pos_cells.append((k, cy, cx, hx, hy))
    sel = care.reshape(-1)
    X = fc.reshape(-1, fc.shape[-1])[sel]
    t = target.reshape(-1)[sel]
    wgt = torch.where(t > 0, torch.tensor(0.5 / (t > 0).sum()), torch.tensor(0.5 / (t == 0).sum()))   # class-balanced
    mu = (X * wgt[:, None]).sum(0)
    Xc = X - mu
    A = (Xc * wgt[:, None]).T @ Xc + RIDGE * torch.eye(Xc.shape[1], dtype=torch.float64) * (Xc.var(0).mean())
    w = torch.linalg.solve(A, (Xc * wgt[:, None]).T @ (t - 0.5))
    score = Xc @ w
    s_pos = float(score[t > 0].mean())
    s_neg = float(score[t == 0].max())
    # head cells -> logit -1 on average (conf 0.27), the highest background cell -> logit -3 (conf 0.047);
    # a modest logit range keeps |w| (and with it the amplification of bf16 feature noise) small
    if os.environ.get("WT_CALIB_DEBUG"):
        full = torch.full((n * gh * gw,), -1e9, dtype=torch.float64)
        full[sel] = torch.where(t == 0, score, torch.tensor(-1e9, dtype=torch.float64))
        for v, i in zip(*torch.topk(full, 6)):
            k, r = divmod(int(i), gh * gw)
            print(f"  neg score {float(v):.3f} view {k} cell (x={r % gw}, y={r // gw}) head cell (x={int(heads[k][0] // 8)}, y={int(heads[k][1] // 8)})")
    alpha = (POS_LOGIT - NEG_LOGIT) / (s_pos - s_neg)
    beta = POS_LOGIT - alpha * s_pos
    w_cls = alpha * w
    b_cls = beta - float(w_cls @ mu)
    out = {"model.22.cv3.0.2": {"weight": [w_cls.float().tolist()], "bias": [b_cls]}}
    for lvl, c in ((1, 128), (2, 128)):
        out[f"model.22.cv3.{lvl}.2"] = {"weight": [[0.0] * c], "bias": [-12.0]}
        out[f"model.22.cv2.{lvl}.2"] = {"weight": [[0.0] * 64 for _ in range(64)], "bias": [1.0] * 64}

    # ---- box logits (stride 8) -----------------------------------------------------------------
    rows, dist = [], []
    for k, cy, cx, hx, hy in pos_cells:
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ax, ay = cx + dx + 0.5, cy + dy + 0.5
                x1, y1, x2, y2 = (hx - 7) / 8, (hy - 7) / 8, (hx + 7) / 8, (hy + 7) / 8
                d = (ax - x1, ay - y1, x2 - ax, y2 - ay)
                if min(d) > 0.05:
                    rows.append(fb[k, cy + dy, cx + dx])
                    dist.append(d)
    Xb = torch.stack(rows)
    D = torch.tensor(dist, dtype=torch.float64)
    mub, mud = Xb.mean(0), D.mean(0)
    Xbc = Xb - mub
    Ab = Xbc.T @ Xbc / Xbc.shape[0] + 1e-2 * torch.eye(Xbc.shape[1], dtype=torch.float64) * (Xbc.var(0).mean())
    Wd = torch.linalg.solve(Ab, Xbc.T @ (D - mud) / Xbc.shape[0])      # (64, 4): d_side = (f - mub) @ Wd + mud
    bd = mud - mub @ Wd
    W = torch.zeros(64, 64, dtype=torch.float64)
    b = torch.zeros(64, dtype=torch.float64)
    for side in range(4):
        for i in range(16):
            W[side * 16 + i] = 2 * KAPPA * i * Wd[:, side]
            b[side * 16 + i] = 2 * KAPPA * i * bd[side] - KAPPA * i * i
    out["model.22.cv2.0.2"] = {"weight": W.float().tolist(), "bias": b.float().tolist()}
    fit = (Xbc @ Wd + mud - D).abs().mean().item()
    pos_logit = (Xc @ w_cls + beta)[t > 0]
    neg_logit = (Xc @ w_cls + beta)[t == 0]
    print(f"ridge {RIDGE}: pos logit min {pos_logit.min():.2f} mean {pos_logit.mean():.2f}; neg logit max {neg_logit.max():.2f} "
          f"frac neg > logit(0.1) {(neg_logit > math.log(0.1 / 0.9)).double().mean():.2e}")
    print(f"seed {seed}: {len(pos_cells)} head cells, cls score pos {s_pos:.3f} / neg max {s_neg:.3f}, "
          f"box fit |err| {fit:.3f} strides, |w_cls| {w_cls.norm():.2f}")
    return out


if __name__ == "__main__":
    seeds = [int(a) for a in sys.argv[1:]] or [0]
    data = json.loads(CALIB_PATH.read_text()) if CALIB_PATH.exists() else {}
    for seed in seeds:
        data[f"s-nc1-seed{seed}"] = calibrate(seed)
    CALIB_PATH.parent.mkdir(exist_ok=True)
    CALIB_PATH.write_text(json.dumps(data))
    print("wrote", CALIB_PATH)
