"""Builds the head constants for the seeded synthetic YOLOv8s weights and writes
models/synthetic_yolov8s_calib.json.

A randomly initialised detector has a head whose outputs barely vary across anchors (the final 1x1
convolutions cancel most of the signal), which is unlike any trained model and turns bf16 rounding
into the dominant term.  To make the synthetic model detector-like, the two final 1x1 convolutions
of every level are REPLACED by weights aligned with the principal directions of their input
features over a few synthetic worm frames (a closed-form stand-in for training the last layer):
  * class logit : first principal direction, scaled to std 1, shifted so ~1.5 % of anchors pass conf 0.1
  * box logits  : seeded random mixtures of the 8 leading whitened directions, std 1.5 per output
Everything is deterministic in the seed.  Uses the fp32 oracle model on the CPU.

    PYTHONPATH=. python tools/calibrate_synthetic.py [seed ...]
"""
import json
import math
import sys

import numpy as np
import torch

from oracle import yolov8_ref as O
from wtracker_b200 import synth
from wtracker_b200.detector.weights import CALIB_PATH, synthetic_state_dict

CLS_STD, BOX_STD, PASS_FRAC, N_DIR = 1.0, 1.5, 0.015, 8


def calibrate(seed: int) -> dict:
    sd = synthetic_state_dict(seed, calibrated=False)
    model = O.build_model(sd)
    track = synth.worm_track(2000, seed)
    views = []
    for i in range(6):
        f = synth.render_frame(i * 300, track, seed)
        pos = (int(track[i * 300, 0]) + 25 * (i - 3), int(track[i * 300, 1]) + 15 * (i - 2))
        views.append(np.ascontiguousarray(synth.camera_view(f, pos, 640)))
    x = O.preprocess(views, 640)
    taps = {}
    with torch.no_grad():
        model.features(x, taps)
    det = model.model[22]
    g = torch.Generator().manual_seed(1000 + seed)
    out = {}
    for lvl, name in enumerate(("x15", "x18", "x21")):
        with torch.no_grad():
            fb = det.cv2[lvl][1](det.cv2[lvl][0](taps[name]))     # (n, 64, h, w)
            fc = det.cv3[lvl][1](det.cv3[lvl][0](taps[name]))     # (n, 128, h, w)
        for key, f in (("cv2", fb), ("cv3", fc)):
            m = f.permute(0, 2, 3, 1).reshape(-1, f.shape[1]).double()
            mu = m.mean(0)
            cov = (m - mu).T @ (m - mu) / m.shape[0]
            lam, vec = torch.linalg.eigh(cov)
            lam, vec = lam.flip(0), vec.flip(1)
            if key == "cv3":
                w = vec[:, 0] * (CLS_STD / math.sqrt(float(lam[0])))
                y = m @ w
                q = float(torch.quantile(y, 1.0 - PASS_FRAC))
                out[f"model.22.cv3.{lvl}.2"] = {"weight": [w.float().tolist()], "bias": [math.log(0.1 / 0.9) - q]}
            else:
                mix = torch.randn(64, N_DIR, generator=g).double() * (BOX_STD / math.sqrt(N_DIR))
                w = mix @ (vec[:, :N_DIR] / lam[:N_DIR].sqrt()).T          # (64 outputs, 64 inputs)
                b = 1.0 - (w @ mu)                                         # centre the logits on 1 (ultralytics' init bias)
                out[f"model.22.cv2.{lvl}.2"] = {"weight": w.float().tolist(), "bias": b.float().tolist()}
    return out


if __name__ == "__main__":
    seeds = [int(a) for a in sys.argv[1:]] or [0]
    data = json.loads(CALIB_PATH.read_text()) if CALIB_PATH.exists() else {}
    for seed in seeds:
        data[f"s-nc1-seed{seed}"] = calibrate(seed)
    CALIB_PATH.parent.mkdir(exist_ok=True)
    CALIB_PATH.write_text(json.dumps(data))
    print("wrote", CALIB_PATH, {k: list(v) for k, v in data.items()})
