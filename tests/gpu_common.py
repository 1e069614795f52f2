"""Shared helpers of the -m gpu parity tests (no reference code is read at run time)."""
import functools

import numpy as np
import torch

from wtracker_b200 import synth


@functools.lru_cache(maxsize=None)
def synthetic_sd(seed=0):
    from wtracker_b200.detector.weights import synthetic_state_dict

    return synthetic_state_dict(seed)


@functools.lru_cache(maxsize=None)
def oracle_model(seed=0):
    from oracle import yolov8_ref as Y

    return Y.build_model(synthetic_sd(seed))


@functools.lru_cache(maxsize=None)
def sample_frames(n=6, seed=0, hw=(1080, 1920)):
    track = synth.worm_track(2000, seed, hw) if hw == (1080, 1920) else synth.worm_track(2000, seed, hw, margin=20)
    idx = [i * (1999 // max(n - 1, 1)) for i in range(n)]
    frames = np.stack([synth.render_frame(i, track, seed, hw) for i in idx])
    return frames, track[idx]


def views_for(size, n=4, seed=0):
    frames, tr = sample_frames(6, seed)
    out = []
    for i in range(n):
        pos = (int(tr[i % 6, 0]) + 17 * i - 20, int(tr[i % 6, 1]) - 11 * i + 8)
        out.append(np.ascontiguousarray(synth.camera_view(frames[i % 6], pos, size)))
    return out


def box_iou(a, b):
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    x2, y2 = min(a[2], b[2]), min(a[3], b[3])
    inter = max(0.0, x2 - x1) * max(0.0, y2 - y1)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 1.0
