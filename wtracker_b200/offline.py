"""Offline detection of a long video sharded by frame range (BASELINE configs[3]).

The crop schedule is fixed in advance (no feedback), so frames are independent: rank r of R takes the contiguous
range ``sharding.frame_range(total, r, R)``, runs it through crop -> YOLOv8s -> decode/NMS in batches and writes one
32-byte row per frame (``wt_result_rows``: what ``YoloController.predict`` would return for it, plus confidence, kept
anchor, frame index and a valid flag) into its on-device table; the ONLY collective is the final
``all_gather_into_tensor`` of those tables — 32 MB for a million frames.
"""

from __future__ import annotations

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.sharding import TABLE_COLS, frame_range, gather_result_table


def detect_range(engine: DetectorEngine, frames: torch.Tensor, schedule, lo: int, hi: int) -> torch.Tensor:
    """Frames [lo, hi) -> i32 [hi - lo][8] rows on the device.  ``schedule(first, n)`` gives the int32 arrays
    (pool frame index, crop x, crop y) of frames [first, first + n); the whole range's descriptors go to the device
    in one copy before the first launch."""
    lib = L.lib()
    dev = engine.device
    n = hi - lo
    table = torch.zeros((max(n, 0), TABLE_COLS), dtype=torch.int32, device=dev)
    if n <= 0:
        return table
    idx, cx, cy = schedule(lo, n)
    with torch.cuda.device(dev):
        desc = torch.from_numpy(np.stack([idx, cx, cy]).astype(np.int32)).pin_memory().to(dev, non_blocking=True)
        s = torch.cuda.current_stream().cuda_stream
        B = engine.batch
        for a in range(0, n, B):
            m = min(B, n - a)
            boxes, count = engine.detect_crops(frames, desc[0, a: a + m], desc[1, a: a + m], desc[2, a: a + m])
            L.check(lib.wt_result_rows(boxes.data_ptr(), count.data_ptr(), engine.max_det, lo + a,
                                       table[a: a + m].data_ptr(), m, s), "wt_result_rows")
    return table


def run_offline(engine: DetectorEngine, frames: torch.Tensor, schedule, total: int, rank: int, world: int):
    """This rank's share of ``total`` frames + the gather.  Returns (full table i32 [total][8] in frame order,
    detect_ms, gather_ms) with both times measured on the device."""
    lo, hi = frame_range(total, rank, world)
    with torch.cuda.device(engine.device):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        local = detect_range(engine, frames, schedule, lo, hi)
        e1.record()
        full = gather_result_table(local, total)
        e2.record()
        torch.cuda.synchronize()
    return full, e0.elapsed_time(e1), e1.elapsed_time(e2)


def decode_rows(table: torch.Tensor) -> dict[str, np.ndarray]:
    """Host view of a result table: xywh f32 [n][4], conf f32, anchor i32, frame i32, valid bool."""
    t = table.cpu().numpy()
    f = t.view(np.float32)
    return dict(xywh=f[:, :4].copy(), conf=f[:, 4].copy(), anchor=t[:, 5].copy(), frame=t[:, 6].copy(),
                valid=t[:, 7].astype(bool))
