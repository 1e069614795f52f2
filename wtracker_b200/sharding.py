"""Multi-GPU plan: frames / experiments are independent, so each rank takes a
contiguous range and holds a full weight replica; the ONLY collective is the final gather of the
per-frame result table (32 B per frame).  Works with NCCL (GPU) and gloo (CPU tests)."""

from __future__ import annotations

import torch
import torch.distributed as dist

TABLE_COLS = 8   # x, y, w, h (view px, NaN = no detection), conf, kept anchor index, frame index, valid


def frame_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) share of ``total`` units for ``rank``."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def result_table_from(boxes: torch.Tensor, counts: torch.Tensor, frame_idx: torch.Tensor) -> torch.Tensor:
    """boxes [n, max_det, 6] (x1, y1, x2, y2, conf, anchor), counts [n] -> table [n, 8] fp32 holding the
    best box as xywh (what YoloController.predict returns), NaN where nothing was detected."""
    b = boxes[:, 0, :].float()
    valid = counts > 0
    nan = torch.full_like(b[:, 0], float("nan"))
    cols = [torch.where(valid, b[:, 0], nan), torch.where(valid, b[:, 1], nan),
            torch.where(valid, b[:, 2] - b[:, 0], nan), torch.where(valid, b[:, 3] - b[:, 1], nan),
            torch.where(valid, b[:, 4], nan), torch.where(valid, b[:, 5], nan), frame_idx.float(), valid.float()]
    return torch.stack(cols, 1).contiguous()


def gather_result_table(local: torch.Tensor, total: int) -> torch.Tensor:
    """All ranks contribute their [n_r, 8] rows; every rank gets the [total, 8] table in frame order.
    Ranges differ by at most one row, so rows are padded to the longest share for all_gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    longest = (total + world - 1) // world
    padded = torch.zeros((longest, TABLE_COLS), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * longest, TABLE_COLS), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    parts = []
    for r in range(world):
        lo, hi = frame_range(total, r, world)
        parts.append(out[r * longest: r * longest + (hi - lo)])
    return torch.cat(parts, 0)
