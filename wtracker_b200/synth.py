"""Seeded synthetic experiment data: grey u8 frames with one dark worm on a noisy
bright background, and the worm's ground-truth trajectory.  Counter-based, so any frame can be
regenerated independently from ``(seed, frame_idx)`` — the oracle, the tests and the bench all see
identical pixels.  Pure numpy; no reference code involved.
"""

from __future__ import annotations

import numpy as np

FRAME_H, FRAME_W = 1080, 1920


def _hash_u32(x: np.ndarray) -> np.ndarray:
    """Integer avalanche hash (uint32 -> uint32)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def worm_track(num_frames: int, seed: int = 0, frame_hw: tuple[int, int] = (FRAME_H, FRAME_W),
               margin: int = 200, border_visit: bool = True) -> np.ndarray:
    """(num_frames, 3) float64: worm head centre x, y and heading angle.  Sum of three sinusoids
    per axis, speed <= ~1.35 px/frame (0.9 mm/s at 90 px/mm, 60 fps); optionally one excursion that
    reaches the frame border to exercise replicate padding."""
    rng = np.random.default_rng(seed)
    h, w = frame_hw
    t = np.arange(num_frames, dtype=np.float64)
    pos = np.zeros((num_frames, 2))
    for axis, extent in enumerate((w, h)):
        span = (extent - 2 * margin) / 2
        amp = rng.uniform(0.15, 0.45, 3)
        amp *= span / amp.sum()
        period = rng.uniform(900, 4000, 3)
        # keep per-axis speed below ~0.9 px/frame so the vector speed stays under 1.35
        max_speed = (amp * 2 * np.pi / period).sum()
        if max_speed > 0.9:
            period *= max_speed / 0.9
        phase = rng.uniform(0, 2 * np.pi, 3)
        pos[:, axis] = extent / 2 + (amp[None, :] * np.sin(2 * np.pi * t[:, None] / period[None, :] + phase[None, :])).sum(1)
    if border_visit and num_frames >= 600:
        # smooth detour of the x coordinate toward the left border in the last third
        s = np.clip((t - 0.7 * num_frames) / (0.25 * num_frames), 0, 1)
        bump = np.sin(np.pi * s) ** 2
        pos[:, 0] = pos[:, 0] * (1 - bump) + 20.0 * bump
    vel = np.gradient(pos, axis=0) if num_frames > 1 else np.zeros_like(pos)
    ang = np.arctan2(vel[:, 1], vel[:, 0] + 1e-12)
    return np.concatenate([pos, ang[:, None]], axis=1)


def render_frame(frame_idx: int, track: np.ndarray, seed: int = 0,
                 frame_hw: tuple[int, int] = (FRAME_H, FRAME_W)) -> np.ndarray:
    """One (h, w) u8 frame: background 200 +- 6 hash noise; worm = tapered body (~80 px long,
    8 px wide at the head, fading from ~70 to ~140 towards the tail) behind a dark ~14 px head disc (~24)."""
    h, w = frame_hw
    yy, xx = np.meshgrid(np.arange(h, dtype=np.uint32), np.arange(w, dtype=np.uint32), indexing="ij")
    key = np.uint32((int(seed) * 0x9E3779B1 ^ int(frame_idx) * 0x85EBCA77) & 0xFFFFFFFF)
    noise = _hash_u32(xx * np.uint32(0x27D4EB2F) ^ yy * np.uint32(0x165667B1) ^ key) % np.uint32(13)
    img = (194 + noise).astype(np.uint8)

    cx, cy, ang = track[frame_idx]
    # local window around the worm
    x0, x1 = int(max(0, cx - 110)), int(min(w, cx + 110))
    y0, y1 = int(max(0, cy - 110)), int(min(h, cy + 110))
    if x1 > x0 and y1 > y0:
        ys, xs = np.meshgrid(np.arange(y0, y1, dtype=np.float64), np.arange(x0, x1, dtype=np.float64), indexing="ij")
        dx, dy = xs - cx, ys - cy
        ca, sa = np.cos(ang), np.sin(ang)
        u = dx * ca + dy * sa          # along heading (head at u = 0, body trails to u = -80)
        v = -dx * sa + dy * ca
        uc = np.clip(u, -80.0, 0.0)
        taper = 1.0 + 0.75 * uc / 80.0                       # 1 at the head .. 0.25 at the tail tip
        body = (u - uc) ** 2 + v ** 2 <= (4.0 * taper) ** 2
        head = dx ** 2 + dy ** 2 <= 7.0 ** 2
        sub = img[y0:y1, x0:x1]
        nz = noise[y0:y1, x0:x1] % np.uint32(5)
        shade = (68.0 - 70.0 * uc / 80.0).astype(np.uint8) + nz.astype(np.uint8)   # body fades towards the tail
        sub[body] = shade[body]
        darker = (22 + nz).astype(np.uint8)                  # the head is the darkest part
        sub[head] = darker[head]
    return img


def head_bbox(track: np.ndarray) -> np.ndarray:
    """Ground-truth head boxes (x, y, w, h) in frame px, 14 x 14 around the head centre."""
    out = np.empty((track.shape[0], 4))
    out[:, 0] = track[:, 0] - 7
    out[:, 1] = track[:, 1] - 7
    out[:, 2] = 14
    out[:, 3] = 14
    return out


def make_frames(num_frames: int, seed: int = 0, frame_hw: tuple[int, int] = (FRAME_H, FRAME_W),
                border_visit: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """(frames u8 [n, h, w], track [n, 3])."""
    track = worm_track(num_frames, seed, frame_hw, border_visit=border_visit)
    frames = np.stack([render_frame(i, track, seed, frame_hw) for i in range(num_frames)])
    return frames, track


def camera_view(frame: np.ndarray, pos_xy: tuple[int, int], size: int) -> np.ndarray:
    """Camera view centred at ``pos_xy`` with replicate border — numpy statement of the crop the
    reference makes with copyMakeBorder + slicing (wtracker/sim/view_controller.py:45-61,143-172)."""
    h, w = frame.shape
    x0 = int(pos_xy[0]) - size // 2
    y0 = int(pos_xy[1]) - size // 2
    ys = np.clip(np.arange(y0, y0 + size), 0, h - 1)
    xs = np.clip(np.arange(x0, x0 + size), 0, w - 1)
    return frame[np.ix_(ys, xs)]


def render_frames_device(track: np.ndarray, seed: int, device, frame_hw: tuple[int, int] = (FRAME_H, FRAME_W),
                         first: int = 0, count: int | None = None):
    """``render_frame`` for frames [first, first + count) of ``track`` written straight into a u8 [count, h, w]
    tensor on ``device`` (the sweep / offline workloads hold hundreds of frames in HBM; rendering them with numpy
    and uploading costs ~30 ms each).  Same integer hash and the same float64 operations one by one (separate
    elementwise kernels: no FMA contraction), so the pixels equal ``render_frame``'s (tests/test_gpu_batched.py).
    Data generation only — not part of the hot path."""
    import torch

    h, w = frame_hw
    count = track.shape[0] - first if count is None else count
    dev = torch.device(device)
    M = 0xFFFFFFFF
    yy = torch.arange(h, dtype=torch.int64, device=dev)[:, None]
    xx = torch.arange(w, dtype=torch.int64, device=dev)[None, :]
    base = ((xx * 0x27D4EB2F) & M) ^ ((yy * 0x165667B1) & M)

    def hash_u32(x):
        x = x ^ (x >> 16)
        x = (x * 0x7FEB352D) & M
        x = x ^ (x >> 15)
        x = (x * 0x846CA68B) & M      # (the int64 product wraps, its low 32 bits are the uint32 product)
        return x ^ (x >> 16)

    out = torch.empty((count, h, w), dtype=torch.uint8, device=dev)
    for j in range(count):
        fi = first + j
        key = (int(seed) * 0x9E3779B1 ^ int(fi) * 0x85EBCA77) & M
        noise = hash_u32(base ^ key) % 13
        img = (194 + noise).to(torch.uint8)
        cx, cy, ang = (float(v) for v in track[fi])
        x0, x1 = int(max(0, cx - 110)), int(min(w, cx + 110))
        y0, y1 = int(max(0, cy - 110)), int(min(h, cy + 110))
        if x1 > x0 and y1 > y0:
            ys = torch.arange(y0, y1, dtype=torch.float64, device=dev)[:, None]
            xs = torch.arange(x0, x1, dtype=torch.float64, device=dev)[None, :]
            dx, dy = xs - cx, ys - cy
            ca, sa = float(np.cos(ang)), float(np.sin(ang))
            u = dx * ca + dy * sa
            v = -dx * sa + dy * ca
            uc = torch.clamp(u, -80.0, 0.0)
            taper = 1.0 + 0.75 * uc / 80.0
            du = u - uc
            body = du * du + v * v <= (4.0 * taper) * (4.0 * taper)
            head = dx * dx + dy * dy <= 7.0 ** 2
            sub = img[y0:y1, x0:x1]
            nz = (noise[y0:y1, x0:x1] % 5).to(torch.uint8)
            shade = (68.0 - 70.0 * uc / 80.0).to(torch.uint8) + nz
            sub = torch.where(body, shade, sub)
            sub = torch.where(head, 22 + nz, sub)
            img[y0:y1, x0:x1] = sub
        out[j] = img
    return out
