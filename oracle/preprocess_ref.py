"""ORACLE (test infrastructure, not product code) — numpy restatement of the camera-view crop and
the letterbox resize on u8 grey frames.

Follows: ViewController.read / _calc_view_bbox / _custom_view
(/root/reference/wtracker/sim/view_controller.py:45-61,143-172) for the crop, and OpenCV's
cv2.resize(INTER_LINEAR, 8-bit) fixed-point algorithm (modules/imgproc/src/resize.cpp:
resizeGeneric_ set-up, HResizeLinear, VResizeLinear with FixedPtCast<int, uchar, 22>) for the
resampling that ultralytics' LetterBox calls.  Pinned against the real cv2 in
tests/test_oracle_cpu.py (cv2 is installed; ultralytics is not).
"""

from __future__ import annotations

import numpy as np


def crop_replicate(frame: np.ndarray, pos_xy: tuple[int, int], size_wh: tuple[int, int]) -> np.ndarray:
    """What ``frame_padded[y:y+h, x:x+w]`` of the reference yields, without building the padded frame."""
    w, h = size_wh
    x0 = int(pos_xy[0]) - w // 2
    y0 = int(pos_xy[1]) - h // 2
    ys = np.clip(np.arange(y0, y0 + h), 0, frame.shape[0] - 1)
    xs = np.clip(np.arange(x0, x0 + w), 0, frame.shape[1] - 1)
    return frame[np.ix_(ys, xs)]


def crop_reference_style(frame: np.ndarray, pos_xy: tuple[int, int], size_wh: tuple[int, int]) -> np.ndarray:
    """Literal restatement (pad whole frame, then slice) — used to pin crop_replicate."""
    w, h = size_wh
    px, py = w // 2, h // 2
    padded = np.pad(frame, ((py, py), (px, px)), mode="edge")
    x = int(pos_xy[0]) + px - w // 2
    y = int(pos_xy[1]) + py - h // 2
    return padded[y: y + w, x: x + h]   # w/h swapped exactly as view_controller.py:171 does


def _coeffs(src: int, dst: int, pin_borders: bool):
    scale = 1.0 / (dst / src)
    ofs = np.empty(dst, np.int64)
    coef = np.empty((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if pin_borders:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0)
        ofs[d] = s
        a0 = np.float32(np.float32(1.0) - f) * np.float32(2048)
        a1 = np.float32(f) * np.float32(2048)
        coef[d] = (int(np.rint(a0)), int(np.rint(a1)))
    return ofs, coef


def resize_linear_u8(img: np.ndarray, new_wh: tuple[int, int]) -> np.ndarray:
    """cv2.resize(img, new_wh, interpolation=INTER_LINEAR) for a 2-D u8 image."""
    new_w, new_h = new_wh
    h, w = img.shape
    xofs, xa = _coeffs(w, new_w, True)
    yofs, yb = _coeffs(h, new_h, False)
    src = img.astype(np.int64)
    x1 = np.minimum(xofs + 1, w - 1)
    hrows = src[:, xofs] * xa[:, 0][None, :] + src[:, x1] * xa[:, 1][None, :]      # (h, new_w), 11-bit scaled
    r0 = hrows[np.clip(yofs, 0, h - 1)]
    r1 = hrows[np.clip(yofs + 1, 0, h - 1)]
    b0, b1 = yb[:, 0][:, None], yb[:, 1][:, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_u8(view: np.ndarray, geom) -> np.ndarray:
    """Letterboxed grey image for a geometry object with src/new/dst/pad fields (pad value 114)."""
    img = view
    if (geom.new_w, geom.new_h) != (geom.src_w, geom.src_h):
        img = resize_linear_u8(view, (geom.new_w, geom.new_h))
    out = np.full((geom.dst_h, geom.dst_w), 114, np.uint8)
    out[geom.pad_top: geom.pad_top + geom.new_h, geom.pad_left: geom.pad_left + geom.new_w] = img
    return out
