"""Host-side geometry of the ultralytics LetterBox (auto=True, stride 32, centred, pad 114) and the
cv2.resize(INTER_LINEAR) fixed-point coefficient tables the pre-process kernel consumes.

The tables are built with the same float32/float64 operations OpenCV uses (imgproc/resize.cpp,
``resizeGeneric_`` set-up for INTER_LINEAR on 8-bit data: 11-bit coefficients), so the kernel's
integer arithmetic reproduces cv2 pixels.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

INTER_BITS = 11
INTER_SCALE = 1 << INTER_BITS


@dataclass(frozen=True)
class Letterbox:
    src_w: int
    src_h: int
    dst_w: int
    dst_h: int
    new_w: int
    new_h: int
    pad_left: int
    pad_top: int

    @property
    def resample(self) -> bool:
        return (self.new_w, self.new_h) != (self.src_w, self.src_h)

    @property
    def gain(self) -> float:
        """scale_boxes gain (net / original)."""
        return min(self.dst_h / self.src_h, self.dst_w / self.src_w)

    @property
    def scale_pad(self) -> tuple[float, float]:
        """scale_boxes padding (x, y): round((net - img * gain) / 2 - 0.1)."""
        g = self.gain
        return float(round((self.dst_w - self.src_w * g) / 2 - 0.1)), float(round((self.dst_h - self.src_h * g) / 2 - 0.1))


def letterbox_for(src_hw: tuple[int, int], imgsz: int, stride: int = 32, auto: bool = True) -> Letterbox:
    h, w = src_hw
    r = min(imgsz / h, imgsz / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - new_w, imgsz - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return Letterbox(w, h, new_w + left + right, new_h + top + bottom, new_w, new_h, left, top)


def _axis_table(src: int, dst: int, clamp_coef: bool) -> tuple[np.ndarray, np.ndarray]:
    """(offsets int32 [dst], coefficients int16 [dst, 2]) for one axis."""
    scale = 1.0 / (float(dst) / float(src))                        # double, like cv::resize
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)               # (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_coef:                                                 # x axis: taps pinned at the borders
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0.0, src - 1
    c = np.stack([np.float32(1.0) - f, f], axis=1).astype(np.float32) * np.float32(INTER_SCALE)
    coef = np.rint(c).astype(np.int32).clip(-32768, 32767).astype(np.int16)   # saturate_cast<short>(cvRound)
    return s, coef


def resize_tables(lb: Letterbox) -> dict[str, np.ndarray]:
    xofs, xcoef = _axis_table(lb.src_w, lb.new_w, clamp_coef=True)
    yofs, ycoef = _axis_table(lb.src_h, lb.new_h, clamp_coef=False)   # rows are clipped when read instead
    return {"xofs": xofs, "xcoef": xcoef.reshape(-1), "yofs": yofs, "ycoef": ycoef.reshape(-1)}
