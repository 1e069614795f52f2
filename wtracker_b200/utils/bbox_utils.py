"""Bounding-box helpers with the reference's names and semantics
(wtracker/utils/bbox_utils.py: BoxFormat :6-18, BoxUtils :21-167, BoxConverter :170-292).

Host-side glue of the hot path (tiny arrays); written column-wise on views instead of the
reference's split/concatenate round trips, same results bit for bit (float64 numpy ops).
"""

from __future__ import annotations

from enum import Enum

import numpy as np


class BoxFormat(Enum):
    XYWH = 0   # (x, y, width, height), x/y = top-left corner
    XYXY = 1   # (x1, y1, x2, y2)
    YOLO = 2   # (centre x, centre y, width, height)


def _cols(b: np.ndarray):
    return b[..., 0], b[..., 1], b[..., 2], b[..., 3]


def _stack(c0, c1, c2, c3) -> np.ndarray:
    return np.stack((c0, c1, c2, c3), axis=-1)


class BoxUtils:
    @staticmethod
    def is_bbox(array: np.ndarray) -> bool:
        return array.shape[-1] == 4

    @staticmethod
    def unpack(bbox: np.ndarray):
        c0, c1, c2, c3 = _cols(bbox)
        return c0.copy(), c1.copy(), c2.copy(), c3.copy()

    @staticmethod
    def pack(c1, c2, c3, c4) -> np.ndarray:
        return _stack(np.asarray(c1), np.asarray(c2), np.asarray(c3), np.asarray(c4))

    @staticmethod
    def center(bboxes: np.ndarray, box_format: BoxFormat = BoxFormat.XYWH) -> np.ndarray:
        """(N, 2) / (2,) box centres: x + w/2, y + h/2 (bbox_utils.py:77-92)."""
        b = BoxConverter.change_format(bboxes, box_format, BoxFormat.XYWH)
        x, y, w, h = _cols(b)
        return np.array([x + w / 2, y + h / 2]).T

    @staticmethod
    def round(bboxes: np.ndarray, box_format: BoxFormat) -> np.ndarray:
        """Outward rounding to int32: floor the top-left, ceil the bottom-right (:94-117)."""
        b = BoxConverter.change_format(bboxes, box_format, BoxFormat.XYXY)
        x1, y1, x2, y2 = _cols(b)
        out = _stack(np.floor(x1).astype(np.int32, copy=False), np.floor(y1).astype(np.int32, copy=False),
                     np.ceil(x2).astype(np.int32, copy=False), np.ceil(y2).astype(np.int32, copy=False))
        return BoxConverter.change_format(out, BoxFormat.XYXY, box_format)

    @staticmethod
    def discretize(bboxes: np.ndarray, bounds: tuple[int, int], box_format: BoxFormat):
        """Integer boxes clipped to ``bounds`` = (h, w) plus a legality mask (:119-167).
        Like the reference this ZEROES non-finite rows of the caller's array in place."""
        finite = np.isfinite(bboxes).all(axis=1)
        bboxes[~finite] = 0
        b = BoxUtils.round(BoxConverter.change_format(bboxes, box_format, BoxFormat.XYXY), BoxFormat.XYXY)
        x1, y1, x2, y2 = _cols(b)
        H, W = bounds
        x1, x2 = np.clip(x1, 0, W), np.clip(x2, 0, W)
        y1, y2 = np.clip(y1, 0, H), np.clip(y2, 0, H)
        out = BoxConverter.change_format(_stack(x1, y1, x2, y2), BoxFormat.XYXY, box_format)
        legal = ((x2 - x1) > 0.0) & ((y2 - y1) > 0.0)
        out[~legal] = 0
        return out.astype(np.int32, copy=False), legal.astype(bool, copy=False)


class BoxConverter:
    @staticmethod
    def change_format(bbox: np.ndarray, src_format: BoxFormat, dst_format: BoxFormat) -> np.ndarray:
        if dst_format == BoxFormat.XYXY:
            return BoxConverter.to_xyxy(bbox, src_format)
        if dst_format in (BoxFormat.XYWH, BoxFormat.YOLO):
            # the reference routes a YOLO destination through to_xywh as well (bbox_utils.py:193-194)
            return BoxConverter.to_xywh(bbox, src_format)
        raise Exception("unsupported bbox format conversion.")

    @staticmethod
    def to_xyxy(bbox: np.ndarray, src_format: BoxFormat) -> np.ndarray:
        if src_format == BoxFormat.XYXY:
            return bbox
        a, b, w, h = _cols(bbox)
        if src_format == BoxFormat.XYWH:
            return _stack(a, b, a + w, b + h)
        if src_format == BoxFormat.YOLO:
            x1, y1 = a - w / 2, b - h / 2
            return _stack(x1, y1, x1 + w, y1 + h)
        raise Exception("unsupported bbox format conversion.")

    @staticmethod
    def to_xywh(bbox: np.ndarray, src_format: BoxFormat) -> np.ndarray:
        if src_format == BoxFormat.XYWH:
            return bbox
        a, b, c, d = _cols(bbox)
        if src_format == BoxFormat.XYXY:
            return _stack(a, b, c - a, d - b)
        if src_format == BoxFormat.YOLO:
            return _stack(a - c / 2, b - d / 2, c, d)
        raise Exception("unsupported bbox format conversion.")

    @staticmethod
    def to_yolo(bbox: np.ndarray, src_format: BoxFormat) -> np.ndarray:
        if src_format == BoxFormat.YOLO:
            return bbox
        a, b, c, d = _cols(bbox)
        if src_format == BoxFormat.XYXY:
            w, h = c - a, d - b
            return _stack(a + w / 2, b + h / 2, w, h)
        if src_format == BoxFormat.XYWH:
            return _stack(a + c / 2, b + d / 2, c, d)
        raise Exception("unsupported bbox format conversion.")
