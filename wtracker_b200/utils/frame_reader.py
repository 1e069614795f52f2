"""Frame sources for the simulator and the batched ingest path.

Interface parity with the reference (wtracker/utils/frame_reader.py: ``FrameReader`` :9-157, ``FrameStream``
:159-244, ``DummyReader`` :247-272 — same constructor arguments, properties and method names, so a reference
script keeps working), organised differently underneath: every source is a ``FrameSource`` (length, frame shape,
random access, batched access), the image-file reader, the constant reader and the in-memory reader are three small
implementations of it, and ``read_batch`` / ``pinned_batches`` feed ``HotPath.run_frames`` — many whole frames in one
pinned buffer per host->device copy instead of one ``cv.imread`` + ``copyMakeBorder`` per simulator step.
"""

from __future__ import annotations

import os
import re
from typing import Iterator, Sequence

import numpy as np


class FrameSource:
    """What a frame source must answer: ``len``, ``frame_shape`` and ``source[i] -> u8 array``."""

    _shape: tuple[int, ...] = ()

    def __len__(self) -> int:
        raise NotImplementedError

    def __getitem__(self, idx: int) -> np.ndarray:
        raise NotImplementedError

    def _check(self, idx: int) -> int:
        idx = int(idx)
        if not 0 <= idx < len(self):
            raise IndexError("index out of bounds")
        return idx

    @property
    def frame_shape(self) -> tuple[int, ...]:
        return self._shape

    @property
    def frame_size(self) -> tuple[int, int]:
        """(h, w)."""
        return self._shape[0], self._shape[1]

    def __iter__(self) -> "FrameStream":
        return FrameStream(self)

    def make_stream(self) -> "FrameStream":
        return FrameStream(self)

    # ---- batched access (ingest) --------------------------------------------------------------------------------
    def read_batch(self, indices: Sequence[int], out: np.ndarray | None = None) -> np.ndarray:
        """Frames ``indices`` stacked into ``out`` (allocated when missing): u8 [n, *frame_shape]."""
        n = len(indices)
        if out is None:
            out = np.empty((n, *self.frame_shape), dtype=np.uint8)
        for slot, idx in enumerate(indices):
            out[slot] = self[idx]
        return out[:n]

    def pinned_batches(self, batch: int, start: int = 0, stop: int | None = None) -> Iterator:
        """Consecutive frames in pinned torch tensors of up to ``batch`` frames, alternating between two buffers so
        that the copy of one batch can be in flight while the next one is read."""
        import torch

        stop = len(self) if stop is None else min(stop, len(self))
        ring = [torch.empty((batch, *self.frame_shape), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for k, first in enumerate(range(start, stop, batch)):
            idx = range(first, min(first + batch, stop))
            buf = ring[k & 1]
            self.read_batch(idx, buf.numpy())
            yield buf[: len(idx)]


_NUMBER = re.compile(r"\{[^{}]*\}")


class FrameReader(FrameSource):
    """Image files of one folder, decoded on access with cv2 (grayscale unless ``read_format`` says otherwise)."""

    def __init__(self, root_folder: str, frame_files: list[str], read_format: int | None = None):
        assert os.path.exists(root_folder)
        assert len(frame_files) > 0
        self._root_folder = root_folder
        self._files = list(frame_files)
        self._read_format = 0 if read_format is None else read_format   # cv.IMREAD_GRAYSCALE == 0
        self._shape = tuple(self._extract_frame_shape())

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return self[0].shape

    @staticmethod
    def _listing(root_folder: str, accept) -> list[str]:
        with os.scandir(root_folder) as entries:
            return sorted(e.name for e in entries if e.is_file() and accept(e.name))

    @staticmethod
    def create_from_template(root_folder: str, name_format: str, read_format: int | None = None) -> "FrameReader":
        """Files whose names are ``name_format`` with its ``{}`` field filled by a number (e.g. ``frame_{:09d}.png``)."""
        head, _, tail = _NUMBER.sub("\0", name_format, count=1).partition("\0")
        pattern = re.compile(re.escape(head) + r"[0-9].*" + re.escape(tail) + r"\Z")
        return FrameReader(root_folder, FrameReader._listing(root_folder, lambda n: pattern.match(n) is not None), read_format)

    @staticmethod
    def create_from_directory(root_folder: str, read_format: int | None = None) -> "FrameReader":
        return FrameReader(root_folder, FrameReader._listing(root_folder, lambda n: "." in n.strip(".")), read_format)

    @property
    def root_folder(self) -> str:
        return self._root_folder

    @property
    def files(self) -> list[str]:
        return self._files

    @property
    def read_format(self) -> int:
        return self._read_format

    def __len__(self) -> int:
        return len(self._files)

    def __getitem__(self, idx: int) -> np.ndarray:
        import cv2 as cv

        path = os.path.join(self._root_folder, self._files[self._check(idx)])
        image = cv.imread(path, self._read_format)
        if image is None:
            raise OSError(f"cannot decode {path}")
        return image.astype(np.uint8, copy=False)


class FrameStream:
    """Forward cursor over a frame source; the frame under the cursor is decoded once and kept until the cursor moves."""

    def __init__(self, frame_reader: FrameSource):
        self._frame_reader = frame_reader
        self._idx = -1
        self.frame: np.ndarray | None = None

    @property
    def index(self) -> int:
        return self._idx

    def __len__(self) -> int:
        return len(self._frame_reader)

    def can_read(self) -> bool:
        return -1 < self._idx < len(self._frame_reader)

    def seek(self, idx: int) -> bool:
        if idx != self._idx:
            self.frame = None
        self._idx = idx
        return self.can_read()

    def progress(self, n: int = 1) -> bool:
        return self.seek(self._idx + n)

    def reset(self) -> None:
        self.seek(-1)

    def read(self) -> np.ndarray:
        if not self.can_read():
            raise IndexError("index out of bounds")
        if self.frame is None:
            self.frame = self._frame_reader[self._idx]
        return self.frame

    def __iter__(self) -> "FrameStream":
        return self

    def __next__(self) -> np.ndarray:
        if not self.progress():
            raise StopIteration()
        return self.read()


class DummyReader(FrameReader):
    """``num_frames`` constant all-255 frames of ``resolution`` (h, w) — a simulation that needs no pixels."""

    def __init__(self, num_frames: int, resolution: tuple[int, int], colored: bool = True):
        self.colored = colored
        self._resolution = tuple(resolution)
        self._root_folder, self._read_format = ".", 0
        self._files = [str(i) for i in range(num_frames)]
        self._shape = self._extract_frame_shape()
        self._frame = np.full(self._shape, 255, dtype=np.uint8)

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return (*self._resolution, 3) if self.colored else self._resolution

    def __getitem__(self, idx: int) -> np.ndarray:
        return self._frame.copy()


class ArrayReader(FrameReader):
    """Frames held in memory as one (n, h, w[, 3]) u8 array (synthetic experiments, device upload)."""

    def __init__(self, frames: np.ndarray):
        assert frames.dtype == np.uint8 and frames.ndim in (3, 4)
        self._frames = frames
        self._root_folder, self._read_format = ".", 0
        self._files = [str(i) for i in range(frames.shape[0])]
        self._shape = tuple(frames.shape[1:])

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return tuple(self._frames.shape[1:])

    def __getitem__(self, idx: int) -> np.ndarray:
        return self._frames[self._check(idx)]

    def read_batch(self, indices, out=None):
        block = self._frames[np.asarray(list(indices), dtype=np.int64)]
        if out is None:
            return block
        out[: block.shape[0]] = block
        return out[: block.shape[0]]

    @property
    def array(self) -> np.ndarray:
        return self._frames
