"""Frame sources (reference: wtracker/utils/frame_reader.py — FrameReader :9-157, FrameStream
:159-244, DummyReader :247-272) plus ``ArrayReader`` for frames already in memory (synthetic
experiments, device upload)."""

from __future__ import annotations

import glob
import os

import numpy as np


class FrameReader:
    """Indexable collection of image files read with cv2 (grayscale by default)."""

    def __init__(self, root_folder: str, frame_files: list[str], read_format: int | None = None):
        assert os.path.exists(root_folder)
        assert len(frame_files) > 0
        self._root_folder = root_folder
        self._files = frame_files
        self._read_format = 0 if read_format is None else read_format   # cv.IMREAD_GRAYSCALE == 0
        self._frame_shape = self._extract_frame_shape()

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return self[0].shape

    @staticmethod
    def create_from_template(root_folder: str, name_format: str, read_format: int | None = None) -> "FrameReader":
        paths = sorted(p for p in glob.glob(name_format.format("[0-9]*"), root_dir=root_folder)
                       if os.path.isfile(os.path.join(root_folder, p)))
        return FrameReader(root_folder, paths, read_format)

    @staticmethod
    def create_from_directory(root_folder: str, read_format: int | None = None) -> "FrameReader":
        paths = sorted(p for p in glob.glob("*.*", root_dir=root_folder) if os.path.isfile(os.path.join(root_folder, p)))
        return FrameReader(root_folder, paths, read_format)

    @property
    def root_folder(self) -> str:
        return self._root_folder

    @property
    def frame_shape(self) -> tuple[int, ...]:
        return self._frame_shape

    @property
    def frame_size(self) -> tuple[int, int]:
        return self._frame_shape[:2]

    @property
    def files(self) -> list[str]:
        return self._files

    @property
    def read_format(self) -> int:
        return self._read_format

    def __len__(self) -> int:
        return len(self._files)

    def __getitem__(self, idx: int) -> np.ndarray:
        if idx < 0 or idx >= len(self._files):
            raise IndexError("index out of bounds")
        import cv2 as cv

        frame = cv.imread(os.path.join(self._root_folder, self._files[idx]), self._read_format)
        return frame.astype(np.uint8, copy=False)

    def __iter__(self):
        return FrameStream(self)

    def make_stream(self):
        return FrameStream(self)


class FrameStream:
    """Cursor over a FrameReader with a one-frame cache."""

    def __init__(self, frame_reader: FrameReader):
        self._frame_reader = frame_reader
        self._idx = -1
        self.frame = None

    @property
    def index(self) -> int:
        return self._idx

    def __len__(self):
        return len(self._frame_reader)

    def __iter__(self):
        return self

    def __next__(self) -> np.ndarray:
        self.progress()
        if not self.can_read():
            raise StopIteration()
        return self.read()

    def can_read(self) -> bool:
        return 0 <= self._idx < len(self._frame_reader)

    def seek(self, idx: int) -> bool:
        self._idx = idx
        self.frame = None
        return self.can_read()

    def read(self) -> np.ndarray:
        if not self.can_read():
            raise IndexError("index out of bounds")
        if self.frame is None:
            self.frame = self._frame_reader[self._idx]
        return self.frame

    def progress(self, n: int = 1) -> bool:
        return self.seek(self._idx + n)

    def reset(self):
        self.seek(-1)


class DummyReader(FrameReader):
    """Constant all-255 frames of a given (h, w) resolution (used when a simulation needs no pixels)."""

    def __init__(self, num_frames: int, resolution: tuple[int, int], colored: bool = True):
        self.colored = colored
        self._resolution = tuple(resolution)
        shape = (*self._resolution, 3) if colored else self._resolution
        self._frame = np.full(shape, 255, dtype=np.uint8)
        super().__init__(".", [str(i) for i in range(num_frames)])

    def __getitem__(self, idx: int) -> np.ndarray:
        return self._frame.copy()

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return (*self._resolution, 3) if self.colored else self._resolution


class ArrayReader(FrameReader):
    """Frames held in memory as one (n, h, w[, 3]) u8 array."""

    def __init__(self, frames: np.ndarray):
        assert frames.dtype == np.uint8 and frames.ndim in (3, 4)
        self._frames = frames
        super().__init__(".", [str(i) for i in range(frames.shape[0])])

    def __getitem__(self, idx: int) -> np.ndarray:
        if idx < 0 or idx >= self._frames.shape[0]:
            raise IndexError("index out of bounds")
        return self._frames[idx]

    def _extract_frame_shape(self) -> tuple[int, ...]:
        return tuple(self._frames.shape[1:])

    @property
    def array(self) -> np.ndarray:
        return self._frames
