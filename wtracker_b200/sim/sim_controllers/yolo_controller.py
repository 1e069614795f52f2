"""YOLO-in-the-loop controller backed by the CUDA detector (reference:
wtracker/sim/sim_controllers/yolo_controller.py — YoloConfig :15-45, YoloController :48-109)."""

from __future__ import annotations

import os
from collections import deque
from dataclasses import dataclass, field
from typing import Any, Collection

import numpy as np

from wtracker_b200.sim.simulator import SimController, Simulator
from wtracker_b200.utils.config_base import ConfigBase


_PRED_KWARGS = {"imgsz", "conf", "iou", "max_det", "verbose", "device", "half"}   # what predict() honours or may ignore safely


class B200Yolo:
    """What ``YoloConfig.load_model()`` returns: detector weights plus lazily built engines, one per
    (view size, imgsz, batch, thresholds).  Stands where ``ultralytics.YOLO`` stands in the reference."""

    def __init__(self, model_path: str, device: str = "cuda:0"):
        from wtracker_b200.detector.weights import load_ultralytics_checkpoint, synthetic_state_dict

        self.device = "cuda:0" if device in ("cuda", "gpu") else device
        if isinstance(model_path, str) and model_path.startswith("synthetic"):
            seed = int(model_path.split(":")[1]) if ":" in model_path else 0
            self.state_dict = synthetic_state_dict(seed)
        elif os.path.exists(model_path):
            self.state_dict = load_ultralytics_checkpoint(model_path)
        else:
            raise FileNotFoundError(f"YOLO weights not found: {model_path} (use 'synthetic:<seed>' for seeded weights)")
        self._engines: dict[tuple, Any] = {}

    def engine(self, view_hw: tuple[int, int], imgsz: int, conf: float, iou: float, max_det: int, batch: int):
        from wtracker_b200.detector.engine import DetectorEngine

        key = (tuple(view_hw), imgsz, conf, iou, max_det)
        eng = self._engines.get(key)
        if eng is None or eng.batch < batch:
            eng = DetectorEngine(self.state_dict, view_hw, imgsz, batch=max(batch, 1), conf=conf, iou=iou,
                                 max_det=max_det, device=self.device)
            self._engines[key] = eng
        return eng


@dataclass
class YoloConfig(ConfigBase):
    model_path: str
    device: str = "cuda:0"          # the reference defaults to "cpu"; this build has no CPU path
    verbose: bool = False
    pred_kwargs: dict = field(default_factory=lambda: {"imgsz": 384, "conf": 0.1})
    model: Any = field(default=None, init=False, repr=False)

    def __getstate__(self) -> dict[str, Any]:
        state = self.__dict__.copy()
        del state["model"]   # never serialise the model
        return state

    def load_model(self) -> B200Yolo:
        if self.model is None:
            if str(self.device) == "cpu":
                raise RuntimeError("wtracker_b200 has no CPU detector; use device='cuda:0'")
            self.model = B200Yolo(self.model_path, self.device)
        return self.model


class LazyView:
    """A camera view that has not been cut out yet: the whole grey frame (a reference, no copy) and the view's origin in
    frame coordinates.  The reference's controller buffers one cropped view per frame (yolo_controller.py:58-59) although
    it detects on one in ``cycle_frame_num`` of them; here the buffer holds descriptors, and the frames that ARE detected
    travel whole to the GPU where the crop kernel cuts the views (replicate borders included).  ``np.asarray(view)``
    still gives the pixels the reference would have buffered."""

    __slots__ = ("frame", "x0", "y0", "w", "h")

    def __init__(self, frame: np.ndarray, x0: int, y0: int, w: int, h: int):
        self.frame, self.x0, self.y0, self.w, self.h = frame, int(x0), int(y0), int(w), int(h)

    @property
    def shape(self) -> tuple[int, int]:
        return self.w, self.h      # rows x columns exactly as ViewController._custom_view slices them (w/h swapped, square views)

    def __array__(self, dtype=None, copy=None):
        f = self.frame
        rows = np.clip(np.arange(self.y0, self.y0 + self.w), 0, f.shape[0] - 1)
        cols = np.clip(np.arange(self.x0, self.x0 + self.h), 0, f.shape[1] - 1)
        out = f[np.ix_(rows, cols)]
        return out if dtype is None else out.astype(dtype)


class YoloController(SimController):
    def __init__(self, timing_config, yolo_config: YoloConfig, lazy_views: bool = True):
        super().__init__(timing_config)
        self.yolo_config = yolo_config
        self._camera_frames = deque(maxlen=timing_config.cycle_frame_num)
        self._model = yolo_config.load_model()
        self._warned = False
        self.lazy_views = lazy_views     # False: buffer cropped views like the reference does

    def on_sim_start(self, sim: Simulator):
        self._camera_frames.clear()

    def on_camera_frame(self, sim: Simulator):
        frame = sim.view.current_frame() if self.lazy_views else None
        if frame is not None and frame.ndim == 2 and sim.view.camera_size[0] == sim.view.camera_size[1]:
            x0, y0, w, h = sim.view.camera_position
            self._camera_frames.append(LazyView(frame, x0, y0, w, h))
        else:
            self._camera_frames.append(sim.camera_view())

    def on_cycle_end(self, sim: Simulator):
        self._camera_frames.clear()

    def predict(self, frames: Collection[np.ndarray]) -> np.ndarray:
        """(n, 4) xywh of the best box per frame in view pixels, NaN row where nothing passes
        ``conf``; float32 if every frame has a box, float64 otherwise (yolo_controller.py:85-90)."""
        assert len(frames) > 0
        frames = list(frames)
        lazy = all(isinstance(f, LazyView) for f in frames)
        if not lazy:
            frames = [np.asarray(f) if isinstance(f, LazyView) else f for f in frames]
        if not lazy and frames[0].ndim == 3:
            # The reference's experiments are grey (FrameReader defaults to IMREAD_GRAYSCALE and predict() replicates the
            # channel, yolo_controller.py:68-69); the CUDA detector folds the three equal channels into layer 0.  A
            # genuinely coloured view would need the 3-channel first layer: refuse it instead of silently using one channel.
            for f in frames:
                if not (np.array_equal(f[..., 0], f[..., 1]) and np.array_equal(f[..., 0], f[..., 2])):
                    raise ValueError("wtracker_b200's detector takes grey views (three equal channels); got a colour frame")
            frames = [np.ascontiguousarray(f[..., 0]) for f in frames]
        kw = dict(self.yolo_config.pred_kwargs)
        unsupported = sorted(set(kw) - _PRED_KWARGS)
        if unsupported and not self._warned:
            import warnings

            warnings.warn(f"YoloController: pred_kwargs {unsupported} are not supported by the CUDA detector and are ignored "
                          f"(supported: {sorted(_PRED_KWARGS)})", stacklevel=2)
            self._warned = True
        eng = self._model.engine(frames[0].shape[:2], int(kw.get("imgsz", 384)), float(kw.get("conf", 0.1)),
                                 float(kw.get("iou", 0.7)), 1, max(len(frames), self.timing_config.cycle_frame_num))
        if lazy:      # whole frames to the device, views cut there
            boxes, counts = eng.detect_frames([f.frame for f in frames], [f.x0 for f in frames], [f.y0 for f in frames])
        else:
            boxes, counts = eng.detect_views(frames)
        rows = []
        for b, c in zip(boxes, counts):
            if c == 0:
                rows.append(np.full([4], np.nan))
            else:
                x1, y1, x2, y2 = b[0, :4]
                rows.append(np.array([x1, y1, x2 - x1, y2 - y1], dtype=np.float32))
        return np.stack(rows, axis=0)

    def begin_movement_prediction(self, sim: Simulator) -> None:
        pass

    def provide_movement_vector(self, sim: Simulator) -> tuple[int, int]:
        frame = self._camera_frames[-self.timing_config.pred_frame_num]
        bbox = self.predict([frame])[0]
        if not np.isfinite(bbox).all():
            return 0, 0
        mid_x, mid_y = bbox[0] + bbox[2] / 2, bbox[1] + bbox[3] / 2
        return round(mid_x - sim.view.camera_size[0] / 2), round(mid_y - sim.view.camera_size[1] / 2)

    def _cycle_predict_all(self, sim: Simulator) -> np.ndarray:
        return self.predict(self._camera_frames)
