"""Log-driven controller base (reference: wtracker/sim/sim_controllers/csv_controller.py)."""

from __future__ import annotations

from collections import deque
from typing import Collection

import numpy as np

from wtracker_b200.sim.simulator import SimController, Simulator
from wtracker_b200.utils.bbox_utils import BoxUtils

WORM_COLS = ["wrm_x", "wrm_y", "wrm_w", "wrm_h"]


class CsvController(SimController):
    """Serves worm boxes from a bbox table.  ``csv_path`` may also be an (F, 4) array (already in
    memory, e.g. the table the CUDA detector just produced)."""

    def __init__(self, timing_config, csv_path):
        super().__init__(timing_config)
        self.csv_path = csv_path
        if isinstance(csv_path, np.ndarray):
            self._csv_data = np.asarray(csv_path, dtype=float).reshape(-1, 4)
        else:
            import pandas as pd

            self._csv_data = pd.read_csv(csv_path, usecols=WORM_COLS).to_numpy(dtype=float)
        self._camera_bboxes = deque(maxlen=timing_config.cycle_frame_num)

    def on_sim_start(self, sim: Simulator):
        self._camera_bboxes.clear()

    def on_camera_frame(self, sim: Simulator):
        self._camera_bboxes.append(sim.view.camera_position)

    def predict(self, frame_nums: Collection[int], relative: bool = True) -> np.ndarray:
        """Rows of the table for ``frame_nums`` (NaN outside the table); with ``relative`` the
        x, y are shifted into the camera view recorded for that frame of the current cycle."""
        assert len(frame_nums) > 0
        idx = np.asanyarray(frame_nums, dtype=int)
        ok = (idx >= 0) & (idx < self._csv_data.shape[0])
        boxes = np.full((idx.shape[0], 4), np.nan)
        boxes[ok] = self._csv_data[idx[ok], :]
        if relative:
            n = self.timing_config.cycle_frame_num
            cams = np.asanyarray([self._camera_bboxes[i % n] for i in idx], dtype=float)
            boxes[:, 0] -= cams[:, 0]
            boxes[:, 1] -= cams[:, 1]
        return boxes

    def begin_movement_prediction(self, sim: Simulator) -> None:
        pass

    def provide_movement_vector(self, sim: Simulator) -> tuple[int, int]:
        bbox = self.predict([sim.frame_number - self.timing_config.pred_frame_num])[0, :]
        if not np.isfinite(bbox).all():
            return 0, 0
        cx, cy = BoxUtils.center(bbox)
        return round(cx - sim.view.camera_size[0] / 2), round(cy - sim.view.camera_size[1] / 2)

    def _cycle_predict_all(self, sim: Simulator) -> np.ndarray:
        n = self.timing_config.cycle_frame_num
        start = (sim.cycle_number - 1) * n
        return self.predict(np.arange(start, min(start + n, len(self._csv_data))))
