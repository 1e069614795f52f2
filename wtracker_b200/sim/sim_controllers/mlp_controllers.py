"""Neural position predictor in the loop (reference: wtracker/sim/sim_controllers/mlp_controllers.py).
The float64 glue (gather, relativise, clip, round) stays on the host exactly as in the reference
(:36-68); the network itself runs through the CUDA ResMLP kernel."""

from __future__ import annotations

import numpy as np

from wtracker_b200.neural.engine import ResMLPEngine
from wtracker_b200.neural.mlp import WormPredictor
from wtracker_b200.sim.sim_controllers.csv_controller import CsvController
from wtracker_b200.sim.simulator import Simulator
from wtracker_b200.utils.bbox_utils import BoxUtils


def relativise(boxes28: np.ndarray) -> None:
    """In place: subtract the first box's x (y) from every x (y) column of a (1, 4k) row."""
    x0, y0 = boxes28[0, 0], boxes28[0, 1]
    boxes28[:, 0::4] -= x0
    boxes28[:, 1::4] -= y0


class MLPController(CsvController):
    def __init__(self, timing_config, csv_path, model: WormPredictor, max_speed: float = 0.9, device: str = "cuda:0"):
        super().__init__(timing_config, csv_path)
        self.model = model
        self.io_config = model.io_config
        self.model.eval()
        self._engine = ResMLPEngine(model, device)
        px_per_frame = max_speed * (timing_config.px_per_mm / timing_config.frames_per_sec)
        self.max_dist_per_pred = px_per_frame * self.io_config.pred_frames[0]

    def provide_movement_vector(self, sim: Simulator) -> tuple[int, int]:
        frames = np.asanyarray(self.io_config.input_frames, dtype=int)
        frames = frames + (sim.frame_number - self.timing_config.pred_frame_num)
        cam_center = BoxUtils.center(np.asanyarray(sim.view.camera_position))
        boxes = self.predict(frames, relative=False).reshape(1, -1)
        if not np.isfinite(boxes).all():
            return 0, 0
        rel_x, rel_y = boxes[0, 0] - cam_center[0], boxes[0, 1] - cam_center[1]
        relativise(boxes)
        pred = self._engine.forward_host(boxes.astype(np.float32)).flatten()
        pred = np.clip(pred, -self.max_dist_per_pred, self.max_dist_per_pred)
        return round(pred[0].item() + rel_x), round(pred[1].item() + rel_y)

    def print_model(self):
        print(self.model)
