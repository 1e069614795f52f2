"""Frame-stepped platform simulator and the controller plug-in surface.

Same contract as the reference (wtracker/sim/simulator.py: Simulator :12-194, SimController
:197-293): per frame the controller hooks fire in the reference's order, the movement vector asked
for at ``cycle_step == imaging_frame_num`` is spread over the moving frames by the motor
controller, and the platform position is clamped to the frame.  Loop state (cycle length, phase
boundaries, controller/motor/view handles) is hoisted out of the per-frame path.
"""

from __future__ import annotations

import abc

import numpy as np

from wtracker_b200.sim.config import ExperimentConfig, TimingConfig
from wtracker_b200.sim.motor_controllers import MotorController, SineMotorController
from wtracker_b200.sim.view_controller import ViewController
from wtracker_b200.utils.frame_reader import DummyReader, FrameReader


class SimController(abc.ABC):
    """Plug-in interface: ten optional hooks + three abstract methods (simulator.py:197-293)."""

    def __init__(self, timing_config: TimingConfig):
        self.timing_config = timing_config

    def on_sim_start(self, sim: "Simulator"):
        pass

    def on_sim_end(self, sim: "Simulator"):
        pass

    def on_cycle_start(self, sim: "Simulator"):
        pass

    def on_cycle_end(self, sim: "Simulator"):
        pass

    def on_camera_frame(self, sim: "Simulator"):
        pass

    def on_imaging_start(self, sim: "Simulator"):
        pass

    def on_micro_frame(self, sim: "Simulator"):
        pass

    def on_imaging_end(self, sim: "Simulator"):
        pass

    def on_movement_start(self, sim: "Simulator"):
        pass

    def on_movement_end(self, sim: "Simulator"):
        pass

    @abc.abstractmethod
    def begin_movement_prediction(self, sim: "Simulator") -> None:
        raise NotImplementedError()

    @abc.abstractmethod
    def provide_movement_vector(self, sim: "Simulator") -> tuple[int, int]:
        """(dx, dy) in pixels by which the platform should move."""
        raise NotImplementedError()

    @abc.abstractmethod
    def _cycle_predict_all(self, sim: "Simulator") -> np.ndarray:
        """(N, 4) xywh worm boxes, camera-relative, one per frame of the finished cycle; NaN row = none."""
        raise NotImplementedError()


class Simulator:
    def __init__(self, timing_config: TimingConfig, experiment_config: ExperimentConfig,
                 sim_controller: SimController, reader: FrameReader = None,
                 motor_controller: MotorController = None) -> None:
        self.timing_config = timing_config
        self.experiment_config = experiment_config
        self._sim_controller = sim_controller
        if reader is None:
            # no pixels needed: constant frames, sized like the reference does (simulator.py:40-44)
            cam = timing_config.camera_size_px
            res = tuple(a + b for a, b in zip(experiment_config.orig_resolution, (cam[0] // 2 * 2, cam[1] // 2 * 2)))
            reader = DummyReader(experiment_config.num_frames, res, colored=True)
        self._motor_controller = motor_controller or SineMotorController(timing_config)
        self._view = ViewController(reader, timing_config.camera_size_px, timing_config.micro_size_px,
                                    experiment_config.init_position)

    # ---- state the controllers read ---------------------------------------------------------
    @property
    def view(self) -> ViewController:
        return self._view

    @property
    def position(self) -> tuple[int, int]:
        return self._view.position

    @property
    def frame_number(self) -> int:
        return self._view.index

    @property
    def cycle_number(self) -> int:
        return self._view.index // self.timing_config.cycle_frame_num

    @property
    def cycle_step(self) -> int:
        return self._view.index % self.timing_config.cycle_frame_num

    def camera_view(self) -> np.ndarray:
        return self._view.camera_view()

    def micro_view(self) -> np.ndarray:
        return self._view.micro_view()

    def _reset(self):
        self._view.reset()
        self._view.set_position(*self.experiment_config.init_position)

    # ---- the loop ---------------------------------------------------------------------------
    def run(self, visualize: bool = False, wait_key: bool = False, progress: bool = False):
        cfg = self.timing_config
        ctrl, motor, view = self._sim_controller, self._motor_controller, self._view
        n_cycle, n_img, n_pred, n_mov = cfg.cycle_frame_num, cfg.imaging_frame_num, cfg.pred_frame_num, cfg.moving_frame_num
        pbar = None
        if progress:
            from tqdm.auto import tqdm

            pbar = tqdm(total=len(view) // n_cycle, desc="Simulation Progress", unit="cycle")

        self._reset()
        ctrl.on_sim_start(self)
        while view.progress():
            step = view.index % n_cycle
            if step == 0:
                if view.index >= n_cycle:
                    ctrl.on_movement_end(self)
                    ctrl.on_cycle_end(self)
                ctrl.on_cycle_start(self)
            ctrl.on_camera_frame(self)
            if step == 0:
                ctrl.on_imaging_start(self)
            if step < n_img:
                ctrl.on_micro_frame(self)
            if step == n_img - n_pred:
                ctrl.begin_movement_prediction(self)
            if step == n_img:
                ctrl.on_imaging_end(self)
                dx, dy = ctrl.provide_movement_vector(self)
                ctrl.on_movement_start(self)
                motor.register_move(dx, dy)
            if n_img <= step < n_img + n_mov:
                view.move_position(*motor.step())
            if pbar is not None and step == n_cycle - 1:
                pbar.update(1)
            if visualize:
                view.visualize_world(timeout=0 if wait_key else 1)
        ctrl.on_sim_end(self)
        if pbar is not None:
            pbar.close()
