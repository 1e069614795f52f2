"""Platform motion profiles (reference: wtracker/sim/motor_controllers.py).  Host-side float64
arithmetic that decides the next crop coordinates, so it must reproduce the reference's integer
steps exactly: half-cosine split of (dx, dy) over ``moving_frame_num`` frames, each step rounded
with Python's round-half-even and the residual carried into the next step (:70-88)."""

from __future__ import annotations

import abc
from collections import deque

import numpy as np


class MotorController(abc.ABC):
    def __init__(self, timing_config):
        self.timing_config = timing_config
        self.movement_steps = timing_config.moving_frame_num

    @abc.abstractmethod
    def register_move(self, dx: int, dy: int):
        pass

    @abc.abstractmethod
    def step(self) -> tuple[int, int]:
        pass


class StepMotorController(MotorController):
    """Whole displacement in one frame, ``move_after_ratio`` of the way through the movement phase."""

    def __init__(self, timing_config, move_after_ratio: float = 0.5):
        assert 0 <= move_after_ratio <= 1
        super().__init__(timing_config)
        self.queue: list = []
        self.move_at_step = round(self.movement_steps * move_after_ratio)

    def register_move(self, dx: int, dy: int):
        self.queue.extend([(0, 0)] * (self.movement_steps - 1))
        self.queue.insert(self.move_at_step, (dx, dy))

    def step(self) -> tuple[int, int]:
        return self.queue.pop(0)


def sine_fractions(steps: int) -> np.ndarray:
    """Fraction of the move done in step i: (cos(i*pi/n) - cos((i+1)*pi/n)) / 2, float64."""
    i = np.arange(steps)
    return (np.cos(i * np.pi / steps) - np.cos((i + 1) * np.pi / steps)) / 2


class SineMotorController(MotorController):
    def __init__(self, timing_config):
        super().__init__(timing_config)
        self.queue: deque = deque()

    def register_move(self, dx: int, dy: int) -> None:
        assert len(self.queue) == 0
        for i in range(self.movement_steps):
            # evaluated per step with scalar numpy calls, like the reference, to keep identical roundings
            frac = (np.cos((i * np.pi) / self.movement_steps) - np.cos(((i + 1) * np.pi) / self.movement_steps)) / 2
            self.queue.append((frac * dx, frac * dy))

    def step(self) -> tuple[int, int]:
        fx, fy = self.queue.popleft()
        ix, iy = round(fx), round(fy)
        if self.queue:
            nx, ny = self.queue[0]
            self.queue[0] = (nx + (fx - ix), ny + (fy - iy))
        return ix, iy
