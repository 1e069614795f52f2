"""Timing / experiment dataclasses, field-compatible with the reference (wtracker/sim/config.py:
TimingConfig :10-71, ExperimentConfig :74-129) so saved JSON configs load unchanged."""

from __future__ import annotations

import math
from dataclasses import dataclass, field

from wtracker_b200.utils.config_base import ConfigBase


@dataclass
class ExperimentConfig(ConfigBase):
    name: str
    num_frames: int
    frames_per_sec: float
    orig_resolution: tuple[int, int]   # (h, w)
    px_per_mm: float
    init_position: tuple[int, int]     # platform centre (x, y) in px
    comments: str = ""
    mm_per_px: float = field(init=False)
    ms_per_frame: float = field(init=False)

    def __post_init__(self):
        self.ms_per_frame = 1000 / self.frames_per_sec
        self.mm_per_px = 1 / self.px_per_mm

    @classmethod
    def from_frame_reader(cls, reader, name: str, frames_per_sec: int, px_per_mm: float,
                          init_position: tuple[int, int]) -> "ExperimentConfig":
        return cls(name=name, num_frames=len(reader), frames_per_sec=frames_per_sec,
                   orig_resolution=reader.frame_size, px_per_mm=px_per_mm, init_position=init_position)


@dataclass
class TimingConfig(ConfigBase):
    """ms -> frame counts (ceil) and mm -> px sizes (round); ``experiment_config`` is consumed by
    ``__post_init__`` and dropped, as in the reference (config.py:41-63)."""

    experiment_config: ExperimentConfig = field(repr=False)
    px_per_mm: int = field(init=False)
    mm_per_px: float = field(init=False)
    frames_per_sec: int = field(init=False)
    ms_per_frame: float = field(init=False)
    imaging_time_ms: float
    imaging_frame_num: int = field(init=False)
    pred_time_ms: float
    pred_frame_num: int = field(init=False)
    moving_time_ms: float
    moving_frame_num: int = field(init=False)
    camera_size_mm: tuple[float, float]
    camera_size_px: tuple[int, int] = field(init=False)
    micro_size_mm: tuple[float, float]
    micro_size_px: tuple[int, int] = field(init=False)

    def __post_init__(self):
        exp = self.experiment_config
        self.frames_per_sec, self.ms_per_frame = exp.frames_per_sec, exp.ms_per_frame
        self.mm_per_px, self.px_per_mm = exp.mm_per_px, exp.px_per_mm
        for phase in ("imaging", "pred", "moving"):
            setattr(self, f"{phase}_frame_num", math.ceil(getattr(self, f"{phase}_time_ms") / self.ms_per_frame))
        self.camera_size_px = tuple(round(self.px_per_mm * s) for s in self.camera_size_mm)
        self.micro_size_px = tuple(round(self.px_per_mm * s) for s in self.micro_size_mm)
        del self.experiment_config

    @property
    def cycle_frame_num(self) -> int:
        return self.imaging_frame_num + self.moving_frame_num

    @property
    def cycle_time_ms(self) -> float:
        return self.cycle_frame_num * self.ms_per_frame
