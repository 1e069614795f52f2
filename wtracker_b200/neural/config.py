"""I/O description of the position predictor (reference: wtracker/neural/config.py:76-103)."""

from __future__ import annotations

from dataclasses import dataclass, field

from wtracker_b200.utils.config_base import ConfigBase


@dataclass
class IOConfig(ConfigBase):
    """``input_frames`` / ``pred_frames`` are frame offsets relative to the prediction frame (0).
    Every input frame contributes a bbox (x, y, w, h); every predicted frame an (x, y) offset."""

    input_frames: list[int]
    pred_frames: list[int]
    in_dim: int = field(init=False)
    out_dim: int = field(init=False)

    def __post_init__(self):
        if 0 not in self.input_frames:
            print("WARNING::IOConfig::__post_init__::input_frames doesn't contain 0 (the prediction frame). "
                  "Please verify your parameters.")
        self.in_dim = 4 * len(self.input_frames)
        self.out_dim = 2 * len(self.pred_frames)
