"""``DataAnalyzer.load`` / ``initialize`` / ``save`` with the reference's interface
(wtracker/eval/data_analyzer.py:11-107): turns a raw ``bboxes.csv`` log into the analysed table.  The derived columns —
box centres, n-lag speed, deviation, bbox error, the final ``round(5)`` — are one CUDA kernel (``wt_analysis_columns``)
over the whole log; ``analysis_table_device`` is the same kernel for log tables that already live on the GPU.
Cleaning, anomaly reports and plots (the rest of the reference class) are outside the accelerated path.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from wtracker_b200 import _lib as L

ANALYSIS_COLUMNS = ["frame", "cycle", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w",
                    "mic_h", "wrm_x", "wrm_y", "wrm_w", "wrm_h", "time", "cycle_step", "wrm_center_x", "wrm_center_y",
                    "mic_center_x", "mic_center_y", "wrm_speed_x", "wrm_speed_y", "wrm_speed", "worm_deviation_x",
                    "worm_deviation_y", "worm_deviation", "bbox_error", "precise_error"]
_INT_COLUMNS = ["frame", "cycle", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w", "mic_h",
                "time", "cycle_step"]


def analysis_table_device(log_table: torch.Tensor, period: int, cycle_frame_num: int) -> torch.Tensor:
    """f64 [n][17] log rows (``wt_log_rows`` layout) on the device -> f64 [n][30] analysed rows (ANALYSIS_COLUMNS)."""
    if not log_table.is_cuda:
        raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
    t = log_table.to(torch.float64).contiguous()
    assert t.dim() == 2 and t.shape[1] == 17
    out = torch.empty((t.shape[0], 30), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.lib().wt_analysis_columns(t.data_ptr(), t.shape[0], int(period), int(cycle_frame_num), out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "wt_analysis_columns")
    return out


class DataAnalyzer:
    device = "cuda:0"

    def __init__(self, time_config, log_data: pd.DataFrame):
        self.time_config = time_config
        self.data = log_data.copy()
        self._orig_data = log_data
        self._unit = "frame"

    @property
    def unit(self) -> str:
        return self._unit

    def save(self, path: str) -> None:
        self._orig_data.to_csv(path, index=False)

    @staticmethod
    def load(time_config, csv_path: str) -> "DataAnalyzer":
        return DataAnalyzer(time_config, pd.read_csv(csv_path))

    def initialize(self, period: int = 10):
        """Adds the derived columns in the reference's order and rounds every float column to 5 decimals."""
        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        data = self._orig_data
        t = data.drop(columns=["phase"]).to_numpy(dtype=np.float64)
        t17 = np.insert(t, 2, (data["phase"] == "moving").to_numpy(dtype=np.float64), axis=1)
        out = analysis_table_device(torch.from_numpy(np.ascontiguousarray(t17)).to(self.device), period,
                                    self.time_config.cycle_frame_num).cpu().numpy()
        res = pd.DataFrame(out, columns=ANALYSIS_COLUMNS)
        for c in _INT_COLUMNS:
            res[c] = res[c].astype(np.int64)
        res.insert(2, "phase", data["phase"].to_numpy())
        self._orig_data = res
        self.data = res.copy()
