"""``DataAnalyzer.load`` / ``initialize`` / ``save`` with the reference's interface
(wtracker/eval/data_analyzer.py:11-107): turns a raw ``bboxes.csv`` log into the analysed table.  The derived columns —
box centres, n-lag speed, deviation, bbox error, the final ``round(5)`` — are one CUDA kernel (``wt_analysis_columns``)
over the whole log; ``analysis_table_device`` is the same kernel for log tables that already live on the GPU.
``clean`` / ``calc_anomalies`` / ``remove_cycle`` / ``reset_changes`` follow (:109-167, :326-374): their row masks are a
second kernel (``wt_analysis_masks``) over the same table; plots and unit changes (the rest of the reference class) are
outside the accelerated path.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import torch

import ctypes as C

from wtracker_b200 import _lib as L

ANALYSIS_COLUMNS = ["frame", "cycle", "plt_x", "plt_y", "cam_x", "cam_y", "cam_w", "cam_h", "mic_x", "mic_y", "mic_w",
                    "mic_h", "wrm_x", "wrm_y", "wrm_w", "wrm_h", "time", "cycle_step", "wrm_center_x", "wrm_center_y",
                    "mic_center_x", "mic_center_y", "wrm_speed_x", "wrm_speed_y", "wrm_speed", "worm_deviation_x",
                    "worm_deviation_y", "worm_deviation", "bbox_error", "precise_error"]
ANOMALY_FLAGS = ["speed_anomaly", "bbox_error_anomaly", "dist_error_anomaly", "width_anomaly", "height_anomaly",
                 "no_pred_anomaly"]
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

    # ---- row selection (data_analyzer.py:109-167) -------------------------------------------------------------
    def remove_cycle(self, cycles) -> None:
        cycles = [cycles] if isinstance(cycles, (int, np.integer)) else list(cycles)
        self.data = self.data[~self.data["cycle"].isin(cycles)]

    def reset_changes(self) -> None:
        self.data = self._orig_data.copy()
        self._unit = "frame"

    def column_names(self) -> list[str]:
        return self.data.columns.to_list()

    def _masks(self, imaging_only=False, bounds=None, no_preds=True, min_bbox_error=np.inf, min_dist_error=np.inf,
               min_speed=np.inf, min_size=np.inf) -> tuple[np.ndarray, np.ndarray]:
        """(keep bool [n], anomaly bits u8 [n]) of the rows of ``self.data`` from ``wt_analysis_masks``."""
        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        data = self.data
        n = len(data)
        if n == 0:
            return np.zeros(0, dtype=bool), np.zeros(0, dtype=np.uint8)
        missing = [c for c in ANALYSIS_COLUMNS if c not in data.columns]
        if missing:
            raise KeyError(f"initialize() has not been run: missing columns {missing[:4]}")
        dev = torch.device(self.device)
        table = torch.from_numpy(np.ascontiguousarray(data[ANALYSIS_COLUMNS].to_numpy(dtype=np.float64))).to(dev)
        moving = torch.from_numpy((data["phase"] == "moving").to_numpy(dtype=np.uint8)).to(dev)
        keep = torch.empty(n, dtype=torch.uint8, device=dev)
        anom = torch.empty(n, dtype=torch.uint8, device=dev)
        b = (C.c_double * 4)(*[float(v) for v in bounds]) if bounds is not None else None
        with torch.cuda.device(dev):
            L.check(L.lib().wt_analysis_masks(table.data_ptr(), moving.data_ptr(), n, int(bool(imaging_only)), b,
                                              int(bool(no_preds)), float(min_bbox_error), float(min_dist_error),
                                              float(min_speed), float(min_size), keep.data_ptr(), anom.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "wt_analysis_masks")
        return keep.cpu().numpy().astype(bool), anom.cpu().numpy()

    def clean(self, trim_cycles: bool = False, imaging_only: bool = False,
              bounds: tuple[float, float, float, float] = None) -> None:
        """Keeps imaging frames only / frames whose worm box (microscope box when there is no prediction) lies inside
        ``bounds`` / drops the first and the last remaining cycle, in that order (:121-159)."""
        if imaging_only or bounds is not None:
            keep, _ = self._masks(imaging_only=imaging_only, bounds=bounds)
            self.data = self.data[keep]
        if trim_cycles:
            cyc = self.data["cycle"]
            self.data = self.data[(cyc != 0) & (cyc != cyc.max())]

    def calc_anomalies(self, no_preds: bool = True, min_bbox_error: float = np.inf, min_dist_error: float = np.inf,
                       min_speed: float = np.inf, min_size: float = np.inf, remove_anomalies: bool = False) -> pd.DataFrame:
        """Rows that cross any threshold, with one boolean column per criterion (:326-374)."""
        _, bits = self._masks(no_preds=no_preds, min_bbox_error=min_bbox_error, min_dist_error=min_dist_error,
                              min_speed=min_speed, min_size=min_size)
        mask = bits != 0
        anomalies = self.data[mask].copy()
        for i, name in enumerate(ANOMALY_FLAGS):
            anomalies[name] = ((bits[mask] >> i) & 1).astype(bool)
        if remove_anomalies:
            self.data = self.data[~mask]
        return anomalies
