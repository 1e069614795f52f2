"""Per-step tracking metrics on the GPU (K10), same static-method surface as the reference's
ErrorCalculator (wtracker/eval/error_calculator.py:163-195 bbox error, :197-212 MSE error).
float64 in, float64 out; the kernels repeat numpy's IEEE operations one for one, so results are
bit-identical to the reference, NaN rows included."""

from __future__ import annotations

import numpy as np
import torch

from wtracker_b200 import _lib as L


def _run(fn_name: str, worm: np.ndarray | torch.Tensor, mic: np.ndarray | torch.Tensor, device: str):
    if not torch.cuda.is_available():
        raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
    lib = L.lib()
    as_numpy = not torch.is_tensor(worm)
    dev = torch.device(device)
    w = torch.as_tensor(worm, dtype=torch.float64).reshape(-1, 4).to(dev).contiguous()
    m = torch.as_tensor(mic, dtype=torch.float64).reshape(-1, 4).to(dev).contiguous()
    assert w.shape == m.shape, "worm and microscope box tables must have the same length"
    out = torch.empty((w.shape[0],), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(getattr(lib, fn_name)(w.data_ptr(), m.data_ptr(), out.data_ptr(), w.shape[0],
                                      torch.cuda.current_stream().cuda_stream), fn_name)
    return out.cpu().numpy() if as_numpy else out


class ErrorCalculator:
    device = "cuda:0"

    @staticmethod
    def calculate_bbox_error(worm_bboxes, mic_bboxes):
        """Fraction of the worm box lying outside the microscope box; 0 where the worm box is empty."""
        return _run("wt_bbox_error", worm_bboxes, mic_bboxes, ErrorCalculator.device)

    @staticmethod
    def calculate_mse_error(worm_bboxes, mic_bboxes):
        """Mean squared distance between the box centres."""
        return _run("wt_mse_error", worm_bboxes, mic_bboxes, ErrorCalculator.device)
