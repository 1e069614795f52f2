"""Per-step tracking metrics on the GPU (K10), same static-method surface as the reference's
ErrorCalculator (wtracker/eval/error_calculator.py:163-195 bbox error, :197-212 MSE error, :63-161 precise error).
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


    # ---- segmentation-based error -----------------------------------------------------
    compact_quirk = True
    """The reference writes the results of the legal rows to ``errors[0 .. n_legal)`` (its loop indexes the result
    array with the index into the FILTERED arrays, error_calculator.py:131-159).  True reproduces that array
    exactly; False returns every row's error in its own position."""

    @staticmethod
    def calculate_precise_device(frames: torch.Tensor, frame_idx: torch.Tensor, background: torch.Tensor,
                                 worm_bboxes: torch.Tensor, mic_bboxes: torch.Tensor, diff_thresh: float = 10,
                                 view_off: torch.Tensor | None = None) -> torch.Tensor:
        """Everything on the device: u8 frames [F][H][W] (or, with ``view_off`` i64 [n], a packed buffer of the rows'
        crops), i32 frame index per row, u8 background [H][W], f64 boxes [n][4] -> f64 error per row (NaN = illegal)."""
        if not frames.is_cuda:
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        dev = frames.device
        n = worm_bboxes.shape[0]
        H, W = background.shape[-2:]
        worm = worm_bboxes.to(device=dev, dtype=torch.float64).contiguous()
        mic = mic_bboxes.to(device=dev, dtype=torch.float64).contiguous()
        assert worm.shape == mic.shape == (n, 4) and background.dtype == torch.uint8 and frames.dtype == torch.uint8
        idx = frame_idx.to(device=dev, dtype=torch.int32).contiguous() if frame_idx is not None else None
        off = view_off.to(device=dev, dtype=torch.int64).contiguous() if view_off is not None else None
        out = torch.empty((n,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().wt_precise_error(frames.data_ptr(), frames.shape[0] if view_off is None else 0, H, W,
                                             idx.data_ptr() if idx is not None else None,
                                             off.data_ptr() if off is not None else None,
                                             background.contiguous().data_ptr(), worm.data_ptr(), mic.data_ptr(),
                                             float(diff_thresh), out.data_ptr(), n,
                                             torch.cuda.current_stream().cuda_stream), "wt_precise_error")
        return out

    @staticmethod
    def calculate_precise(background: np.ndarray, worm_bboxes: np.ndarray, mic_bboxes: np.ndarray,
                          frame_nums: np.ndarray, worm_reader, diff_thresh: float = 10) -> np.ndarray:
        """Drop-in for the reference call: ``worm_reader[frame_num]`` is the worm view of a row (the frame cropped at
        its discretized worm box, grey).  Like the reference this discretizes — and thereby zeroes the non-finite
        rows of — the caller's box arrays in place."""
        from wtracker_b200.utils.bbox_utils import BoxFormat, BoxUtils

        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        assert frame_nums.ndim == 1
        assert len(frame_nums) == worm_bboxes.shape[0] == mic_bboxes.shape[0]
        if background.ndim != 2:
            raise ValueError("wtracker_b200 computes the precise error on grey frames")
        bounds = background.shape[:2]
        wb, legal = BoxUtils.discretize(worm_bboxes, bounds=bounds, box_format=BoxFormat.XYWH)
        mb, _ = BoxUtils.discretize(mic_bboxes, bounds=bounds, box_format=BoxFormat.XYWH)
        n = len(frame_nums)
        off = np.zeros(n, dtype=np.int64)
        parts, pos = [], 0
        for i in np.nonzero(legal)[0]:
            view = np.asarray(worm_reader[frame_nums[i]])
            assert view.shape[:2] == (wb[i, 3], wb[i, 2])
            if view.ndim != 2:
                raise ValueError("wtracker_b200 computes the precise error on grey frames")
            off[i] = pos
            parts.append(np.ascontiguousarray(view, dtype=np.uint8).reshape(-1))
            pos += parts[-1].size
        dev = torch.device(ErrorCalculator.device)
        packed = torch.from_numpy(np.concatenate(parts) if parts else np.zeros(1, np.uint8)).to(dev)
        err = ErrorCalculator.calculate_precise_device(packed, None, torch.from_numpy(np.ascontiguousarray(background)).to(dev),
                                                       torch.from_numpy(wb.astype(np.float64)),
                                                       torch.from_numpy(mb.astype(np.float64)), diff_thresh,
                                                       view_off=torch.from_numpy(off)).cpu().numpy()
        if not ErrorCalculator.compact_quirk:
            return err
        out = np.zeros(n, dtype=float)
        out[~legal] = np.nan
        out[:int(legal.sum())] = err[legal]
        return out
