"""Parallel sweep of simulated experiments (BASELINE configs[4]: 4096 experiments, YOLO + ResMLP controllers,
sharded by experiment over the GPUs of one box — SURVEY.md 8e).

One experiment of the reference is two simulator runs: ``Simulator(LoggingController(YoloController))`` writes the
worm boxes of every frame (workflows/initialize_experiment.ipynb), then ``Simulator(MLPController(bboxes.csv))``
replays the video with the ResMLP predictor in the loop (workflows/simulate.ipynb) and the tracking error is
``ErrorCalculator.calculate_bbox_error`` of the logged worm boxes against the microscope view.  Here the K
experiments of a rank run both passes in lock-step (sim/batched.py): pass 1 keeps its bbox table on the device, pass 2
reads it there, the error is one kernel over all K x F rows.  Ranks take experiment ids ``rank::world``-style
contiguous blocks (sharding.frame_range) and only the per-experiment summary rows are gathered.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200 import synth
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.sim.batched import BatchedMLPController, BatchedSimulator, BatchedYoloController
from wtracker_b200.sim.config import ExperimentConfig, TimingConfig

SUMMARY_COLS = 6   # experiment id, mean bbox error (imaging frames), detected fraction, final x, final y, cycles


def sweep_timing(num_frames: int, px_per_mm: int = 90, imaging_ms: float = 100, frame_hw=(synth.FRAME_H, synth.FRAME_W)):
    """The reference's default experiment geometry (60 fps, 90 px/mm, 4 mm camera -> 360 px view,
    0.32 mm microscope, 100-40-50 ms -> 9-frame cycles)."""
    exp = ExperimentConfig("sweep", num_frames, 60, frame_hw, px_per_mm, (frame_hw[1] // 2, frame_hw[0] // 2))
    return exp, TimingConfig(exp, imaging_ms, 40, 50, (4.0, 4.0), (0.32, 0.32))


def experiment_plan(exp_ids: np.ndarray, n_videos: int, num_frames: int, tracks: list[np.ndarray]):
    """Experiment e replays video e % n_videos from a platform start that is the worm's first position plus a
    deterministic offset of up to +-40 px (so experiments on the same video still differ in every crop)."""
    e = exp_ids.astype(np.int64)
    vid = e % n_videos
    h = (e * 2654435761) & 0xFFFFFFFF
    jx = ((h >> 7) % 81) - 40
    jy = ((h >> 15) % 81) - 40
    start = np.stack([np.rint(np.array([tracks[v][0, 0] for v in vid])).astype(np.int64) + jx,
                      np.rint(np.array([tracks[v][0, 1] for v in vid])).astype(np.int64) + jy], axis=1)
    return vid * num_frames, start


def run_sweep(exp_ids: np.ndarray, num_frames: int, state_dict: dict, predictor, device: str = "cuda:0",
              n_videos: int = 2, imgsz: int = 384, engine_batch: int = 256, seed: int = 0, frames=None, tracks=None):
    """Runs experiments ``exp_ids`` on ``device``.  Returns (summary f64 [K][SUMMARY_COLS] on the device, info dict)."""
    lib = L.lib()
    dev = torch.device(device)
    exp, timing = sweep_timing(num_frames)
    K = int(exp_ids.shape[0])
    with torch.cuda.device(dev):
        t0 = time.perf_counter()
        if frames is None:
            tracks = [synth.worm_track(num_frames, seed + v, border_visit=False) for v in range(n_videos)]
            frames = torch.cat([synth.render_frames_device(tracks[v], seed + v, dev) for v in range(n_videos)])
        torch.cuda.synchronize()
        t_render = time.perf_counter() - t0
        video_base, start = experiment_plan(exp_ids, n_videos, num_frames, tracks)
        cam = timing.camera_size_px
        engine = DetectorEngine(state_dict, (cam[1], cam[0]), imgsz, batch=min(engine_batch, K * timing.cycle_frame_num),
                                max_det=1, device=device)
        launches0 = L.launch_count()
        # ---- pass 1: YOLO in the loop + per-cycle logging of every frame's box
        yolo = BatchedYoloController(timing, engine, frames, video_base, num_frames, log_cycles=True, csv_zero_rows=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r1 = BatchedSimulator(timing, num_frames, start, exp.orig_resolution, yolo).run()
        torch.cuda.synchronize()
        t_pass1 = time.perf_counter() - t0
        # ---- pass 2: ResMLP in the loop over the table pass 1 left on the device
        mlp = BatchedMLPController(timing, yolo.worm_table, predictor)
        t0 = time.perf_counter()
        r2 = BatchedSimulator(timing, num_frames, start, exp.orig_resolution, mlp).run()
        # ---- tracking error of pass 2: logged worm boxes vs the microscope box of every frame
        mw, mh = timing.micro_size_px
        pos = r2["pos_trace"]
        mic = np.empty((num_frames, K, 4), dtype=np.float64)
        mic[..., 0], mic[..., 1], mic[..., 2], mic[..., 3] = pos[..., 0] - mw // 2, pos[..., 1] - mh // 2, mw, mh
        d_mic = torch.from_numpy(mic).to(dev)
        d_err = torch.empty((num_frames, K), dtype=torch.float64, device=dev)
        L.check(lib.wt_bbox_error(yolo.worm_table.data_ptr(), d_mic.data_ptr(), d_err.data_ptr(), num_frames * K,
                                  torch.cuda.current_stream().cuda_stream), "wt_bbox_error")
        err = d_err.cpu().numpy()
        worm = yolo.worm_table.cpu().numpy()
        t_pass2 = time.perf_counter() - t0
        launches = L.launch_count() - launches0
    N, n_img = timing.cycle_frame_num, timing.imaging_frame_num
    logged = (num_frames - 1) // N * N                     # the last cycle is never logged (reference quirk)
    imaging = (np.arange(num_frames) % N < n_img) & (np.arange(num_frames) < logged)
    found = worm[:logged, :, 2] > 0
    summary = np.zeros((K, SUMMARY_COLS))
    summary[:, 0] = exp_ids
    with np.errstate(invalid="ignore"):
        summary[:, 1] = np.nanmean(np.where(found[imaging[:logged]], err[:logged][imaging[:logged]], np.nan), axis=0) \
            if imaging.any() else np.nan
    summary[:, 2] = found.mean(axis=0) if logged else 0.0
    summary[:, 3:5] = pos[-1]
    summary[:, 5] = r2["vec_trace"].shape[0]
    info = dict(experiments=K, frames_per_experiment=num_frames, cycle_frames=N, render_s=t_render, pass1_s=t_pass1,
                pass2_s=t_pass2, detections=int(yolo.det.launched), resmlp_evals=int(mlp.launched), launches=int(launches),
                imgsz=imgsz, view=int(cam[0]), engine_batch=engine.batch, pos_pass1=r1["pos_trace"], pos_pass2=pos,
                vec_pass2=r2["vec_trace"])
    return torch.from_numpy(summary).to(dev), info
