"""The lock-step engine on the GPU (BASELINE configs[4]): K experiments through one detector pass per cycle must trace
exactly what K separate ``Simulator + YoloController`` / ``MLPController`` runs of this package trace."""
import numpy as np
import pytest
import torch

from gpu_common import synthetic_sd
from wtracker_b200 import synth
from wtracker_b200.paths import RESMLP_100
from wtracker_b200.sim import ExperimentConfig, Simulator, TimingConfig
from wtracker_b200.sim.batched import BatchedMLPController, BatchedSimulator, BatchedYoloController
from wtracker_b200.sim.sim_controllers import MLPController, YoloConfig, YoloController
from wtracker_b200.utils.frame_reader import ArrayReader

pytestmark = pytest.mark.gpu


def make_timing(n, im=100, ppm=90, hw=(1080, 1920), init=(960, 540)):
    exp = ExperimentConfig("t", n, 60, hw, ppm, init)
    return exp, TimingConfig(exp, im, 40, 50, (4.0, 4.0), (0.32, 0.32))


class Recorder:
    """Records positions, movement vectors and — like LoggingController — the absolute boxes of
    ``_cycle_predict_all`` at every cycle end."""

    def __init__(self, inner, log=False):
        self.inner, self.pos, self.vec, self.log, self.rows, self.cams = inner, [], [], log, {}, []

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def on_camera_frame(self, sim):
        self.pos.append(tuple(int(v) for v in sim.position))
        self.cams.append(sim.view.camera_position)
        return self.inner.on_camera_frame(sim)

    def on_cycle_end(self, sim):
        if self.log:
            n = sim.timing_config.cycle_frame_num
            first = (sim.cycle_number - 1) * n
            rel = np.asarray(self.inner._cycle_predict_all(sim), dtype=np.float64)
            for j in range(rel.shape[0]):
                r = rel[j].copy()
                r[0] += self.cams[first + j][0]
                r[1] += self.cams[first + j][1]
                self.rows[first + j] = r
        return self.inner.on_cycle_end(sim)

    def provide_movement_vector(self, sim):
        v = self.inner.provide_movement_vector(sim)
        self.vec.append((int(v[0]), int(v[1])))
        return v


@pytest.fixture(scope="module")
def video():
    n = 9 * 7
    frames, track = synth.make_frames(n, seed=3, border_visit=False)
    return frames, track


def test_device_renderer_equals_numpy(video):
    frames, track = video
    dev = synth.render_frames_device(track, 3, "cuda:0", first=10, count=3)
    assert np.array_equal(dev.cpu().numpy(), frames[10:13])


@pytest.mark.parametrize("K", [1, 3])
def test_lockstep_yolo_equals_separate_simulators(video, K):
    from wtracker_b200.detector.engine import DetectorEngine

    frames, track = video
    n = frames.shape[0]
    exp, timing = make_timing(n)
    inits = np.array([[int(track[0, 0]) + 11 * e, int(track[0, 1]) - 7 * e] for e in range(K)])
    d_frames = torch.from_numpy(frames).cuda()
    eng = DetectorEngine(synthetic_sd(), (360, 360), 384, batch=8, max_det=1)
    ctrl = BatchedYoloController(timing, eng, d_frames, np.zeros(K, dtype=np.int64), n, log_cycles=True)
    res = BatchedSimulator(timing, n, inits, exp.orig_resolution, ctrl).run()
    torch.cuda.synchronize()
    table = ctrl.worm_table.cpu().numpy()
    for e in range(K):
        exp_e = ExperimentConfig("t", n, 60, (1080, 1920), 90, tuple(int(v) for v in inits[e]))
        rec = Recorder(YoloController(timing, YoloConfig("synthetic:0")), log=True)
        Simulator(timing, exp_e, rec, reader=ArrayReader(frames)).run()
        assert np.array_equal(res["pos_trace"][:, e], np.array(rec.pos)), e
        assert np.array_equal(res["vec_trace"][:, e], np.array(rec.vec)), e
        # every logged frame: identical box (same detector, same crop); the last cycle is never logged
        logged = sorted(rec.rows)
        assert logged == list(range((n // timing.cycle_frame_num - 1) * timing.cycle_frame_num))
        want = np.stack([rec.rows[f] for f in logged])
        assert np.array_equal(table[: len(logged), e], want, equal_nan=True)
        assert np.isnan(table[len(logged):, e]).all()
    assert np.isfinite(table[:9]).all(), "the synthetic worm is detected in the first cycle"
    assert len({tuple(v) for v in res["pos_trace"][0]}) == K      # different starts (they converge on the same worm)


def test_lockstep_mlp_equals_separate_simulators(golden):
    from wtracker_b200.neural.mlp import load_worm_predictor

    K, n = 3, 900
    exp, timing = make_timing(n)
    table = golden["trace_csv_table"][:n]
    inits = np.array([[960, 540], [940, 555], [1010, 500]])
    d_table = torch.from_numpy(np.repeat(table[:, None, :], K, axis=1).copy()).cuda()
    pred = load_worm_predictor(RESMLP_100)
    res = BatchedSimulator(timing, n, inits, exp.orig_resolution, BatchedMLPController(timing, d_table, pred)).run()
    for e in range(K):
        exp_e = ExperimentConfig("t", n, 60, (1080, 1920), 90, tuple(int(v) for v in inits[e]))
        rec = Recorder(MLPController(timing, table, pred))
        Simulator(timing, exp_e, rec).run()
        assert np.array_equal(res["vec_trace"][:, e], np.array(rec.vec)), e
        assert np.array_equal(res["pos_trace"][:, e], np.array(rec.pos)), e
    assert np.abs(res["vec_trace"]).max() > 0


def test_small_sweep_runs_both_passes():
    from wtracker_b200.neural.mlp import load_worm_predictor
    from wtracker_b200.sweep import SUMMARY_COLS, run_sweep

    ids = np.arange(8, 14)
    summary, info = run_sweep(ids, 54, synthetic_sd(), load_worm_predictor(RESMLP_100), n_videos=2, engine_batch=16)
    s = summary.cpu().numpy()
    assert s.shape == (6, SUMMARY_COLS) and np.array_equal(s[:, 0], ids)
    assert info["detections"] == 6 * (6 + 5 * 9)        # one per cycle + every frame of the five logged cycles
    assert info["resmlp_evals"] == 6 * 6
    assert (s[:, 2] > 0.9).all(), "the synthetic worm is found in the logged frames"
    assert np.isfinite(s[:, 1]).all() and ((s[:, 1] >= 0) & (s[:, 1] <= 1)).all()
    assert info["pos_pass2"].shape == (54, 6, 2)
