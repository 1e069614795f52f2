"""Host logic of the lock-step engine (wtracker_b200/sim/batched.py) without a GPU: K experiments stepped together
must trace exactly what K separate ``Simulator`` runs trace — motor steps with carried residuals, position clamps,
cycle bookkeeping.  The controller here is a host-only stand-in (boxes from a table, like CsvController); the CUDA
controllers are covered by tests/test_gpu_batched.py."""
import numpy as np

from wtracker_b200.sim import ExperimentConfig, Simulator, TimingConfig
from wtracker_b200.sim.batched import BatchedController, BatchedSimulator
from wtracker_b200.sim.sim_controllers import CsvController


def make_timing(n, im=100, ppm=90, hw=(1080, 1920), init=(960, 540)):
    exp = ExperimentConfig("t", n, 60, hw, ppm, init)
    return exp, TimingConfig(exp, im, 40, 50, (4.0, 4.0), (0.32, 0.32))


class Recorder:
    def __init__(self, inner):
        self.inner, self.pos, self.vec = inner, [], []

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def on_camera_frame(self, sim):
        self.pos.append(tuple(int(v) for v in sim.position))
        return self.inner.on_camera_frame(sim)

    def provide_movement_vector(self, sim):
        v = self.inner.provide_movement_vector(sim)
        self.vec.append((int(v[0]), int(v[1])))
        return v


class BatchedCsv(BatchedController):
    """CsvController.provide_movement_vector (csv_controller.py:51-57) over K experiments: the box of frame
    ``frame_number - pred_frame_num`` relative to the camera view recorded for that frame."""

    def __init__(self, timing, tables):
        self.t, self.tables = timing, tables             # tables: [K][F][4] absolute boxes
        self.cams = {}

    def on_camera_frame(self, sim):
        self.cams[sim.frame_number % self.t.cycle_frame_num] = sim.positions - np.array(self.t.camera_size_px) // 2

    def provide_movement_vectors(self, sim):
        f = sim.frame_number - self.t.pred_frame_num
        K = sim.K
        dx, dy = np.zeros(K, np.int64), np.zeros(K, np.int64)
        cam = self.cams[f % self.t.cycle_frame_num]
        for e in range(K):
            if not (0 <= f < self.tables.shape[1]):
                continue
            b = self.tables[e, f].copy()
            b[0] -= cam[e, 0]
            b[1] -= cam[e, 1]
            if np.isfinite(b).all():
                dx[e] = round(b[0] + b[2] / 2 - self.t.camera_size_px[0] / 2)
                dy[e] = round(b[1] + b[3] / 2 - self.t.camera_size_px[1] / 2)
        return dx, dy


def _tables(K, F, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(F)[None, :, None]
    start = rng.uniform(200, 800, (K, 1, 2))
    vel = rng.uniform(-1.3, 1.3, (K, 1, 2))
    xy = start + vel * t + 25 * np.sin(t / rng.uniform(20, 60, (K, 1, 2)))
    tab = np.concatenate([xy, np.full((K, F, 1), 14.3), np.full((K, F, 1), 12.7)], axis=2)
    tab[rng.uniform(size=(K, F)) < 0.03] = np.nan          # gaps -> (0, 0) vectors
    tab[0, :, 0] -= np.linspace(0, 1500, F)                # one experiment runs into the left border (position clamp)
    return tab


def test_lockstep_equals_separate_simulators():
    K, F = 5, 9 * 40
    for im in (100, 200):
        exp, timing = make_timing(F, im)
        tabs = _tables(K, F, seed=im)
        inits = np.stack([np.rint(tabs[:, 0, 0]).astype(int) + 3, np.rint(tabs[:, 0, 1]).astype(int) - 5], axis=1)
        res = BatchedSimulator(timing, F, inits, exp.orig_resolution, BatchedCsv(timing, tabs)).run()
        for e in range(K):
            exp_e = ExperimentConfig("t", F, 60, (1080, 1920), 90, tuple(int(v) for v in inits[e]))
            rec = Recorder(CsvController(timing, tabs[e]))
            Simulator(timing, exp_e, rec).run()
            assert np.array_equal(res["pos_trace"][:, e], np.array(rec.pos)), (im, e)
            assert np.array_equal(res["vec_trace"][:, e], np.array(rec.vec)), (im, e)
        assert res["vec_trace"].shape == ((F - timing.imaging_frame_num - 1) // timing.cycle_frame_num + 1, K, 2)


def test_sweep_plan_is_deterministic_and_sharded():
    from wtracker_b200.sharding import frame_range
    from wtracker_b200.sweep import experiment_plan, sweep_timing

    exp, timing = sweep_timing(900)
    assert timing.cycle_frame_num == 9 and timing.camera_size_px == (360, 360) and timing.micro_size_px == (29, 29)
    tracks = [np.array([[900.2, 500.7, 0.0]]), np.array([[1000.0, 400.0, 0.0]])]
    ids = np.arange(4096)
    base, start = experiment_plan(ids, 2, 900, tracks)
    assert set(np.unique(base)) == {0, 900}
    assert np.abs(start[ids % 2 == 0] - np.array([900, 501])).max() <= 40
    parts = [experiment_plan(ids[slice(*frame_range(4096, r, 8))], 2, 900, tracks) for r in range(8)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), start)


def test_device_renderer_equals_numpy_on_cpu():
    from wtracker_b200 import synth

    tr = synth.worm_track(700, 3)
    a = synth.render_frames_device(tr, 3, "cpu", first=640, count=2).numpy()
    b = np.stack([synth.render_frame(640 + i, tr, 3) for i in range(2)])
    assert np.array_equal(a, b)
