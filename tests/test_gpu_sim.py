"""The plug-in surface on the GPU: MLPController / YoloController inside the Simulator loop."""
import numpy as np
import pytest
import torch

from gpu_common import oracle_model, views_for
from oracle import yolov8_ref as Y
from wtracker_b200 import synth
from wtracker_b200.paths import RESMLP_100, RESMLP_200
from wtracker_b200.sim import ExperimentConfig, SimController, Simulator, TimingConfig
from wtracker_b200.sim.sim_controllers import MLPController, YoloConfig, YoloController
from wtracker_b200.utils.frame_reader import ArrayReader

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("tag,im,path", [("100", 100, RESMLP_100), ("200", 200, RESMLP_200)])
def test_mlp_controller_trace_matches_reference(golden, tag, im, path):
    """Whole-loop golden made by the unmodified reference (torch CPU ResMLP): chosen movement
    vectors and platform positions must be identical integers."""
    from wtracker_b200.neural.mlp import load_worm_predictor

    exp, t = make_timing(1800, im)
    rec = Recorder(MLPController(t, golden["trace_csv_table"], load_worm_predictor(path)))
    Simulator(t, exp, rec).run()
    vec, ref = np.array(rec.vec), golden[f"trace_mlp_{tag}_vec"]
    assert vec.shape == ref.shape
    mism = np.nonzero((vec != ref).any(1))[0]
    assert len(mism) == 0, f"{len(mism)} of {len(ref)} movement vectors differ, first at cycle {mism[:3]}"
    assert np.array_equal(np.array(rec.pos), golden[f"trace_mlp_{tag}_pos"])


def test_yolo_controller_predict_contract():
    cfg = YoloConfig("synthetic:0", pred_kwargs={"imgsz": 384, "conf": 0.1})
    _, t = make_timing(100)
    ctrl = YoloController(t, cfg)
    views = views_for(360, 3)
    out = ctrl.predict(views)
    ref = Y.YoloOracle(oracle_model(), 384, max_det=1).predict(views)
    assert out.shape == ref.shape == (3, 4) and out.dtype == ref.dtype
    assert np.allclose(out, ref, atol=0.5), np.abs(out - ref).max()
    # a frame with nothing above conf -> NaN row and float64 result, as in the reference
    strict = YoloController(t, YoloConfig("synthetic:0", pred_kwargs={"imgsz": 384, "conf": 0.99}))
    out = strict.predict(views[:2])
    assert out.dtype == np.float64 and np.isnan(out).all()
    with pytest.raises(AssertionError):
        ctrl.predict([])
    import pickle
    assert "model" not in pickle.loads(pickle.dumps(cfg)).__dict__


class OracleYoloController(YoloController):
    """Same controller with the CUDA detector swapped for the fp32 oracle (test only)."""

    def __init__(self, timing_config, imgsz):
        SimController.__init__(self, timing_config)
        from collections import deque

        self._camera_frames = deque(maxlen=timing_config.cycle_frame_num)
        self._oracle = Y.YoloOracle(oracle_model(), imgsz, max_det=1)
        self.lazy_views = False      # the oracle takes cropped views, as the reference's controller buffers them

    def predict(self, frames):
        return self._oracle.predict(list(frames))


def test_closed_loop_yolo_controller_matches_oracle_loop():
    """YOLO in the loop: crop k+1 depends on detection k.  The integer host logic is bit-exact
    (CSV / MLP traces above); the detection floats differ from the fp32 oracle by < 0.5 px, which can
    flip round() at a .5 boundary, so movement vectors may differ by one pixel — and because every
    cycle re-centres on the worm the difference must not accumulate."""
    n = 9 * 30
    frames, track = synth.make_frames(n, seed=3, border_visit=False)
    init = (int(track[0, 0]), int(track[0, 1]))
    exp, t = make_timing(n, init=init)
    reader = ArrayReader(frames)
    a = Recorder(YoloController(t, YoloConfig("synthetic:0")))
    Simulator(t, exp, a, reader=reader).run()
    b = Recorder(OracleYoloController(t, 384))
    Simulator(t, exp, b, reader=ArrayReader(frames)).run()
    assert len(a.vec) == len(b.vec) == 30
    dv = np.abs(np.array(a.vec) - np.array(b.vec))
    dp = np.abs(np.array(a.pos) - np.array(b.pos))
    assert dv.max() <= 1 and dp.max() <= 2, (a.vec, b.vec)
    assert (dv == 0).all(axis=1).mean() >= 0.5
    # and it does track: the platform ends within a few pixels of the worm head
    assert np.abs(np.array(a.pos[-1]) - track[-1, :2]).max() < 15


def test_lazy_views_equal_buffered_views():
    """The controller buffers (frame, origin) descriptors and lets the GPU cut the views (frame ingest); with
    ``lazy_views=False`` it buffers cropped views like the reference.  Both must give the same loop, and a descriptor
    must materialise to exactly the view the reference would have buffered — also where the view hangs over the border."""
    from wtracker_b200.sim.sim_controllers.yolo_controller import LazyView

    n = 9 * 6
    frames, track = synth.make_frames(n, seed=5, border_visit=False)
    init = (int(track[0, 0]), int(track[0, 1]))
    exp, t = make_timing(n, init=init)
    runs = []
    for lazy in (True, False):
        rec = Recorder(YoloController(t, YoloConfig("synthetic:0"), lazy_views=lazy))
        Simulator(t, exp, rec, reader=ArrayReader(frames)).run()
        runs.append(rec)
    assert runs[0].vec == runs[1].vec and runs[0].pos == runs[1].pos and len(runs[0].vec) == 6
    for origin in ((-40, 500), (1700, -25), (800, 900), (300, 200)):
        lv = LazyView(frames[3], origin[0], origin[1], 360, 360)
        want = synth.camera_view(frames[3], (origin[0] + 180, origin[1] + 180), 360)
        assert np.array_equal(np.asarray(lv), want) and lv.shape == (360, 360)


def test_cycle_predict_all_matches_oracle():
    """``_cycle_predict_all`` (what LoggingController logs, yolo_controller.py:108-109): all N buffered views of a cycle in
    one detector pass, against the fp32 oracle on the same views."""
    n = 9 * 2
    frames, track = synth.make_frames(n, seed=7, border_visit=False)
    exp, t = make_timing(n, init=(int(track[0, 0]), int(track[0, 1])))

    class Grab(YoloController):
        def on_cycle_end(self, sim):
            self.got = self._cycle_predict_all(sim)
            self.views = [np.asarray(v) for v in self._camera_frames]
            super().on_cycle_end(sim)

    ctrl = Grab(t, YoloConfig("synthetic:0"))
    Simulator(t, exp, ctrl, reader=ArrayReader(frames)).run()
    assert ctrl.got.shape == (9, 4) and len(ctrl.views) == 9
    ref = Y.YoloOracle(oracle_model(), 384, max_det=1).predict(ctrl.views)
    assert ref.dtype == ctrl.got.dtype
    assert np.allclose(ctrl.got, ref, atol=0.5, equal_nan=True), np.nanmax(np.abs(ctrl.got - ref))
    assert np.isfinite(ctrl.got).all()
