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


class ViewRecorder(Recorder):
    """Recorder that also keeps the pixels of the view each movement vector was decided on."""

    def __init__(self, inner):
        super().__init__(inner)
        self.views, self.det_pos = [], []

    def provide_movement_vector(self, sim):
        pfn = self.inner.timing_config.pred_frame_num
        lv = self.inner._camera_frames[-pfn]
        self.views.append(np.ascontiguousarray(np.asarray(lv)))
        self.det_pos.append(self.pos[-pfn])      # platform position when that view was taken
        return super().provide_movement_vector(sim)


MARGIN, BOX_TOL = 2e-2, 0.5      # as in test_gpu_parity64.py


@torch.no_grad()
def oracle_candidates(view, imgsz=384, conf_thres=0.1):
    """fp32 oracle on one view: (confidence of every anchor, kept anchor or -1, its xyxy box in view pixels, whether the
    decision is AMBIGUOUS: another anchor within MARGIN of the best confidence whose box centre lies more than two pixels
    away — bf16 feature noise, or a one-pixel shift of the crop, may then pick the other box)."""
    x = Y.preprocess([view], imgsz)
    pred = oracle_model()(x)
    conf = pred[0, 4]
    rows, idx = Y.non_max_suppression(pred, conf_thres, 0.7, 1)[0]
    if rows.shape[0] == 0:
        return conf, -1, None, bool(conf.max() > conf_thres - MARGIN)
    best = int(idx[0])
    box = Y.scale_boxes(x.shape[2:], rows[:, :4], view.shape[:2])[0].numpy()
    near = torch.nonzero(conf > conf[best] - MARGIN).flatten()
    d = (pred[0, :2, near] - pred[0, :2, best:best + 1]).abs().amax(0)
    ambiguous = bool((d > 2.0).any()) or bool(abs(float(conf[best]) - conf_thres) < MARGIN)
    return conf, best, box, ambiguous


def test_closed_loop_yolo_controller_matches_oracle_loop():
    """YOLO in the loop over 30 cycles: crop k+1 depends on detection k.  The integer host logic is bit-exact (CSV / MLP
    traces above); the detection floats differ from the fp32 oracle by < 0.5 px, which can flip round() at a .5 boundary.

    (1) Teacher-forced, every cycle: the oracle on the very view the CUDA loop decided on.  Each cycle is IDENTICAL
    (same anchor: box within 0.5 px, movement vector within one pixel), SWAPPED inside the margin (the oracle scores the
    GPU's anchor within 2e-2 of its own best — test_gpu_parity64.py's accounting) or WRONG.  No cycle may be wrong, at most
    10 % swapped; the counts are printed and written to gpurun_out/closed_loop.json.
    (2) Free-running: the same loop with the oracle as the detector.  The two loops see crops that may be a pixel apart
    (a flipped round()), so they are compared in ABSOLUTE coordinates: where each loop places the worm (platform position
    of the detected view + movement vector).  On every cycle whose decision is unambiguous in both loops the two places
    agree within three pixels (one from each round() and one of sub-pixel sensitivity to the shifted crop); at most half of
    the cycles may be ambiguous (the synthetic net scores neighbouring anchors of one worm within 2e-2 of each other); over the whole run the platforms stay within
    a worm box of each other (every cycle re-centres: differences must not accumulate) and both end on the worm."""
    import json
    import os

    n_cycles = 30
    n = 9 * n_cycles
    frames, track = synth.make_frames(n, seed=3, border_visit=False)
    init = (int(track[0, 0]), int(track[0, 1]))
    exp, t = make_timing(n, init=init)
    ctrl = YoloController(t, YoloConfig("synthetic:0"))
    a = ViewRecorder(ctrl)
    Simulator(t, exp, a, reader=ArrayReader(frames)).run()
    b = ViewRecorder(OracleYoloController(t, 384))
    Simulator(t, exp, b, reader=ArrayReader(frames)).run()
    assert len(a.vec) == len(b.vec) == len(a.views) == n_cycles

    # ---- (1) teacher-forced
    eng = ctrl._model.engine((360, 360), 384, 0.1, 0.7, 1, t.cycle_frame_num)
    stats = dict(identical=0, swapped=0, wrong=0, none_both=0, max_box_err=0.0, max_vec_diff=0, worst_margin=0.0)
    failures, amb_a = [], []
    for k, view in enumerate(a.views):
        boxes, counts = eng.detect_views([view])
        conf, o_idx, want, ambiguous = oracle_candidates(view)
        amb_a.append(ambiguous)
        if counts[0] == 0 or o_idx < 0:
            if counts[0] == 0 and o_idx < 0:
                stats["none_both"] += 1
            elif not ambiguous:
                stats["wrong"] += 1
                failures.append(f"cycle {k}: count {counts[0]} vs oracle anchor {o_idx}")
            continue
        got = boxes[0, 0]
        if int(got[5]) == o_idx:
            stats["identical"] += 1
            be = float(np.abs(got[:4] - want).max())
            ovec = (round((want[0] + want[2]) / 2 - 180), round((want[1] + want[3]) / 2 - 180))
            vd = int(np.abs(np.array(ovec) - np.array(a.vec[k])).max())
            stats["max_box_err"], stats["max_vec_diff"] = max(stats["max_box_err"], be), max(stats["max_vec_diff"], vd)
            if be >= BOX_TOL or vd > 1:
                failures.append(f"cycle {k}: same anchor, box off by {be:.3f} px, vector {a.vec[k]} vs {ovec}")
        else:
            margin = float(conf[o_idx] - conf[int(got[5])])
            stats["worst_margin"] = max(stats["worst_margin"], margin)
            if margin < MARGIN:
                stats["swapped"] += 1
            else:
                stats["wrong"] += 1
                failures.append(f"cycle {k}: anchor {int(got[5])} vs oracle {o_idx}, oracle margin {margin:.4f}")

    # ---- (2) free-running
    amb_b = [oracle_candidates(v)[3] for v in b.views]
    amb = np.array(amb_a) | np.array(amb_b)
    va, vb, pa, pb = np.array(a.vec), np.array(b.vec), np.array(a.pos), np.array(b.pos)
    place = np.abs((np.array(a.det_pos) + va) - (np.array(b.det_pos) + vb)).max(axis=1)     # per cycle
    dv, dp = np.abs(va - vb), np.abs(pa - pb)
    stats.update(ambiguous_cycles=int(amb.sum()), place_diff=place.tolist(), ambiguous=amb.astype(int).tolist(),
                 max_place_diff_unambiguous=int(place[~amb].max(initial=0)), max_place_diff=int(place.max()),
                 max_dv=int(dv.max()), max_dp=int(dp.max()), equal_vectors=float((dv == 0).all(axis=1).mean()))
    print("[closed loop] " + json.dumps(stats))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "closed_loop.json"), "w") as f:
        f.write(json.dumps(stats) + "\n" + "\n".join(failures) + "\n")
    assert not failures, failures[:6]
    assert stats["wrong"] == 0 and stats["swapped"] <= 0.1 * n_cycles, stats
    assert stats["identical"] >= 0.8 * n_cycles, stats
    assert stats["max_place_diff_unambiguous"] <= 3 and stats["ambiguous_cycles"] <= n_cycles // 2, (stats, a.vec, b.vec)
    assert stats["max_dp"] <= 16 and stats["equal_vectors"] >= 0.5, (stats, a.vec, b.vec)
    # and both track: the platform ends within a few pixels of the worm head
    for rec in (a, b):
        assert np.abs(np.array(rec.pos[-1]) - track[-1, :2]).max() < 15


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
    # frame ingest at the borders: the window of the frame + the device crop must feed the net exactly the buffered view
    eng = runs[0].inner._model.engine((360, 360), 384, 0.1, 0.7, 1, t.cycle_frame_num)
    origins = [(-40, 500), (1700, -25), (800, 900), (300, 200), (-400, -400), (1919, 1079), (1560, 720), (0, 0)]
    views = [synth.camera_view(frames[i % 6], (x + 180, y + 180), 360) for i, (x, y) in enumerate(origins)]
    for chunk in (slice(0, 1), slice(0, 8), slice(3, 6)):
        fb, fc = eng.detect_frames([frames[i % 6] for i in range(8)][chunk], [o[0] for o in origins][chunk],
                                   [o[1] for o in origins][chunk])
        net_in = eng.input_view[: chunk.stop - chunk.start].cpu().numpy().copy()
        vb, vc = eng.detect_views(views[chunk])
        assert np.array_equal(net_in, eng.input_view[: chunk.stop - chunk.start].cpu().numpy())
        assert np.array_equal(fc, vc) and np.array_equal(fb, vb)


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
