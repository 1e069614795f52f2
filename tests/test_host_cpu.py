"""Host-side mirror of the reference interface against golden vectors made by the real reference."""
import numpy as np
import pytest

from wtracker_b200.sim import ExperimentConfig, SimController, Simulator, SineMotorController, TimingConfig, ViewController
from wtracker_b200.sim.sim_controllers.csv_controller import CsvController
from wtracker_b200.utils.bbox_utils import BoxConverter, BoxFormat, BoxUtils
from wtracker_b200.utils.frame_reader import ArrayReader, DummyReader


def make_timing(fps=60, ppm=90, im=100, pr=40, mv=50, cam=4.0, mic=0.32, n=1800):
    exp = ExperimentConfig("t", n, fps, (1080, 1920), ppm, (960, 540))
    return exp, TimingConfig(exp, im, pr, mv, (cam, cam), (mic, mic))


def test_timing_config_derivations(golden):
    for row in golden["timing"]:
        fps, ppm, im, pr, mv, cam, mic = row[:7]
        _, t = make_timing(fps, ppm, im, pr, mv, cam, mic)
        got = [t.imaging_frame_num, t.pred_frame_num, t.moving_frame_num, t.camera_size_px[0], t.micro_size_px[0],
               t.cycle_frame_num]
        assert got == [int(v) for v in row[7:]]
    assert not hasattr(t, "experiment_config")


def test_sine_motor_steps(golden):
    for row in golden["motor_steps"]:
        n_mov, dx, dy = int(row[0]), int(row[1]), int(row[2])
        _, t = make_timing(mv=n_mov * 1000 / 60 - 1)
        assert t.moving_frame_num == n_mov
        m = SineMotorController(t)
        m.register_move(dx, dy)
        got = [v for _ in range(n_mov) for v in m.step()]
        assert got == [int(v) for v in row[3:3 + 2 * n_mov]]
        assert sum(got[0::2]) == dx and sum(got[1::2]) == dy


def test_view_controller_crops(golden):
    frames = golden["view_frames"]
    vc = ViewController(ArrayReader(frames), camera_size=(36, 36), micro_size=(5, 5), init_position=(96, 54))
    for i, p in enumerate(golden["view_positions"]):
        vc.seek(i % 3)
        vc.set_position(*p)
        assert np.array_equal(vc.camera_view(), golden["view_cam"][i])
        assert np.array_equal(vc.micro_view(), golden["view_mic"][i])
        boxes = list(vc.camera_position) + list(vc.micro_position) + list(vc.position)
        assert [int(v) for v in boxes] == [int(v) for v in golden["view_boxes"][i]]
        x0, y0 = vc.camera_crop_origin()
        assert (x0, y0) == (int(vc.position[0]) - 18, int(vc.position[1]) - 18)


def test_padded_read_matches_reference_layout():
    frames = np.arange(2 * 6 * 8, dtype=np.uint8).reshape(2, 6, 8)
    vc = ViewController(ArrayReader(frames), camera_size=(4, 4), micro_size=(2, 2), init_position=(3, 3))
    vc.seek(1)
    padded = vc.read()
    assert padded.shape == (10, 12)
    x, y, w, h = vc._calc_view_bbox(4, 4)
    assert np.array_equal(padded[y:y + w, x:x + h], vc.camera_view())


class _Recorder:
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


@pytest.mark.parametrize("tag,im", [("200", 200), ("100", 100)])
def test_simulator_csv_controller_trace(golden, tag, im):
    exp, t = make_timing(im=im)
    rec = _Recorder(CsvController(t, golden["trace_csv_table"]))
    Simulator(t, exp, rec).run()
    assert np.array_equal(np.array(rec.pos), golden[f"trace_csv_{tag}_pos"])
    assert np.array_equal(np.array(rec.vec), golden[f"trace_csv_{tag}_vec"])


def test_hook_order_and_dropped_last_cycle():
    exp, t = make_timing(n=3 * 9 + 4)
    calls = []

    class Spy(SimController):
        def begin_movement_prediction(self, sim):
            calls.append(("pred", sim.frame_number))

        def provide_movement_vector(self, sim):
            calls.append(("vec", sim.frame_number))
            return 3, -2

        def _cycle_predict_all(self, sim):
            return np.zeros((t.cycle_frame_num, 4))

        def on_cycle_end(self, sim):
            calls.append(("end", sim.frame_number))

        def on_cycle_start(self, sim):
            calls.append(("start", sim.frame_number))

    Simulator(t, exp, Spy(t)).run()
    assert t.cycle_frame_num == 9 and t.imaging_frame_num == 6 and t.pred_frame_num == 3
    assert calls[:4] == [("start", 0), ("pred", 3), ("vec", 6), ("end", 9)]
    assert [c for c in calls if c[0] == "end"] == [("end", 9), ("end", 18), ("end", 27)]   # cycle 3 never ends


def test_dummy_reader_default_resolution():
    exp, t = make_timing(n=20)
    sim = Simulator(t, exp, CsvController(t, np.zeros((20, 4))))
    assert sim.view._frame_reader.frame_shape == (1080 + 360, 1920 + 360, 3)
    assert isinstance(sim.view._frame_reader, DummyReader)


def test_bbox_utils(golden):
    b = golden["disc_in"].copy()
    d, legal = BoxUtils.discretize(b, (1080, 1920), BoxFormat.XYWH)
    assert d.dtype == np.int32 and legal.dtype == bool
    assert np.array_equal(d, golden["disc_out"]) and np.array_equal(legal, golden["disc_legal"])
    assert np.all(b[np.isnan(golden["disc_in"]).any(1)] == 0)   # input mutated like the reference
    clean = np.nan_to_num(golden["disc_in"])
    assert np.array_equal(BoxUtils.center(clean), golden["center_out"])
    assert np.array_equal(BoxUtils.round(clean, BoxFormat.XYWH), golden["round_out"])
    xyxy = BoxConverter.to_xyxy(clean, BoxFormat.XYWH)
    assert np.allclose(BoxConverter.to_xywh(xyxy, BoxFormat.XYXY), clean)
    assert np.allclose(BoxConverter.to_yolo(clean, BoxFormat.XYWH)[:, :2], golden["center_out"])
    assert BoxUtils.center(np.array([1.0, 2.0, 4.0, 6.0])).tolist() == [3.0, 5.0]


def test_file_reader_listing_stream_and_batches(tmp_path):
    """FrameReader over image files (reference: utils/frame_reader.py:9-157) + the batched access the ingest path uses."""
    import cv2

    from wtracker_b200.utils.frame_reader import FrameReader, FrameStream

    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, (24, 32), dtype=np.uint8) for _ in range(5)]
    for i, im in enumerate(imgs):
        cv2.imwrite(str(tmp_path / f"frame_{i:04d}.png"), im)
    (tmp_path / "notes").write_text("x")
    (tmp_path / "other.txt").write_text("x")
    (tmp_path / "sub.dir").mkdir()
    r = FrameReader.create_from_template(str(tmp_path), "frame_{:04d}.png")
    assert r.files == [f"frame_{i:04d}.png" for i in range(5)] and len(r) == 5
    assert r.frame_shape == (24, 32) and r.frame_size == (24, 32) and r.read_format == 0
    assert FrameReader.create_from_directory(str(tmp_path)).files == r.files + ["other.txt"]
    assert all(np.array_equal(r[i], imgs[i]) for i in range(5))
    with pytest.raises(IndexError):
        r[5]
    s = r.make_stream()
    assert isinstance(s, FrameStream) and s.index == -1 and not s.can_read()
    assert [int(f.sum()) for f in s] == [int(im.sum()) for im in imgs]
    assert s.seek(2) and np.array_equal(s.read(), imgs[2]) and s.read() is s.read()
    s.reset()
    assert s.index == -1
    assert np.array_equal(r.read_batch([4, 0, 2]), np.stack([imgs[4], imgs[0], imgs[2]]))
    a = ArrayReader(np.stack(imgs))
    assert np.array_equal(a.read_batch(range(1, 4)), np.stack(imgs[1:4])) and a.frame_shape == (24, 32)
    d = DummyReader(3, (10, 12), colored=False)
    assert len(d) == 3 and d.frame_shape == (10, 12) and (d[1] == 255).all()


def test_ingest_window_plus_clamped_crop_equals_camera_view():
    """Frame ingest (DetectorEngine.detect_frames): a view-sized window of the frame travels to the GPU and the crop kernel
    cuts the view out of it with clamp-addressing.  Host statement of that kernel here: for origins inside the frame, over
    every border, over corners and entirely outside, window + clamped crop must give exactly the replicate-border view the
    reference buffers (view_controller.py:45-61,158-172)."""
    from wtracker_b200 import synth
    from wtracker_b200.detector.engine import ingest_window

    rng = np.random.default_rng(0)
    fh, fw = 108, 192
    frame = rng.integers(0, 256, (fh, fw), dtype=np.uint8)
    for size in (36, 64, 108):
        origins = [(10, 20), (-7, 30), (fw - 20, 5), (40, -9), (60, fh - 11), (-15, -15), (fw - 3, fh - 2), (-500, 40),
                   (fw + 300, fh + 300), (0, 0), (fw - size, fh - size)]
        for x0, y0 in origins:
            wx, wy, cx, cy = ingest_window(x0, y0, size, size, fw, fh)
            assert 0 <= wx <= fw - size and 0 <= wy <= fh - size
            window = frame[wy: wy + size, wx: wx + size]
            ys = np.clip(cy + np.arange(size), 0, size - 1)          # the crop kernel's addressing inside the window
            xs = np.clip(cx + np.arange(size), 0, size - 1)
            want = synth.camera_view(frame, (x0 + size // 2, y0 + size // 2), size)
            assert np.array_equal(window[np.ix_(ys, xs)], want), (size, x0, y0)


def test_sine_motor_steps_sum_to_the_requested_move():
    """Property of the residual-carrying half-cosine profile (motor_controllers.py:70-88) that the lock-step engine's
    vectorised motor state relies on: whatever the move, the integer steps of one movement phase add up to it exactly
    (fractions sum to 1 and every rounding error is carried into the next step).  (That the numpy-over-K form of
    ``BatchedSimulator`` takes the same steps as K scalar controllers is test_batched_cpu.py's subject.)"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from wtracker_b200.sim import ExperimentConfig, TimingConfig
    from wtracker_b200.sim.motor_controllers import SineMotorController

    def timing(moving_ms):
        exp = ExperimentConfig("t", 100, 60, (1080, 1920), 90, (960, 540))
        return TimingConfig(exp, 100, moving_ms, 50, (4.0, 4.0), (0.32, 0.32))

    @settings(max_examples=200, deadline=None)
    @given(st.integers(-400, 400), st.integers(-400, 400), st.sampled_from([17, 40, 67, 100, 150]))
    def check(dx, dy, moving_ms):
        t = timing(moving_ms)
        m = SineMotorController(t)
        m.register_move(dx, dy)
        steps = [m.step() for _ in range(t.moving_frame_num)]
        assert all(isinstance(v, int) for s in steps for v in s)
        assert (sum(s[0] for s in steps), sum(s[1] for s in steps)) == (dx, dy)
        assert len(m.queue) == 0

    check()
