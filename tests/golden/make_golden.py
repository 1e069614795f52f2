"""Generates tests/golden/reference_golden.npz by running the UNMODIFIED reference
(/root/reference/wtracker) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

tkinter and ultralytics are absent here; three-line stub modules let everything except the YOLO
call itself import.  The detector has no reference-runnable golden (parity unpinned).
"""
import os
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, "/root/reference")
for name in ("tkinter", "tkinter.filedialog", "seaborn"):   # absent here; only imported, never called
    _m = sys.modules.setdefault(name, types.ModuleType(name))
    _m.Tk = object
sys.modules["tkinter"].filedialog = sys.modules["tkinter.filedialog"]
# wtracker/eval/__init__.py pulls in matplotlib/seaborn viewers; register the package without running it
_pkg = types.ModuleType("wtracker.eval")
_pkg.__path__ = ["/root/reference/wtracker/eval"]
sys.modules["wtracker.eval"] = _pkg
_u = types.ModuleType("ultralytics")
_u.YOLO = object
sys.modules.setdefault("ultralytics", _u)

import pandas as pd  # noqa: E402
import torch  # noqa: E402
from wtracker.eval.error_calculator import ErrorCalculator  # noqa: E402
from wtracker.sim.config import ExperimentConfig, TimingConfig  # noqa: E402
from wtracker.sim.motor_controllers import SineMotorController  # noqa: E402
from wtracker.sim.sim_controllers.csv_controller import CsvController  # noqa: E402
from wtracker.sim.sim_controllers.mlp_controllers import MLPController  # noqa: E402
from wtracker.sim.simulator import Simulator  # noqa: E402
from wtracker.sim.view_controller import ViewController  # noqa: E402
from wtracker.utils.bbox_utils import BoxFormat, BoxUtils  # noqa: E402
from wtracker.utils.frame_reader import DummyReader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
G = {}
rng = np.random.default_rng(1234)

# ---- 1. timing derivations -----------------------------------------------------------------
timing_cases = [(60, 90, 200, 40, 50, 4.0, 0.32), (60, 90, 100, 40, 50, 4.0, 0.32), (30, 57.3, 133, 61, 99, 3.3, 0.5),
                (60, 160, 100, 40, 50, 4.0, 0.32)]
rows = []
for fps, ppm, im, pr, mv, cam, mic in timing_cases:
    exp = ExperimentConfig("g", 1000, fps, (1080, 1920), ppm, (960, 540))
    t = TimingConfig(exp, im, pr, mv, (cam, cam), (mic, mic))
    rows.append([fps, ppm, im, pr, mv, cam, mic, t.imaging_frame_num, t.pred_frame_num, t.moving_frame_num,
                 t.camera_size_px[0], t.micro_size_px[0], t.cycle_frame_num])
G["timing"] = np.array(rows, dtype=np.float64)


# ---- 2./3. whole-loop traces ----------------------------------------------------------------
def synth_csv(n, seed):
    r = np.random.default_rng(seed)
    t = np.arange(n)
    x = 960 + 500 * np.sin(2 * np.pi * t / 2100 + r.uniform(0, 6)) + 150 * np.sin(2 * np.pi * t / 370 + r.uniform(0, 6))
    y = 540 + 300 * np.sin(2 * np.pi * t / 1700 + r.uniform(0, 6)) + 90 * np.sin(2 * np.pi * t / 290 + r.uniform(0, 6))
    w = 14 + r.uniform(-1, 1, n)
    h = 14 + r.uniform(-1, 1, n)
    tab = np.stack([x - w / 2, y - h / 2, w, h], 1)
    tab[r.uniform(size=n) < 0.01] = np.nan      # missed detections
    return tab


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


tab = synth_csv(1800, 7)
G["trace_csv_table"] = tab
tmp = tempfile.mkdtemp()
csv_path = os.path.join(tmp, "bboxes.csv")
pd.DataFrame(tab, columns=["wrm_x", "wrm_y", "wrm_w", "wrm_h"]).to_csv(csv_path, index=False)
for tag, (im, pr, mv) in {"200": (200, 40, 50), "100": (100, 40, 50)}.items():
    exp = ExperimentConfig("g", 1800, 60, (1080, 1920), 90, (960, 540))
    t = TimingConfig(exp, im, pr, mv, (4.0, 4.0), (0.32, 0.32))
    rec = Recorder(CsvController(t, csv_path))
    Simulator(t, exp, rec).run()
    G[f"trace_csv_{tag}_pos"] = np.array(rec.pos, dtype=np.int64)
    G[f"trace_csv_{tag}_vec"] = np.array(rec.vec, dtype=np.int64)
    model = torch.load(f"/root/reference/models/ResMLP(imaging-{tag}ms_pred-40ms_moving-50ms).pt", weights_only=False)
    rec = Recorder(MLPController(t, csv_path, model))
    Simulator(t, exp, rec).run()
    G[f"trace_mlp_{tag}_pos"] = np.array(rec.pos, dtype=np.int64)
    G[f"trace_mlp_{tag}_vec"] = np.array(rec.vec, dtype=np.int64)

# ---- 4. ResMLP known answers ------------------------------------------------------------------
for tag in ("100", "200"):
    model = torch.load(f"/root/reference/models/ResMLP(imaging-{tag}ms_pred-40ms_moving-50ms).pt", weights_only=False)
    model.eval()
    torch.manual_seed(int(tag))
    x = torch.randn(257, 28) * 6
    x[:, 0:2] = 0            # inputs are relative to the first box
    with torch.no_grad():
        y = model(x)
    G[f"resmlp_{tag}_x"] = x.numpy()
    G[f"resmlp_{tag}_y"] = y.numpy()

# ---- 5. metrics ------------------------------------------------------------------------------
n = 600
worm = np.stack([rng.uniform(0, 1900, n), rng.uniform(0, 1000, n), rng.uniform(0, 30, n), rng.uniform(0, 30, n)], 1)
mic = worm + np.stack([rng.normal(0, 12, n), rng.normal(0, 12, n), rng.uniform(0, 20, n), rng.uniform(0, 20, n)], 1)
worm[::37] = np.nan
worm[5::53, 2] = 0.0
worm[9::61, 3] = 0.0
mic[11::71] = np.nan
worm[100:110] = np.round(worm[100:110])
mic[100:110] = np.round(mic[100:110])
G["metric_worm"], G["metric_mic"] = worm.copy(), mic.copy()
with np.errstate(all="ignore"):
    G["metric_bbox_error"] = ErrorCalculator.calculate_bbox_error(worm.copy(), mic.copy())
    G["metric_mse_error"] = ErrorCalculator.calculate_mse_error(worm.copy(), mic.copy())

# ---- 6. bbox utils ---------------------------------------------------------------------------
b = np.stack([rng.uniform(-30, 1950, 200), rng.uniform(-30, 1100, 200), rng.uniform(-2, 40, 200), rng.uniform(-2, 40, 200)], 1)
b[::17] = np.nan
G["disc_in"] = b.copy()
d, legal = BoxUtils.discretize(b.copy(), (1080, 1920), BoxFormat.XYWH)
G["disc_out"], G["disc_legal"] = d, legal
G["center_out"] = BoxUtils.center(np.nan_to_num(G["disc_in"]))
G["round_out"] = BoxUtils.round(np.nan_to_num(G["disc_in"]), BoxFormat.XYWH)


# ---- 7. camera views -------------------------------------------------------------------------
class NoiseReader(DummyReader):
    def __init__(self, frames):
        self._frames_arr = frames
        super().__init__(frames.shape[0], frames.shape[1:], colored=False)

    def __getitem__(self, idx):
        return self._frames_arr[idx]


frames = rng.integers(0, 256, (3, 108, 192), dtype=np.uint8)
G["view_frames"] = frames
positions = [(0, 0), (191, 107), (96, 54), (5, 100), (188, 3), (40, 20), (500, -7)]
G["view_positions"] = np.array(positions)
views_cam, views_mic, campos = [], [], []
vc = ViewController(NoiseReader(frames), camera_size=(36, 36), micro_size=(5, 5), init_position=(96, 54))
for i, p in enumerate(positions):
    vc.seek(i % 3)
    vc.set_position(*p)
    views_cam.append(np.ascontiguousarray(vc.camera_view()))
    views_mic.append(np.ascontiguousarray(vc.micro_view()))
    campos.append(list(vc.camera_position) + list(vc.micro_position) + list(vc.position))
G["view_cam"], G["view_mic"], G["view_boxes"] = np.stack(views_cam), np.stack(views_mic), np.array(campos)

# ---- 8. motor --------------------------------------------------------------------------------
moves = [(7, -3), (0, 0), (-13, 13), (1, 1), (100, -57), (5, 5), (-1, 2), (33, 0)]
steps = []
for n_mov in (3, 4, 7):
    exp = ExperimentConfig("g", 10, 60, (1080, 1920), 90, (0, 0))
    t = TimingConfig(exp, 100, 40, n_mov * 1000 / 60 - 1, (4.0, 4.0), (0.32, 0.32))
    assert t.moving_frame_num == n_mov
    m = SineMotorController(t)
    for dx, dy in moves:
        m.register_move(dx, dy)
        steps.append([n_mov, dx, dy] + [v for _ in range(n_mov) for v in m.step()] + [0] * (2 * (7 - n_mov)))
G["motor_steps"] = np.array(steps, dtype=np.int64)

np.savez_compressed(OUT, **G)
print("wrote", OUT, {k: v.shape for k, v in G.items()})
