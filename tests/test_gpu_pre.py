"""K1-K4 parity: crop (replicate border) + letterbox resize — bit-exact against the oracle
(numpy restatement pinned to cv2.resize and to the reference's ViewController crops)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_common import sample_frames
from oracle import preprocess_ref as P
from oracle import yolov8_ref as Y

pytestmark = pytest.mark.gpu


def run_pre(frames_np, idx, xs, ys, view_hw, imgsz, want_f32=False):
    from wtracker_b200 import _lib as L
    from wtracker_b200.detector.letterbox import letterbox_for, resize_tables

    lib = L.lib()
    lb = letterbox_for(view_hw, imgsz)
    dev = torch.device("cuda:0")
    frames = torch.from_numpy(frames_np).to(dev)
    n = len(idx)
    t_idx = torch.tensor(idx, dtype=torch.int32, device=dev)
    t_x = torch.tensor(xs, dtype=torch.int32, device=dev)
    t_y = torch.tensor(ys, dtype=torch.int32, device=dev)
    tabs = {k: torch.from_numpy(v).to(dev) for k, v in resize_tables(lb).items()} if lb.resample else {}
    ptr = lambda k: tabs[k].data_ptr() if k in tabs else 0  # noqa: E731
    lbc = L.WtLetterbox(lb.src_w, lb.src_h, lb.dst_w, lb.dst_h, lb.new_w, lb.new_h, lb.pad_left, lb.pad_top,
                        ptr("xofs"), ptr("xcoef"), ptr("yofs"), ptr("ycoef"))
    out = torch.zeros((n, lb.dst_h, lb.dst_w), dtype=torch.uint8, device=dev)
    outf = torch.zeros((n, 3, lb.dst_h, lb.dst_w), dtype=torch.float32, device=dev) if want_f32 else None
    L.check(lib.wt_preprocess(frames.data_ptr(), frames.shape[0], frames.shape[1], frames.shape[2], t_idx.data_ptr(),
                              t_x.data_ptr(), t_y.data_ptr(), n, C.byref(lbc), out.data_ptr(),
                              outf.data_ptr() if want_f32 else 0, 0), "wt_preprocess")
    torch.cuda.synchronize()
    return lb, out.cpu().numpy(), (outf.cpu() if want_f32 else None)


CASES = [  # view (h, w), imgsz
    ((640, 640), 640),    # identity (the 640x640 bench geometry)
    ((360, 360), 384),    # reference default: 360 -> 384 up-scale
    ((360, 360), 640),
    ((300, 500), 384),    # non-square: resample + 114 padding
    ((251, 251), 384),    # odd size (the ViewController default)
    ((700, 700), 640),    # down-scale
    ((1000, 1000), 384),  # strong down-scale: the four columns of a work item reach further than two words apart
    ((1080, 1920), 640),  # the whole frame as the view
]


@pytest.mark.parametrize("view_hw,imgsz", CASES)
def test_preprocess_bit_exact(view_hw, imgsz):
    frames, tr = sample_frames(6, 0)
    h, w = view_hw
    H, W = frames.shape[1:]
    # interior, every border, corners, fully outside
    centres = [(int(tr[0, 0]), int(tr[0, 1])), (0, 0), (W - 1, H - 1), (5, H // 2), (W - 3, 7), (W // 2, 2),
               (-400, 100), (W + 300, H + 300)]
    idx = [i % 6 for i in range(len(centres))]
    xs = [c[0] - w // 2 for c in centres]
    ys = [c[1] - h // 2 for c in centres]
    lb, got, _ = run_pre(frames, idx, xs, ys, view_hw, imgsz)
    for i, c in enumerate(centres):
        view = P.crop_replicate(frames[idx[i]], c, (w, h))
        assert view.shape == (h, w)
        want = P.letterbox_u8(view, lb)
        assert np.array_equal(got[i], want), f"case {i}: {np.abs(got[i].astype(int) - want.astype(int)).max()}"


def test_preprocess_f32_tensor_matches_reference_layout():
    """The optional fp32 NCHW output equals what ultralytics would feed the network."""
    frames, tr = sample_frames(6, 0)
    c = (int(tr[1, 0]) + 30, int(tr[1, 1]) - 12)
    lb, _, f32 = run_pre(frames, [1, 2], [c[0] - 180, 100], [c[1] - 180, 900], (360, 360), 384, want_f32=True)
    views = [P.crop_replicate(frames[1], c, (360, 360)), P.crop_replicate(frames[2], (280, 1080), (360, 360))]
    want = Y.preprocess(views, 384)
    assert torch.equal(f32, want)


def test_preprocess_empty_batch_is_a_noop():
    frames, _ = sample_frames(6, 0)
    run_pre(frames, [], [], [], (360, 360), 384)
