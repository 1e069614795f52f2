"""HBM / FMA roofline of the non-conv kernels at sizes large enough to leave the launch-latency regime
(the bench batch of 64 frames moves only ~50 MB through them).  Algorithmic bytes per unit are SURVEY.md §8(d)'s;
times are CUDA events around `reps` launches on fresh, larger-than-L2 data.

    python tools/gpu_simt_roofline.py > gpurun_out/simt_roofline.txt
"""
import ctypes as C
import json
import sys

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.letterbox import letterbox_for, resize_tables

lib = L.lib()
dev = torch.device("cuda:0")
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))
except Exception:
    PEAK = {}
HBM = float(PEAK.get("hbm_gbs", 6543.7))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def report(name, ms, nbytes, note=""):
    gbs = nbytes / ms / 1e6
    print(f"{name:34s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.1f} GB/s  {100*gbs/HBM:5.1f}% of {HBM:.0f} GB/s  {note}")


def pre_case(view, imgsz, n, n_frames=200):
    lb = letterbox_for((view, view), imgsz)
    frames = torch.randint(0, 255, (n_frames, 1080, 1920), dtype=torch.uint8, device=dev)    # 415 MB: larger than L2
    g = torch.Generator(device="cpu").manual_seed(0)
    idx = torch.randint(0, n_frames, (n,), generator=g, dtype=torch.int32).to(dev)
    xs = torch.randint(0, 1920 - view, (n,), generator=g, dtype=torch.int32).to(dev)
    ys = torch.randint(0, 1080 - view, (n,), generator=g, dtype=torch.int32).to(dev)
    tabs = {k: torch.from_numpy(v).to(dev) for k, v in resize_tables(lb).items()} if lb.resample else {}
    ptr = lambda k: tabs[k].data_ptr() if k in tabs else 0  # noqa: E731
    lbc = L.WtLetterbox(lb.src_w, lb.src_h, lb.dst_w, lb.dst_h, lb.new_w, lb.new_h, lb.pad_left, lb.pad_top,
                        ptr("xofs"), ptr("xcoef"), ptr("yofs"), ptr("ycoef"))
    out = torch.empty((n, lb.dst_h, lb.dst_w), dtype=torch.uint8, device=dev)
    def run():
        L.check(lib.wt_preprocess(frames.data_ptr(), n_frames, 1080, 1920, idx.data_ptr(), xs.data_ptr(), ys.data_ptr(), n,
                                  C.byref(lbc), out.data_ptr(), 0, 0), "wt_preprocess")
    ms = timed(run)
    report(f"pre_kernel {view}->{imgsz} x{n}", ms, n * (view * view + lb.dst_h * lb.dst_w),
           "read c^2 u8 crop + write S^2 u8 grey (C=1 folded)")


def metrics_case(n):
    worm = torch.rand((n, 4), dtype=torch.float64, device=dev) * 100
    mic = torch.rand((n, 4), dtype=torch.float64, device=dev) * 100
    err = torch.empty((n,), dtype=torch.float64, device=dev)
    for fn in ("wt_bbox_error", "wt_mse_error"):
        ms = timed(lambda: L.check(getattr(lib, fn)(worm.data_ptr(), mic.data_ptr(), err.data_ptr(), n, 0), fn))
        report(f"{fn[3:]}_kernel x{n}", ms, n * 72, "8 f64 in + 1 f64 out per row")


def resmlp_case(n):
    from wtracker_b200.neural.engine import ResMLPEngine
    from wtracker_b200.paths import RESMLP_100
    from wtracker_b200.neural.mlp import load_worm_predictor
    model = load_worm_predictor(RESMLP_100)
    eng = ResMLPEngine(model)
    x = torch.randn((n, 28), dtype=torch.float32, device=dev)
    out = torch.empty((n, 2), dtype=torch.float32, device=dev)
    ms = timed(lambda: eng.forward(x, out))
    flops = 9440 * n
    print(f"{'resmlp_kernel x%d' % n:34s} {ms*1e3:9.1f} us  {n*120/1e6:9.1f} MB  {n*120/ms/1e6:8.1f} GB/s  "
          f"{flops/ms/1e9:7.2f} TFLOP/s fp32 FMA (9,440 FLOP + 120 B per sample)")


def post_case(n, net=640):
    hw = [(net // s, net // s) for s in (8, 16, 32)]
    keep, total = [], 0
    lv = (L.WtHeadLevel * 3)()
    g = torch.Generator(device="cpu").manual_seed(1)
    for i, (h, w) in enumerate(hw):
        box = (torch.randn((n, h * w, 64), generator=g) * 2).to(dev)
        logit = (torch.randn((n, h * w), generator=g) * 1.2 - 4.5).to(dev)        # ~0.3 % of anchors above conf 0.1
        keep += [box, logit]
        lv[i] = L.WtHeadLevel(box.data_ptr(), None, logit.data_ptr(), h, w, (8, 16, 32)[i], L.WT_DT_F32, 0, None, 0.0)
        total += h * w
    pp = L.WtPostParams(0.1, 0.7, 1, net, net, net, net, 1.0, 0.0, 0.0)
    out = torch.zeros((n, 1, 6), dtype=torch.float32, device=dev)
    cnt = torch.zeros((n,), dtype=torch.int32, device=dev)
    scratch = torch.zeros(lib.wt_post_scratch_bytes(n, total), dtype=torch.uint8, device=dev)
    ms = timed(lambda: L.check(lib.wt_decode_nms(lv, 3, n, C.byref(pp), out.data_ptr(), cnt.data_ptr(), scratch.data_ptr(), 0),
                               "wt_decode_nms"))
    surv = 0.003 * total
    report(f"post_kernel {net}^2 max_det=1 x{n}", ms, n * (total * 4 + surv * 64 * 4),
           "read A f32 logits + 64 f32 box logits per surviving anchor (conf-first lower bound)")


if __name__ == "__main__":
    print(f"# non-conv kernels against the measured HBM peak ({HBM:.0f} GB/s, MEASURED_PEAKS.json); CUDA events, median of 5")
    pre_case(640, 640, 64)
    pre_case(640, 640, 2048)
    pre_case(360, 384, 4096)
    pre_case(360, 640, 2048)
    post_case(64)
    post_case(2048)
    metrics_case(64)
    metrics_case(1 << 24)
    resmlp_case(64)
    resmlp_case(1 << 20)
