"""One large resize launch (R geometry, 4096 views) and one batch-64 forward (for the SPPF pool) for ncu:
ncu --set full --clock-control none --import-source on -k regex:"pre_resize|sppf_pool" -c 3 -o gpurun_out/x python tools/ncu_simt.py"""
import ctypes as C

import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.detector.letterbox import letterbox_for, resize_tables
from wtracker_b200.detector.weights import synthetic_state_dict

lib = L.lib()
dev = torch.device("cuda:0")
n, n_frames, view, imgsz = 4096, 200, 360, 384
lb = letterbox_for((view, view), imgsz)
frames = torch.randint(0, 255, (n_frames, 1080, 1920), dtype=torch.uint8, device=dev)
idx = torch.randint(0, n_frames, (n,), dtype=torch.int32, device=dev)
xs = torch.randint(0, 1920 - view, (n,), dtype=torch.int32, device=dev)
ys = torch.randint(0, 1080 - view, (n,), dtype=torch.int32, device=dev)
tabs = {k: torch.from_numpy(v).to(dev) for k, v in resize_tables(lb).items()}
lbc = L.WtLetterbox(lb.src_w, lb.src_h, lb.dst_w, lb.dst_h, lb.new_w, lb.new_h, lb.pad_left, lb.pad_top,
                    tabs["xofs"].data_ptr(), tabs["xcoef"].data_ptr(), tabs["yofs"].data_ptr(), tabs["ycoef"].data_ptr())
out = torch.zeros((n, lb.dst_h, lb.dst_w), dtype=torch.uint8, device=dev)
for _ in range(2):
    L.check(lib.wt_preprocess(frames.data_ptr(), n_frames, 1080, 1920, idx.data_ptr(), xs.data_ptr(), ys.data_ptr(), n,
                              C.byref(lbc), out.data_ptr(), 0, 0), "wt_preprocess")
torch.cuda.synchronize()
eng = DetectorEngine(synthetic_state_dict(0), (640, 640), 640, batch=64, max_det=1)
eng.input_view.random_(0, 255)
eng.forward(64)
torch.cuda.synchronize()
print("ok")
