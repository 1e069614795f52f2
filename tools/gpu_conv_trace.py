"""Event trace of every tcgen05 conv launch of ONE forward pass (tuning build, wt_debug_conv_trace): per CTA and warp role
(producer, MMA issuer, the two epilogue leaders) clock64 stamps of the pipeline events.  Writes gpurun_out/trace_<tag>.npz;
tools/conv_trace_report.py turns it into a per-layer table.  Usage: python tools/gpu_conv_trace.py <tag> [batch] [imgsz]"""
import ctypes
import sys

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.detector.weights import synthetic_state_dict

tag = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
imgsz = int(sys.argv[3]) if len(sys.argv) > 3 else 640
CAP, MAXL = 160, 56
eng = DetectorEngine(synthetic_state_dict(0), (imgsz, imgsz), imgsz, batch=batch, max_det=1)
eng.input_view.random_(0, 255)
for _ in range(20):
    eng.forward(batch)
torch.cuda.synchronize()
buf = torch.zeros((MAXL, 148, 4, CAP), dtype=torch.int64, device="cuda")
lib = L.lib()
lib.wt_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    eng.forward(batch)
lib.wt_debug_conv_trace(buf.data_ptr(), CAP, MAXL)
e0.record()
eng.forward(batch)
e1.record()
torch.cuda.synchronize()
lib.wt_debug_conv_trace(None, 0, 0)
names = [o["name"] for o in eng.program.ops if o["kind"] in (0, 1)]
shapes = [f'{o["cin"]}->{o["cout"]} k{o["k"]}s{o["stride"]} @{eng.program.bufs[o["dst"]][0]}' for o in eng.program.ops if o["kind"] in (0, 1)]
np.savez_compressed(f"gpurun_out/trace_{tag}.npz", trace=buf.cpu().numpy().view(np.uint64), names=np.array(names), shapes=np.array(shapes),
                    forward_ms=e0.elapsed_time(e1))
print(f"traced forward: {e0.elapsed_time(e1):.3f} ms, {len(names)} conv launches")
