"""Small detector forward through the multi-op launches: per-op output checksums (compare runs with different WT_MEGA_MAX)."""
import sys
import torch
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.detector.weights import synthetic_state_dict

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
imgsz = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = DetectorEngine(synthetic_state_dict(0), (imgsz, imgsz), imgsz, batch=batch, max_det=1)
torch.manual_seed(0)
eng.input_view.random_(0, 255)
for it in range(2):
    eng.forward(batch)
    torch.cuda.synchronize()
print("forward ok", flush=True)
for i, o in enumerate(eng.program.ops):
    t = eng.buffer_tensor(o["dst"], batch).float()
    print(f"op {i:2d} {o['name'][:34]:34s} dst {o['dst']:2d} sum {float(t.double().sum()):+.6e} nan {int(torch.isnan(t).sum())}", flush=True)
