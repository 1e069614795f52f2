"""Per-op timing of the detector program (CUDA events, warm, batch N): python tools/gpu_layer_times.py [batch] [imgsz]"""
import sys
import torch
from wtracker_b200.detector.weights import synthetic_state_dict
from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200 import _lib as L

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgsz = int(sys.argv[2]) if len(sys.argv) > 2 else 640
sd = synthetic_state_dict(0)
eng = DetectorEngine(sd, (imgsz, imgsz), imgsz, batch=batch, max_det=1)
eng.input_view.random_(0, 255)
torch.cuda.synchronize()
ops = eng.program.ops
specs = {s.name: s for s in eng.arch.conv_specs()}
for _ in range(3):
    eng.forward(batch)
torch.cuda.synchronize()
def timeit(fn, iters=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]
total_ms = timeit(lambda: eng.forward(batch))
tot_flop = 2 * eng.arch.macs_per_image(imgsz, imgsz) * batch
print(f"whole forward: {total_ms:.3f} ms  -> {batch / total_ms * 1e3:.0f} img/s, {tot_flop / total_ms / 1e9:.1f} TFLOP/s")
# per-op times INSIDE a whole pass (events between consecutive ops), median over passes: every op sees the
# cache state the real forward gives it (re-running one op back to back would find its input in L2)
import numpy as np
reps = 7
per = np.zeros((reps, len(ops)))
for r in range(reps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
    evs[0].record()
    for i in range(len(ops)):
        eng.forward(batch, i, i + 1)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per[r] = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(ops))]
per = np.median(per, axis=0)
rows = []
for i, o in enumerate(ops):
    ms = float(per[i])
    h, w, c, _ = eng.program.bufs[o["dst"]]
    flop = 2 * h * w * o["cout"] * o["cin"] * o["k"] ** 2 * batch if o["kind"] in (0, 1) else 0
    if o.get("chain_w_off", -1) >= 0 and o["kind"] == 1:      # (kind 0 = layer 0: its chain_w_off is the tcgen05 weight matrix)
        k2 = o["cat_c"] + o["cout"] if o.get("cat_buf", -1) >= 0 else o["cout"]
        n2 = o["chain_cout"] if o.get("cat_buf", -1) >= 0 else o["cout"]
        flop += 2 * h * w * n2 * k2 * batch
    rows.append((ms, i, o["name"], o["cin"], o["cout"], o["k"], o["stride"], h, flop))
s = sum(r[0] for r in rows)
print(f"sum of per-op times: {s:.3f} ms")
for ms, i, name, cin, cout, k, st, h, flop in rows:
    print(f"{i:3d} {name:22s} {cin:5d}->{cout:4d} k{k} s{st} out{h:4d}  {ms*1e3:8.1f} us  {flop/ms/1e9 if flop else 0:7.1f} TF/s  {100*ms/s:5.1f}%")
