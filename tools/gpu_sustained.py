"""Forward pass of the detector at steady state: runs eng.forward(batch) back to back for ~SECONDS of device time and
prints the mean time per forward and the clocks / power sampled meanwhile (the power-capped regime the bench measures).
Usage: python tools/gpu_sustained.py [batch] [imgsz] [seconds]"""
import statistics
import subprocess
import sys
import threading
import time

import torch

from wtracker_b200.detector.engine import DetectorEngine
from wtracker_b200.detector.weights import synthetic_state_dict

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgsz = int(sys.argv[2]) if len(sys.argv) > 2 else 640
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
eng = DetectorEngine(synthetic_state_dict(0), (imgsz, imgsz), imgsz, batch=batch, max_det=1)
eng.input_view.random_(0, 255)
for _ in range(5):
    eng.forward(batch)
torch.cuda.synchronize()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: rows.extend(l.split(",") for l in proc.stdout), daemon=True).start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < seconds:
    for _ in range(50):
        eng.forward(batch)
    n += 50
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
proc.terminate()
ms = e0.elapsed_time(e1) / n
clk = [float(r[0]) for r in rows[4:] if len(r) == 2]
pw = [float(r[1]) for r in rows[4:] if len(r) == 2]
print(f"sustained forward: {ms:.4f} ms -> {batch / ms * 1e3:.0f} img/s over {n} passes; sm {statistics.median(clk) if clk else 0:.0f} MHz "
      f"(min {min(clk) if clk else 0:.0f}), power {statistics.median(pw) if pw else 0:.0f} W")
