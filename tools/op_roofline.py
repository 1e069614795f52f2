"""Per-op bound vs measured (tools/gpu_layer_times.py log): python tools/op_roofline.py layers.log [batch]"""
import re, sys
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
PEAK_TF, HBM = 1371.6e12, 6543.7e9
tot_m = tot_b = 0
rows = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*(\d+)\s+(\S+)\s+(\d+)->\s*(\d+)\s+k(\d) s(\d) out\s*(\d+)\s+([\d.]+) us", line)
    if not m: continue
    i, name, cin, cout, k, s, ho, us = int(m[1]), m[2], int(m[3]), int(m[4]), int(m[5]), int(m[6]), int(m[7]), float(m[8])
    px = B * ho * ho
    flops = 2 * px * cin * cout * k * k if k != 5 else 0
    in_b = px * s * s * cin * (1 if cin == 1 else 2)
    out_b = px * cout * 2
    t_t = flops / PEAK_TF * 1e6
    # smem-bound UMMA shapes: N=64 -> 48 clk per 32 clk of math, N=32 -> 40 per 16
    n_eff = cout if cout <= 64 else 128
    if k == 3 and cin > 1 and n_eff == 64: t_t *= 1.5
    if k == 3 and cin > 1 and n_eff == 32: t_t *= 2.5
    t_h = (in_b + out_b) / HBM * 1e6
    bound = max(t_t, t_h)
    meas = us - 3.5
    rows.append((meas - bound, i, name, f"{cin}->{cout} k{k}s{s} @{ho}", meas, t_t, t_h))
    tot_m += meas; tot_b += bound
for ex, i, name, shp, meas, t_t, t_h in sorted(rows, reverse=True):
    print(f"{i:3d} {name[:30]:30s} {shp:20s} meas {meas:6.1f}  tensor {t_t:6.1f}  hbm {t_h:6.1f}  excess {ex:6.1f}")
print(f"sum measured {tot_m:.0f} us, sum of bounds {tot_b:.0f} us")
