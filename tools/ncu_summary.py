"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per profiled launch.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
if rep.endswith(".csv"):      # already a raw page (ncu --csv --page raw --log-file): skip ncu's ==PROF== lines
    lines = open(rep).read().splitlines()
    out = "\n".join(lines[next(i for i, l in enumerate(lines) if l.startswith('"ID"')):])
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[col[k]].replace(",", ""))
    except Exception: return d
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised replays: compare shares, not absolutes)")
print(f"{'id':>3} {'kernel':42s} {'grid':>5} {'us':>8} {'dramR_MB':>9} {'dramW_MB':>9} {'dram_TB/s':>9} {'tensor%':>8} {'lts%':>6} {'L2hit%':>7} {'regs':>5}")
tot_us = tot_rd = tot_wr = 0.0
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    name = name.replace("void wt::<unnamed>::", "").replace("(wt::<unnamed>::ConvTcParams)", "").replace("(int)", "")
    us = g(r, "gpu__time_duration.sum")
    u = units[col["gpu__time_duration.sum"]]
    us = us * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    def mb(k):
        v = g(r, k); un = units[col[k]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(un, 1.0)
    rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    tot_us += us; tot_rd += rd; tot_wr += wr
    print(f"{r[col['ID']]:>3} {name[:42]:42s} {int(g(r,'launch__grid_size')):5d} {us:8.1f} {rd:9.1f} {wr:9.1f} {(rd+wr)/us if us else 0:9.2f} "
          f"{g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.1f} {g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r,'lts__t_sector_hit_rate.pct'):7.1f} {int(g(r,'launch__registers_per_thread')):5d}")
print(f"# total: {len(rows) - 2} launches, {tot_us:.1f} us, DRAM read {tot_rd:.1f} MB + write {tot_wr:.1f} MB = {tot_rd + tot_wr:.1f} MB")
if len(sys.argv) > 2:      # write the per-step DRAM traffic for bench.py's roofline.traffic
    import json
    batch, imgsz = int(sys.argv[3]), int(sys.argv[4])
    json.dump({"dram_bytes_per_step": int((tot_rd + tot_wr) * 1e6), "launches": len(rows) - 2, "batch": batch, "imgsz": imgsz,
               "source": rep, "how": "sum of dram__bytes_read.sum + dram__bytes_write.sum over every conv launch of ONE timed step "
               "(tools/gpu_profile_r2.sh: ncu --set full --clock-control none -k regex:conv_ -s 260 -c 52 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --min-seconds 0)"},
              open(sys.argv[2], "w"), indent=1)
