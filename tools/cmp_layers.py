"""Side-by-side per-op times of two tools/gpu_layer_times.py logs: python tools/cmp_layers.py old.log new.log"""
import re, sys
def load(p):
    rows = []
    for line in open(p):
        m = re.match(r"\s*(\d+)\s+(\S+)\s+(\d+)->\s*(\d+)\s+k(\d) s(\d) out\s*(\d+)\s+([\d.]+) us", line)
        if m: rows.append((int(m[1]), m[2], f"{m[3]}->{m[4]} k{m[5]}s{m[6]} @{m[7]}", float(m[8])))
    return rows
a, b = load(sys.argv[1]), load(sys.argv[2])
for (i, n, s, ta), (_, _, _, tb) in zip(a, b):
    print(f"{i:3d} {n[:34]:34s} {s:22s} {ta:7.1f} -> {tb:7.1f}  {tb - ta:+6.1f}")
print("sum", sum(r[3] for r in a), "->", sum(r[3] for r in b))
print(open(sys.argv[1]).readline().strip(), "|", open(sys.argv[2]).readline().strip())
