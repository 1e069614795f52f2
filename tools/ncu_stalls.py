"""Per-warp-role stall picture of one profiled kernel: python tools/ncu_stalls.py rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
hdr = rows[hi]
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
seen = set()
for i, r in enumerate(rows[hi + 1:]):
    if len(r) > isamp and r[isamp].isdigit():
        if r[0] in seen: continue      # the page lists the function twice (SASS + source views)
        seen.add(r[0])
        data.append((int(r[isamp]), len(data), r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, i, r in sorted(data, reverse=True)[:topn]:
    st = sorted([(int(r[j]), hdr[j][6:]) for j in stall if r[j] not in ('', '0')], reverse=True)[:3]
    print(f"{i:5d} {n:5d} {100*n/tot:5.1f}% exec={r[iex]:>8s} {r[ia].strip()[:64]:64s} {st}")
if len(sys.argv) > 4:
    lo, hi2 = int(sys.argv[3]), int(sys.argv[4])
    for n, i, r in data:
        if lo <= i <= hi2:
            st = sorted([(int(r[j]), hdr[j][6:]) for j in stall if r[j] not in ('', '0')], reverse=True)[:2]
            print(f"{i:5d} {n:4d} exec={r[iex]:>7s} {r[ia].strip()[:80]:80s} {st}")
