"""Aggregates an ncu launch list (--metrics ... --csv --log-file) by kernel.
Usage: python tools/launch_table.py gpurun_out/launches_X.csv [first_launch_id last_launch_id]  (id range = one step)"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID': hdr = r; start = i + 1; break
ik, im, iv, iu, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 30)
per = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= iv: continue
    lid = int(r[iid])
    if not (lo <= lid <= hi): continue
    v = float(r[iv].replace(',', '') or 0)
    u = r[iu]
    scale = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
    per.setdefault(lid, {'name': r[ik].split('(')[0].replace('void wt::<unnamed>::', '').replace('wt::<unnamed>::', '')})[r[im]] = v * scale
agg = collections.OrderedDict()
for lid, d in per.items():
    a = agg.setdefault(d['name'], dict(n=0, us=0.0, rd=0.0, wr=0.0, tc=0.0))
    a['n'] += 1; a['us'] += d.get('gpu__time_duration.sum', 0); a['rd'] += d.get('dram__bytes_read.sum', 0); a['wr'] += d.get('dram__bytes_write.sum', 0)
    a['tc'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0) * d.get('gpu__time_duration.sum', 0)
tot = sum(a['us'] for a in agg.values())
print(f"# {sys.argv[1]} launches {lo}..{hi if hi < 1 << 30 else 'end'}: {sum(a['n'] for a in agg.values())} launches, {tot:.1f} us of kernel time (ncu: serialised, cold caches — compare shares)")
print(f"{'kernel':48s} {'n':>4} {'us':>9} {'share':>6} {'dramR MB':>9} {'dramW MB':>9} {'TB/s':>6} {'tensor%':>8}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    print(f"{k[:48]:48s} {a['n']:4d} {a['us']:9.1f} {100 * a['us'] / tot:5.1f}% {a['rd']:9.1f} {a['wr']:9.1f} {(a['rd'] + a['wr']) / a['us'] if a['us'] else 0:6.2f} {a['tc'] / a['us'] if a['us'] else 0:8.1f}")
conv = [a for k, a in agg.items() if k.startswith('conv_tc') or k.startswith('conv_halo')]
print(f"# conv kernels: {sum(a['n'] for a in conv)} launches, {sum(a['us'] for a in conv):.1f} us, DRAM {sum(a['rd'] + a['wr'] for a in conv):.1f} MB")
