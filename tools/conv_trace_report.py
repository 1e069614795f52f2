"""Per-launch table of a conv event trace (tools/gpu_conv_trace.py).  Times in microseconds, clock counts converted at
GHZ (default 1.92).  Usage: python tools/conv_trace_report.py gpurun_out/trace_X.npz [GHZ]"""
import sys

import numpy as np

z = np.load(sys.argv[1])
GHZ = float(sys.argv[2]) if len(sys.argv) > 2 else 1.92
tr, names, shapes = z["trace"], list(z["names"]), list(z["shapes"])
M40 = (1 << 40) - 1
GT_BASE = int(min(int(v) for v in tr[:, :, :, 0].ravel() if v != 0))      # float64 cannot hold ns since boot + fractions


def role_events(a):
    """a: [cap] u64 of one role -> (gt0_ns, [(tag, val, t_ns)...]) with t relative to globaltimer"""
    gt0, ck0 = int(a[0]), int(a[1])
    if gt0 == 0:
        return None
    gt0 -= GT_BASE
    ev = []
    for v in a[2:]:
        v = int(v)
        if v == 0:
            break
        clk = v & M40
        d = (clk - (ck0 & M40)) & M40
        ev.append((v >> 56, (v >> 40) & 0xFFFF, gt0 + d / GHZ, clk))
    return gt0, ev


print(f"# {sys.argv[1]}: traced forward {float(z['forward_ms']):.3f} ms; clock counts converted at {GHZ} GHz")
print("# span = first CTA start -> last epilogue end; gap = this launch's first 'dependency resolved' - previous launch's last end;")
print("# dep = CTA start -> dependency resolved (median); fill = dependency resolved -> first operands landed (median);")
print("# per work item of the MMA issuer (median over CTAs of the per-CTA mean, us): item = tempty wait + A wait + B wait + rest;")
print("# pro = kernel entry -> prologue done (median); dep counts from the end of the prologue; relaunch = per SM: CTA entry - last store complete of the previous launch's CTA there (median)")
print("# epi = epilogue per item (tfull -> stores issued); drain = last MMA commit -> CTA's last store complete; spread = last CTA end - median CTA end")
print(f"{'#':>2} {'op':34s} {'shape':20s} {'span':>6} {'gap':>5} {'dep':>5} {'fill':>5} {'items':>5} {'item':>6} {'wT':>5} {'wA':>5} {'wB':>5} {'rest':>5} {'epi':>5} {'drain':>5} {'spread':>6} {'pro':>5} {'relaunch':>8}")
prev_end = None
prev_exits = None
tot_span = tot_gap = 0.0
for li in range(tr.shape[0]):
    if li >= len(names):
        break
    pros, sm_end, sm_start = [], {}, {}
    starts, ends, deps, fills, items, item_t, wT, wA, wB, epi, drain, depres = [], [], [], [], [], [], [], [], [], [], [], []
    for c in range(tr.shape[1]):
        r0, r1 = role_events(tr[li, c, 0]), role_events(tr[li, c, 1])
        r2, r3 = role_events(tr[li, c, 2]), role_events(tr[li, c, 3])
        if r1 is None or r0 is None:
            continue
        gts = [r[0] for r in (r0, r1, r2, r3) if r]
        starts.append(min(gts))
        t_pro = next((t for tag, v, t, _ in r0[1] if tag == 30), None)
        smid = next((v for tag, v, t, _ in r0[1] if tag == 32), None)
        if t_pro is not None:
            pros.append(t_pro - r0[0])
        t_dep = next((t for tag, v, t, _ in r0[1] if tag == 1), None)
        if t_dep is not None:
            deps.append(t_dep - (t_pro if t_pro is not None else min(gts))); depres.append(t_dep)
        ev1 = r1[1]
        first_op = next((t for tag, v, t, _ in ev1 if tag in (12, 13)), None)
        if t_dep is not None and first_op is not None:
            fills.append(first_op - t_dep)
        # per item breakdown
        n_items = sum(1 for e in ev1 if e[0] == 14)
        items.append(n_items)
        if n_items:
            t10 = [e[2] for e in ev1 if e[0] == 10]; t14 = [e[2] for e in ev1 if e[0] == 14]
            item_t.append(np.mean([b - a for a, b in zip(t10, t14)]))
            t11 = [e[2] for e in ev1 if e[0] == 11]
            wT.append(np.mean([b - a for a, b in zip(t10, t11)]))
            wb = sum(e[1] * 16 / GHZ for e in ev1 if e[0] == 13) / n_items
            wB.append(wb)
            # A wait: tag 12 minus the previous event (11 or 13)
            wa = 0.0
            for i, e in enumerate(ev1):
                if e[0] == 12 and i > 0:
                    wa += e[2] - ev1[i - 1][2]
            wA.append(wa / n_items)
        e_end = []
        for r in (r2, r3):
            if not r:
                continue
            t20 = [e[2] for e in r[1] if e[0] == 20]; t21 = [e[2] for e in r[1] if e[0] == 21]
            epi += [b - a for a, b in zip(t20, t21)]
            t22 = [e[2] for e in r[1] if e[0] == 22]
            if t22:
                e_end.append(t22[0])
        if e_end:
            ends.append(max(e_end))
            if smid is not None:
                sm_end[smid] = max(e_end)
            t14 = [e[2] for e in ev1 if e[0] == 14]
            if t14:
                drain.append(max(e_end) - t14[-1])
        if smid is not None:
            sm_start[smid] = r0[0]
    if not starts:
        print(f"{li:2d} {names[li][:34]:34s} {shapes[li]:20s}  (no trace)")
        continue
    s0, e1 = min(starts), (max(ends) if ends else max(starts))
    span = (e1 - s0) / 1e3
    gap = (min(depres) - prev_end) / 1e3 if (prev_end is not None and depres) else 0.0
    prev_end = e1
    tot_span += span; tot_gap += gap
    med = lambda x: (np.median(x) / 1e3 if len(x) else 0.0)
    it = med(item_t); a = med(wT); b = med(wA); c_ = med(wB)
    spread = (max(ends) - np.median(ends)) / 1e3 if ends else 0.0
    rl = [sm_start[k] - prev_exits[k] for k in sm_start if prev_exits and k in prev_exits]
    relaunch = np.median(rl) / 1e3 if rl else 0.0
    prev_exits = sm_end
    print(f"{li:2d} {names[li][:34]:34s} {shapes[li]:20s} {span:6.1f} {gap:5.1f} {med(deps):5.1f} {med(fills):5.1f} {max(items):5d} {it:6.2f} {a:5.2f} {b:5.2f} {c_:5.2f} {it - a - b - c_:5.2f} {med(epi):5.2f} {med(drain):5.2f} {spread:6.1f} {med(pros):5.2f} {relaunch:8.2f}")
print(f"# sum of spans {tot_span:.1f} us, sum of gaps {tot_gap:.1f} us")
