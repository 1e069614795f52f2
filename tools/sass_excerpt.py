"""SASS evidence for profiles/: per kernel of libwtracker_b200.so the counts of the tcgen05 / TMA / TMEM mnemonics and an
excerpt of the main MMA loop of three representative convolution kernels.
Usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

LIB = "wtracker_b200/_native/libwtracker_b200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
arch = None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None and "/*" in line and ";" in line:
        funcs[cur].append(line)
demangled = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(funcs, demangled)) if len(demangled) == len(funcs) else {k: k for k in funcs}
names = {k: v.replace("(int)", "").replace("(bool)", "").replace("wt::<unnamed>::", "") for k, v in names.items()}
MN = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "UTCATOMSWS", "SYNCS", "MUFU.TANH", "FFMA2"]
print(f"# cuobjdump -sass {LIB}   (arch {arch}); mnemonic counts per kernel")
print(f"{'kernel':86s} {'instr':>6} " + " ".join(f"{m:>10s}" for m in MN))
tot = collections.Counter()
for f, lines in funcs.items():
    n = names[f].replace("(ConvTcParams)", "").replace("void ", "")
    c = {m: sum(1 for l in lines if re.search(r"\b" + re.escape(m), l)) for m in MN}
    for m in MN:
        tot[m] += c[m]
    if c["UTCHMMA"] or c["UTMALDG"] or "post_" in n or "pre_" in n or "hot_tail" in n or "resmlp" in n:
        print(f"{n[:86]:86s} {len(lines):6d} " + " ".join(f"{c[m]:10d}" for m in MN))
print(f"{'TOTAL (all kernels of the library)':86s} {sum(len(l) for l in funcs.values()):6d} " + " ".join(f"{tot[m]:10d}" for m in MN))
two_cta = sum(1 for lines in funcs.values() for l in lines if "2CTA" in l)
print(f"# instructions mentioning 2CTA (cta_group::2): {two_cta}")
for want in ("conv_halo_kernel<128, 64, 0, 2, 2>", "conv_tc_kernel<256, 64>", "conv0_tc_kernel"):
    for f, lines in funcs.items():
        if want in names[f]:
            idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
            if not idx:
                continue
            lo, hi = max(0, idx[0] - 22), min(len(lines), idx[min(len(idx) - 1, 7)] + 6)
            print(f"\n# ---- {names[f]}: instructions {lo}..{hi} (first MMA group of the issue loop)")
            for l in lines[lo:hi]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
            break
