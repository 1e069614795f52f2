"""Runs wt_selftest_conv (tcgen05 conv vs scalar validation conv) over the YOLOv8s layer shapes.

Each case runs in its own process under a timeout so a hung kernel cannot stall the whole sweep.
Usage: python tools/gpu_conv_selftest.py [--quick] [--one b,h,w,cin,cout,k,s,act,res,f32]
"""

from __future__ import annotations

import subprocess
import sys
import time

# cases run with the source slice == whole buffer (verbose bit 1 of wt_selftest_conv): the stride-2 pixel-pair halo kernel
TIGHT_CASES = [
    (2, 32, 32, 32, 64, 3, 2, 1, 0, 0),
    (3, 64, 48, 32, 64, 3, 2, 1, 0, 0),      # ragged rows: 24 output rows in 16-row tiles
    (40, 320, 320, 32, 64, 3, 2, 1, 0, 0),   # layer 1 at depth: ~54 tiles per CTA
    (2, 64, 64, 32, 32, 3, 2, 0, 0, 0),
]

CASES = [
    # batch, h, w, cin, cout, k, stride, act, res, f32
    (2, 16, 16, 64, 64, 1, 1, 0, 0, 0),      # smallest: 1x1, no act
    (2, 16, 16, 64, 64, 1, 1, 1, 0, 0),      # + SiLU
    (2, 16, 16, 64, 64, 3, 1, 1, 0, 0),      # 3x3 taps + zero padding
    (2, 16, 16, 64, 64, 3, 1, 1, 1, 0),      # + residual
    (2, 16, 16, 64, 64, 1, 1, 0, 0, 1),      # f32 output
    (2, 32, 32, 64, 128, 3, 2, 1, 0, 0),     # stride 2 parity views
    (2, 16, 16, 32, 32, 3, 1, 1, 1, 0),      # BK=32 / BN=32 (SWIZZLE_64B everywhere)
    (2, 32, 32, 32, 64, 3, 2, 1, 0, 0),      # layer 1 shape family
    (2, 16, 16, 96, 64, 1, 1, 1, 0, 0),      # cin 96 -> BK 32
    (2, 16, 16, 128, 256, 1, 1, 1, 0, 0),    # BN 256
    (3, 20, 20, 256, 512, 3, 2, 1, 0, 0),    # ragged tiles, 2 N blocks, odd batch
    (4, 20, 20, 256, 256, 3, 1, 1, 1, 0),    # 20x20 patch (4,4,8)
    (2, 40, 40, 128, 128, 3, 1, 1, 0, 0),
    (2, 80, 80, 64, 64, 3, 1, 1, 1, 0),
    (2, 80, 80, 384, 128, 1, 1, 1, 0, 0),
    (2, 20, 20, 1024, 512, 1, 1, 1, 0, 0),
    (2, 80, 80, 64, 64, 1, 1, 0, 0, 1),      # head box logits f32
    (8, 160, 160, 32, 64, 3, 2, 1, 0, 0),    # many tiles per CTA (persistence, phases)
    (16, 80, 80, 128, 128, 3, 1, 1, 0, 0),
    (48, 80, 80, 64, 64, 3, 1, 1, 1, 0),     # resident weights, two MMA issuers, ~16 tiles per CTA, residual
    (40, 160, 160, 32, 32, 3, 1, 1, 1, 0),   # BN 32 resident, ~110 tiles per CTA
    (64, 40, 40, 256, 128, 3, 1, 1, 0, 0),   # 4 halo blocks per tile, streamed weights
    (64, 20, 20, 512, 512, 1, 1, 1, 0, 0),   # many K blocks per tile through a short ring
    # tile groups (NT pixel tiles per weight pass): odd batches leave a phantom tile in the last group
    (3, 40, 40, 128, 128, 3, 1, 1, 0, 0),
    (5, 80, 80, 64, 64, 3, 1, 1, 1, 0),      # resident weights + residual
    (7, 160, 160, 32, 32, 3, 1, 1, 1, 0),    # BN 32 / BK 32 (groups of four)
    (9, 40, 40, 256, 64, 3, 1, 1, 0, 0),
    (33, 80, 80, 128, 128, 3, 1, 1, 1, 0),   # ~11 groups per CTA, streamed weights, residual
    (6, 80, 80, 128, 192, 3, 1, 1, 0, 0),    # BN 192: one tile per pass
    # 40 x 40 maps: 8 x 8 tiles of two interleaved images (odd batches: a phantom second image)
    (5, 40, 40, 128, 128, 3, 1, 1, 1, 0),
    (4, 40, 40, 256, 192, 3, 1, 1, 0, 0),
    (7, 40, 40, 64, 64, 3, 1, 1, 1, 0),
    (64, 40, 40, 128, 128, 3, 1, 1, 1, 0),
    # 1x1 convs with the weights of an N block resident in shared memory
    (64, 40, 40, 512, 256, 1, 1, 1, 0, 0),
    (8, 40, 40, 384, 256, 1, 1, 1, 0, 0),
    (3, 40, 40, 256, 256, 1, 1, 0, 0, 1),    # f32 output
    (5, 80, 80, 256, 128, 1, 1, 1, 0, 0),
]

# chained conv + 1x1 (wt_selftest_conv_chain): batch, h, w, cin, cout, k, stride
CHAIN_CASES = [
    (2, 32, 32, 64, 64, 1, 1),        # generic kernel, one K block each
    (2, 32, 32, 32, 64, 3, 2),        # layer 1 family: stride-2 pixel-pair halo kernel + chain
    (2, 32, 32, 64, 128, 3, 2),       # layer 3 family: generic stride-2 kernel, two staging units
    (3, 64, 48, 32, 64, 3, 2),        # ragged rows
    (2, 32, 32, 64, 64, 3, 1),        # stride-1 halo kernel + chain
    (40, 320, 320, 32, 64, 3, 2),     # layer 1 at depth
    (24, 160, 160, 64, 128, 3, 2),    # layer 3 at depth
]

# concat chain (wt_selftest_conv_cat): batch, h, w
CAT_CASES = [(2, 16, 16), (2, 32, 32), (3, 48, 40), (40, 160, 160)]

CAT_SNIPPET = """
import ctypes, sys
from wtracker_b200._lib import lib
args = [int(v) for v in sys.argv[1].split(',')]
d = ctypes.c_double(-1.0)
rc = lib().wt_selftest_conv_cat(*args, 1, ctypes.byref(d))
if rc != 0:
    print('ERROR', lib().wt_last_error().decode())
    sys.exit(2)
sys.exit(0 if d.value == 0.0 else 3)
"""

CHAIN_SNIPPET = """
import ctypes, sys
from wtracker_b200._lib import lib
args = [int(v) for v in sys.argv[1].split(',')]
d = ctypes.c_double(-1.0)
rc = lib().wt_selftest_conv_chain(*args, 1, ctypes.byref(d))
if rc != 0:
    print('ERROR', lib().wt_last_error().decode())
    sys.exit(2)
sys.exit(0 if d.value == 0.0 else 3)
"""

SNIPPET = """
import ctypes, sys
from wtracker_b200._lib import lib
args = [int(v) for v in sys.argv[1].split(',')]
d = ctypes.c_double(-1.0)
rc = lib().wt_selftest_conv(*args, 1 | (2 if len(sys.argv) > 2 and sys.argv[2] == 'tight' else 0), ctypes.byref(d))
if rc != 0:
    print('ERROR', lib().wt_last_error().decode())
    sys.exit(2)
sys.exit(0 if d.value < 0.07 else 3)
"""


def main() -> int:
    cases = CASES
    if "--quick" in sys.argv:
        cases = CASES[:6]
    if "--one" in sys.argv:
        cases = [tuple(int(v) for v in sys.argv[sys.argv.index("--one") + 1].split(","))]
    failed = 0
    import os
    tight = set() if "--one" in sys.argv else set(TIGHT_CASES)
    if "--one" not in sys.argv and "--quick" not in sys.argv:
        cases = cases + TIGHT_CASES
    if "--chain-only" in sys.argv or "--cat-only" in sys.argv:
        cases = []
    for case in cases:
        arg = ",".join(str(v) for v in case)
        t0 = time.time()
        extra = ["tight"] if case in tight else []
        try:
            res = subprocess.run([sys.executable, "-c", SNIPPET, arg, *extra], capture_output=True, text=True, timeout=int(__import__("os").environ.get("WT_CASE_TIMEOUT", "40")))
            status = {0: "OK", 2: "ERROR", 3: "MISMATCH"}.get(res.returncode, f"rc={res.returncode}")
            out = (res.stdout.strip() + " " + res.stderr.strip()[-400:]).strip()
        except subprocess.TimeoutExpired:
            status, out = "TIMEOUT", ""
        if status != "OK":
            failed += 1
        print(f"[{status}] {arg} ({time.time() - t0:.1f}s) {out}", flush=True)
    n_chain = 0
    if "--one" not in sys.argv:
        chain_cases = [(CHAIN_SNIPPET, c) for c in (CHAIN_CASES[:3] if "--quick" in sys.argv else CHAIN_CASES)]
        chain_cases += [(CAT_SNIPPET, c) for c in (CAT_CASES[:2] if "--quick" in sys.argv else CAT_CASES)]
        if "--cat-only" in sys.argv:
            chain_cases = [(CAT_SNIPPET, c) for c in CAT_CASES]
        for CHAIN_SNIPPET_, case in chain_cases:
            arg = ",".join(str(v) for v in case)
            t0 = time.time()
            n_chain += 1
            try:
                res = subprocess.run([sys.executable, "-c", CHAIN_SNIPPET_, arg], capture_output=True, text=True,
                                     timeout=int(os.environ.get("WT_CASE_TIMEOUT", "40")))
                status = {0: "OK", 2: "ERROR", 3: "MISMATCH"}.get(res.returncode, f"rc={res.returncode}")
                out = (res.stdout.strip() + " " + res.stderr.strip()[-400:]).strip()
            except subprocess.TimeoutExpired:
                status, out = "TIMEOUT", ""
            if status != "OK":
                failed += 1
            print(f"[{status}] chain {arg} ({time.time() - t0:.1f}s) {out}", flush=True)
    cases = list(cases) + [None] * n_chain
    print(f"{len(cases) - failed}/{len(cases)} conv selftests passed")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
