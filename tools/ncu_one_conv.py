import ctypes, sys
from wtracker_b200._lib import lib
args = [int(v) for v in sys.argv[1].split(',')]
d = ctypes.c_double(-1.0)
rc = lib().wt_selftest_conv(*args, 1, ctypes.byref(d))
print(rc, d.value)
