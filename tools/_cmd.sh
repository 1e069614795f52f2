export PYTHONPATH=$PWD
i=0
for c in 64,80,80,64,64,3,1,1,0,0 64,80,80,128,128,3,1,1,0,0 64,160,160,64,64,1,1,1,0,0 64,320,320,32,64,3,2,1,0,0; do
  i=$((i+1))
  python tools/gpu_conv_selftest.py --one $c > gpurun_out/st_plain_$i.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_halo\|conv_tc -c 1 -o gpurun_out/k$i python -c "
import ctypes, sys
from wtracker_b200._lib import lib
args = [int(v) for v in '$c'.split(',')]
d = ctypes.c_double(-1.0)
rc = lib().wt_selftest_conv(*args, 1, ctypes.byref(d))
print(rc, d.value)
" > gpurun_out/st_ncu_$i.log 2>&1
  tail -2 gpurun_out/st_ncu_$i.log | head -1
done
