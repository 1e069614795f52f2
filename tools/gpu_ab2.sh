# Selftests of the tuning library under a variant, then same-GPU A/B of several variants at steady state.
# Usage: bash tools/gpu_ab2.sh <tag> "<selftest env>" "<variant 1>" "<variant 2>" ...
export PYTHONPATH=$PWD
export WTRACKER_B200_LIB=$PWD/wtracker_b200/_native/tuning/libwtracker_b200.so
tag=$1; shift
st="$1"; shift
if [ -n "$st" ]; then
  env $st timeout 600 python tools/gpu_conv_selftest.py > gpurun_out/selftest_$tag.log 2>&1
  tail -1 gpurun_out/selftest_$tag.log; grep -v "^\[OK\]" gpurun_out/selftest_$tag.log | head -8
fi
bash tools/gpu_r2c.sh $tag "$@"
