# Round-2 check after the kernel-boundary study: GPU tests on the product library, conv selftests and the event trace on the
# tuning library.  Usage: bash tools/gpu_r2d.sh <tag>
export PYTHONPATH=$PWD
tag=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
export WTRACKER_B200_LIB=$PWD/wtracker_b200/_native/tuning/libwtracker_b200.so
timeout 400 python tools/gpu_conv_selftest.py > gpurun_out/selftest_$tag.log 2>&1
tail -1 gpurun_out/selftest_$tag.log; grep -v "^\[OK\]" gpurun_out/selftest_$tag.log | head -8
timeout 100 python tools/gpu_conv_trace.py $tag 2>&1 | tail -2
timeout 120 python tools/gpu_sustained.py 64 640 3 2>&1 | tail -1
