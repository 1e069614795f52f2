# One GPU round-trip: conv selftests, per-layer times, GPU tests.  Usage: bash tools/gpu_round.sh <tag>
# every step under its own timeout: the whole script stays below 8 minutes even if a kernel hangs
export PYTHONPATH=$PWD
tag=${1:-x}
timeout 240 python tools/gpu_conv_selftest.py > gpurun_out/selftest_$tag.log 2>&1; tail -1 gpurun_out/selftest_$tag.log
grep -v "^\[OK\]" gpurun_out/selftest_$tag.log | head -5
timeout 60 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1 || echo "layer times FAILED/timeout"
head -2 gpurun_out/layers_$tag.log
timeout 150 python -m pytest tests -m gpu -q -x --timeout 60 2>&1 | tail -15 > gpurun_out/tests_$tag.log; tail -3 gpurun_out/tests_$tag.log
