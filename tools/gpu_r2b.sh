# Round-2 kernel A/B trip: conv selftests, per-layer times for tile-group variants, then the usual tests + bench.
export PYTHONPATH=$PWD
tag=${1:-x}
mkdir -p gpurun_out
timeout 400 python tools/gpu_conv_selftest.py > gpurun_out/selftest_$tag.log 2>&1; tail -1 gpurun_out/selftest_$tag.log
grep -v "^\[OK\]" gpurun_out/selftest_$tag.log | head -8
for v in "1 0" "4 0" "1 1" "4 1"; do
  set -- $v
  WT_CONV_NT=$1 WT_CONV_IL=$2 timeout 90 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_${tag}_nt$1_il$2.log 2>&1 || echo "layer times nt$1 il$2 FAILED"
  echo "nt=$1 il=$2: $(head -1 gpurun_out/layers_${tag}_nt$1_il$2.log)"
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err || { echo "bench FAILED"; tail -5 gpurun_out/bench_$tag.err; }
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 4), "clk", d["clocks"])
    print("roofline", {k: r[k] for k in ("achieved", "frac", "frac_burst", "frac_sustained", "kernel_ms_per_step", "share_of_step")}, "stage", d["stage_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --workload sweep --experiments 128 --sim-frames 450 > gpurun_out/sweep_$tag.log 2> gpurun_out/sweep_$tag.err || { echo "sweep FAILED"; tail -5 gpurun_out/sweep_$tag.err; }
tail -1 gpurun_out/sweep_$tag.log | cut -c1-1500
