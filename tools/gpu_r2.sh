# Round-2 GPU round trip: GPU tests, the bench line, small sweep / offline runs.  Usage: bash tools/gpu_r2.sh <tag> [quick]
export PYTHONPATH=$PWD
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
timeout 90 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1 || echo "layer times FAILED/timeout"
head -2 gpurun_out/layers_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err || { echo "bench FAILED"; tail -5 gpurun_out/bench_$tag.err; }
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 4), "R", d["timed_region"], "clk", d["clocks"])
    print("roofline", {k: r[k] for k in ("achieved", "frac", "frac_burst", "frac_sustained", "kernel_ms_per_step", "share_of_step")})
    print("stage", d["stage_ms"], "lib", d.get("library_baseline"), "ingest", d.get("ingest"))
    print("plugin", d.get("plugin")); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
if [ "$2" != "quick" ]; then
timeout 300 python bench.py --workload sweep --experiments 128 --sim-frames 450 > gpurun_out/sweep_$tag.log 2> gpurun_out/sweep_$tag.err || { echo "sweep FAILED"; tail -5 gpurun_out/sweep_$tag.err; }
tail -1 gpurun_out/sweep_$tag.log | cut -c1-1200
timeout 300 python bench.py --workload offline --frames 200000 > gpurun_out/offline_$tag.log 2> gpurun_out/offline_$tag.err || { echo "offline FAILED"; tail -5 gpurun_out/offline_$tag.err; }
tail -1 gpurun_out/offline_$tag.log | cut -c1-1200
fi
