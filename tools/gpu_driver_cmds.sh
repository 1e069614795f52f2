# The driver's own commands at N ranks (both arms), then the offline and sweep workloads.  Usage: bash tools/gpu_driver_cmds.sh <tag> <N> [offline frames] [experiments per GPU] [frames per experiment]
export PYTHONPATH=$PWD
tag=$1; N=$2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 600 $T 29521 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/ref_${tag}_n$N.log 2> gpurun_out/ref_${tag}_n$N.err; echo "reference arm rc=$?"; grep "^{" gpurun_out/ref_${tag}_n$N.log | tail -1 | cut -c1-300
timeout 900 $T 29522 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n$N.log 2> gpurun_out/bench_${tag}_n$N.err; echo "b200 arm rc=$?"; tail -3 gpurun_out/bench_${tag}_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${tag}_n$N.log") if l.startswith("{")][-1])
print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 4), "gather_ms", d["gather_ms"], "numa", d["numa"], "clk", d["clocks"]["sm_mhz"], "frac", d["roofline"]["frac"], "plugin" in d, "library_baseline" in d)
PY
timeout 600 $T 29523 bench.py --gpus $N --workload offline --frames ${3:-400000} > gpurun_out/offline_${tag}_n$N.log 2> gpurun_out/offline_${tag}_n$N.err || { echo "offline FAILED"; tail -5 gpurun_out/offline_${tag}_n$N.err; }
grep "^{" gpurun_out/offline_${tag}_n$N.log | tail -1 | cut -c1-900
timeout 900 $T 29524 bench.py --gpus $N --workload sweep --experiments ${4:-128} --sim-frames ${5:-450} > gpurun_out/sweep_${tag}_n$N.log 2> gpurun_out/sweep_${tag}_n$N.err || { echo "sweep FAILED"; tail -5 gpurun_out/sweep_${tag}_n$N.err; }
grep "^{" gpurun_out/sweep_${tag}_n$N.log | tail -1 | cut -c1-900
