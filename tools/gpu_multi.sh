# Multi-GPU trip (gpurun --gpus N): the bench line, the offline job (configs[3]) and the sweep (configs[4]) at N ranks.
# Usage: bash tools/gpu_multi.sh <tag> <N> <offline frames> <experiments per GPU> <frames per experiment>
export PYTHONPATH=$PWD
tag=$1; N=$2; FR=${3:-400000}; EX=${4:-128}; SF=${5:-450}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29511 --steps 20 --warmup 3 --no-extras" > gpurun_out/bench_${tag}_n$N.log 2> gpurun_out/bench_${tag}_n$N.err || { echo "bench N=$N FAILED"; tail -5 gpurun_out/bench_${tag}_n$N.err; }
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${tag}_n$N.log") if l.startswith("{")][-1])
    print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 4), "gather_ms", d["gather_ms"], "numa", d["numa"], "clk", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 bash -c "$(declare -f run); N=$N; run 29512 --workload offline --frames $FR" > gpurun_out/offline_${tag}_n$N.log 2> gpurun_out/offline_${tag}_n$N.err || { echo "offline N=$N FAILED"; tail -5 gpurun_out/offline_${tag}_n$N.err; }
grep "^{" gpurun_out/offline_${tag}_n$N.log | tail -1 | cut -c1-1300
timeout 900 bash -c "$(declare -f run); N=$N; run 29513 --workload sweep --experiments $EX --sim-frames $SF" > gpurun_out/sweep_${tag}_n$N.log 2> gpurun_out/sweep_${tag}_n$N.err || { echo "sweep N=$N FAILED"; tail -5 gpurun_out/sweep_${tag}_n$N.err; }
grep "^{" gpurun_out/sweep_${tag}_n$N.log | tail -1 | cut -c1-1500
