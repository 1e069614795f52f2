# Round-2 late check: GPU tests on the product library, smoke, default bench.  Usage: bash tools/gpu_r2e.sh <tag>
export PYTHONPATH=$PWD
tag=$1
timeout 300 python -m pytest tests/test_gpu_sim.py -m gpu -q --timeout 250 -x 2>&1 | tail -25 > gpurun_out/tests_sim_$tag.log; tail -6 gpurun_out/tests_sim_$tag.log
cat gpurun_out/closed_loop.json
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks','cpu_baseline')})
print(d['e2e']); print(d['roofline'])"
