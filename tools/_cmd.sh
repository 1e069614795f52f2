export PYTHONPATH=$PWD
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks','cpu_baseline')})
print(d['e2e']); print(d['roofline'])"
