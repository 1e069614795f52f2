# Round-2 late check: GPU tests, smoke, plugin timings + bench.  Usage: bash tools/gpu_r2g.sh <tag>
export PYTHONPATH=$PWD
tag=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/tests_$tag.log; tail -6 gpurun_out/tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python tools/gpu_sustained.py 64 640 3 2>&1 | tail -1
timeout 90 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1; head -1 gpurun_out/layers_$tag.log; grep -i "model.9.m" gpurun_out/layers_$tag.log | head
timeout 400 python bench.py > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst']); print(d['plugin']); print(d['stage_ms'])"
