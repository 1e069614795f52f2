export PYTHONPATH=$PWD
bash tools/gpu_round.sh u22
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_u22.log 2>gpurun_out/bench_u22.err; tail -3 gpurun_out/bench_u22.err
python -c "import json; d=json.loads(open('gpurun_out/bench_u22.log').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'], d['gpu_launches'], d['clocks'])"
