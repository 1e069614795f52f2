# Round-2 final pass: GPU tests, SIMT rooflines, bench, then the ncu profile pass.  Usage: bash tools/gpu_r2h.sh <tag>
export PYTHONPATH=$PWD
tag=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
timeout 300 python tools/gpu_simt_roofline.py > gpurun_out/simt_roofline_$tag.txt 2>gpurun_out/simt_roofline_$tag.err; tail -2 gpurun_out/simt_roofline_$tag.err
timeout 400 python bench.py > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst']); print(d['plugin']); print(d['stage_ms'])"
bash tools/gpu_profile_r2.sh $tag 2>&1 | tail -12
