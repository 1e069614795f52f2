# End-of-round verification on one GPU: every GPU test, smoke, the default bench, the non-conv rooflines, and a fresh
# launch list (metrics pass) of the final code.
export PYTHONPATH=$PWD
tag=${1:-final}
timeout 300 python -m pytest tests -m gpu -q -x --timeout 90 2>&1 | tail -4 > gpurun_out/tests_$tag.log; tail -2 gpurun_out/tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks','stage_ms')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launches_per_step'], d['cpu_baseline']['value'])"
timeout 200 python tools/gpu_simt_roofline.py > gpurun_out/simt_roofline.txt 2>&1; cut -c1-110 gpurun_out/simt_roofline.txt
timeout 60 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1; head -2 gpurun_out/layers_$tag.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 400 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
ls -la gpurun_out/launches_$tag.csv
