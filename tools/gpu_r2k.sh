# Final pass of round 2: GPU tests, smoke, bench, then the ncu launch list + conv-step capture of the final code.
export PYTHONPATH=$PWD
tag=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/tests_$tag.log; tail -3 gpurun_out/tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err; tail -1 gpurun_out/bench_$tag.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst']); print(d['plugin']['R']); print(d['cpu_baseline'])"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --min-seconds 0"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 400 ncu --metrics $M --clock-control none -c 1300 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:conv_ -s 260 -c 52 --csv --page raw --log-file gpurun_out/conv_step_$tag.csv $CMD > gpurun_out/ncu_step_$tag.log 2>&1
ls -la gpurun_out | grep $tag | head
