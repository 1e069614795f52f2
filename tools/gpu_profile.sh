# Profile pass of one round (run under gpurun; ONE ncu-wrapped program invocation at a time, each after a plain run
# of the same command exited 0).  Outputs stay small (gpurun merges at most 64 MiB back).
#   1. every launch of a short bench run with duration / DRAM bytes / tensor-pipe activity  -> launches_<tag>.csv
#   2. ncu --set full with source for a handful of representative conv launches of a timed step -> conv_<tag>.ncu-rep
export PYTHONPATH=$PWD
tag=${1:-rX}
CONVS_PER_STEP=${2:-55}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 500 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -s $((3 * CONVS_PER_STEP)) -c 8 -o gpurun_out/conv_a_$tag $CMD > gpurun_out/ncu_full_a_$tag.log 2>&1
$CMD > gpurun_out/plain3_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -s $((3 * CONVS_PER_STEP + 46)) -c 4 -o gpurun_out/conv_b_$tag $CMD > gpurun_out/ncu_full_b_$tag.log 2>&1
ls -la gpurun_out/ | tail -12
du -sh gpurun_out
