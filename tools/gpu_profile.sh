# Profile pass of one round (run under gpurun; ONE ncu-wrapped program invocation at a time, each after a plain run
# of the same command exited 0).  Outputs stay small (gpurun merges at most 64 MiB back).
#   1. every launch of a short bench run with duration / DRAM bytes / tensor-pipe activity  -> launches_<tag>.csv
#   2. ncu --set full of EVERY conv launch of one timed step, raw page as csv (DRAM traffic of the step) -> conv_step_<tag>.csv
#   3. ncu --set full with source for a handful of representative conv launches of a timed step -> conv_{a,b}_<tag>.ncu-rep
#   4. ncu --set full with source for the non-conv kernels of a timed step -> simt_<tag>.ncu-rep
export PYTHONPATH=$PWD
tag=${1:-rX}
CONVS_PER_STEP=${2:-52}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 500 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
timeout 400 ncu --set full --clock-control none -k regex:conv_ -s $((3 * CONVS_PER_STEP)) -c $CONVS_PER_STEP --csv --page raw --log-file gpurun_out/conv_step_$tag.csv $CMD > gpurun_out/ncu_step_$tag.log 2>&1
$CMD > gpurun_out/plain3_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -s $((3 * CONVS_PER_STEP)) -c 8 -o gpurun_out/conv_a_$tag $CMD > gpurun_out/ncu_full_a_$tag.log 2>&1
$CMD > gpurun_out/plain4_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -s $((3 * CONVS_PER_STEP + 36)) -c 6 -o gpurun_out/conv_b_$tag $CMD > gpurun_out/ncu_full_b_$tag.log 2>&1
$CMD > gpurun_out/plain5_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv0|pre_|post_kernel|sppf|resmlp|mlp_gather|bbox_error|track_rows" -s 24 -c 8 -o gpurun_out/simt_$tag $CMD > gpurun_out/ncu_simt_$tag.log 2>&1
ls -la gpurun_out/ | grep $tag
du -sh gpurun_out
