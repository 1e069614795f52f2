# Round-2 profile pass (run under gpurun; ONE ncu-wrapped program invocation at a time, each after a plain run of the same
# command exited 0).  Step = 57 launches: pre_crop, conv0_tc, 52 x conv_tc / conv_halo, sppf_pool, post_top1, hot_tail.
# The bench steps in order: 3 warm-up, 2 (estimation repeat), 2 timed, 2 (result rows), then the e2e loops.
export PYTHONPATH=$PWD
tag=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --min-seconds 0"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__cycles_elapsed.avg
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 500 ncu --metrics $M --clock-control none -c 1300 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
timeout 500 ncu --set full --clock-control none -k regex:conv_ -s 260 -c 52 --csv --page raw --log-file gpurun_out/conv_step_$tag.csv $CMD > gpurun_out/ncu_step_$tag.log 2>&1
$CMD > gpurun_out/plain3_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 260 -c 9 -o gpurun_out/conv_a_$tag $CMD > gpurun_out/ncu_full_a_$tag.log 2>&1
$CMD > gpurun_out/plain4_$tag.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv0|pre_|post_|sppf|hot_tail" -s 25 -c 5 -o gpurun_out/simt_$tag $CMD > gpurun_out/ncu_simt_$tag.log 2>&1
ls -la gpurun_out/ | grep $tag
du -sh gpurun_out
