export PYTHONPATH=$PWD
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r1c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv0|cls_logit|resmlp|post_kernel|pre_kernel|upsample|sppf' -s 24 -c 8 -o gpurun_out/misc_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc_r1.log 2>&1
tail -2 gpurun_out/ncu_misc_r1.log | cut -c1-300
