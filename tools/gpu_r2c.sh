# Same-GPU A/B of conv variants at steady state (power-capped regime) + per-layer burst times.
# Usage: bash tools/gpu_r2c.sh <tag> "<env assignments variant 1>" "<variant 2>" ...
# Needs the tuning build (python -m wtracker_b200.build --tuning, done HERE before gpurun: it travels with the snapshot).
export PYTHONPATH=$PWD
export WTRACKER_B200_LIB=$PWD/wtracker_b200/_native/tuning/libwtracker_b200.so
tag=$1; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 120 python tools/gpu_sustained.py 64 640 3 > gpurun_out/sust_${tag}_$i.log 2>&1 || echo "sustained $v FAILED"
  echo "[$v] $(tail -1 gpurun_out/sust_${tag}_$i.log)"
  env $v timeout 90 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_${tag}_$i.log 2>&1 || echo "layers $v FAILED"
  echo "[$v] $(head -1 gpurun_out/layers_${tag}_$i.log)"
done
