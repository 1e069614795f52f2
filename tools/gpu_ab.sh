# A/B of env knobs on the whole forward (tools/gpu_layer_times.py prints the whole-forward time first)
export PYTHONPATH=$PWD
for cfg in "WT_LANES=0" "WT_LANES=1" "WT_LANES=1 WT_CONV0_OCC=4" "$@"; do
  echo "== $cfg"; env $cfg timeout 60 python tools/gpu_layer_times.py 64 640 2>&1 | head -1
done
