# A/B of env knobs on the whole forward (tools/gpu_layer_times.py prints the whole-forward time first, then per-op)
# usage: bash tools/gpu_ab.sh "ENV1=a ENV2=b" "ENV3=c" ...   ("-" = defaults)
export PYTHONPATH=$PWD
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg="WT_NOOP=1"
  echo "== $cfg"; env $cfg timeout 60 python tools/gpu_layer_times.py 64 640 2>&1 | sed -n '1p;3p'
done
