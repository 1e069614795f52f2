# concat-chain bring-up: its selftests first; the full round only if they pass, else numbers without it
export PYTHONPATH=$PWD
tag=${1:-x}
WT_CASE_TIMEOUT=25 timeout 200 python tools/gpu_conv_selftest.py --cat-only > gpurun_out/cat_$tag.log 2>&1
cat gpurun_out/cat_$tag.log | cut -c1-300
if tail -1 gpurun_out/cat_$tag.log | grep -q "^4/4"; then
  bash tools/gpu_round.sh $tag; bash tools/_cmd.sh
else
  export WT_CHAIN_EXIT=0
  timeout 60 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1; head -2 gpurun_out/layers_$tag.log
fi
