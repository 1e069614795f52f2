# chained-conv bring-up: chain selftests first; the full round only if they pass, else the unchained numbers
export PYTHONPATH=$PWD
tag=${1:-x}
WT_CASE_TIMEOUT=25 timeout 240 python tools/gpu_conv_selftest.py --chain-only > gpurun_out/chain_$tag.log 2>&1
cat gpurun_out/chain_$tag.log | cut -c1-300
if tail -1 gpurun_out/chain_$tag.log | grep -q "^7/7"; then
  bash tools/gpu_round.sh $tag; bash tools/_cmd.sh
else
  export WT_CHAIN=0
  timeout 60 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1; head -2 gpurun_out/layers_$tag.log
  timeout 150 python -m pytest tests -m gpu -q -x --timeout 60 2>&1 | tail -3
fi
