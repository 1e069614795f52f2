export PYTHONPATH=$PWD
bash tools/gpu_round.sh u14
WT_CONV_PDL=0 timeout 60 python tools/gpu_layer_times.py 64 640 2>&1 | head -1
timeout 60 python tools/gpu_layer_times.py 64 640 2>&1 | head -1
