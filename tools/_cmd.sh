export PYTHONPATH=$PWD
python tools/gpu_detector_check.py 640 640 2>&1 | grep -E "^x|^m[0-9]|head|oracle" | head -40 > gpurun_out/det_exact.log
WT_SILU_TANH=1 python tools/gpu_detector_check.py 640 640 2>&1 | grep -E "^x|^m[0-9]|head|oracle" | head -40 > gpurun_out/det_tanh.log
python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_exact.log 2>&1; head -1 gpurun_out/layers_exact.log
WT_SILU_TANH=1 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_tanh.log 2>&1; head -1 gpurun_out/layers_tanh.log
python -m pytest tests -m gpu -q 2>&1 | tail -3
