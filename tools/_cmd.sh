export PYTHONPATH=$PWD
python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_u7.log 2>&1; head -2 gpurun_out/layers_u7.log
