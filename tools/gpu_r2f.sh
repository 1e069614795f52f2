# Round-2 late check of the SIMT kernels: GPU tests, SIMT rooflines, sustained forward, layer times.  Usage: bash tools/gpu_r2f.sh <tag>
export PYTHONPATH=$PWD
tag=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/tests_$tag.log; tail -6 gpurun_out/tests_$tag.log
cat gpurun_out/closed_loop.json | head -1
timeout 300 python tools/gpu_simt_roofline.py > gpurun_out/simt_roofline_$tag.txt 2>gpurun_out/simt_roofline_$tag.err; cat gpurun_out/simt_roofline_$tag.txt; tail -3 gpurun_out/simt_roofline_$tag.err
timeout 120 python tools/gpu_sustained.py 64 640 3 2>&1 | tail -1
timeout 90 python tools/gpu_layer_times.py 64 640 > gpurun_out/layers_$tag.log 2>&1; head -1 gpurun_out/layers_$tag.log; grep -i "pool\|model.9" gpurun_out/layers_$tag.log | head
