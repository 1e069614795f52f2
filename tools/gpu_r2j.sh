# Split-epilogue rule check: selftests, A/B of the three settings, detector / fusion / layer tests on the product library.
export PYTHONPATH=$PWD
tag=$1
bash tools/gpu_ab2.sh $tag "WT_EPI_SPLIT=1" "WT_EPI_SPLIT=0" "WT_EPI_SPLIT=1" "WT_EPI_SPLIT=2" "WT_EPI_SPLIT=1"
unset WTRACKER_B200_LIB
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_fusion.py tests/test_gpu_layers.py tests/test_gpu_parity64.py tests/test_gpu_conv.py -m gpu -q --timeout 600 -x 2>&1 | tail -4
