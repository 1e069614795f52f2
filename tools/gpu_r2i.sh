# Split-epilogue check: conv selftests (tuning library, both settings), same-GPU A/B, then the GPU tests on the product library.
export PYTHONPATH=$PWD
tag=$1
bash tools/gpu_ab2.sh $tag "WT_EPI_SPLIT=1" "WT_EPI_SPLIT=0" "WT_EPI_SPLIT=1" "WT_EPI_SPLIT=0" "WT_EPI_SPLIT=1"
unset WTRACKER_B200_LIB
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8 > gpurun_out/tests_$tag.log; tail -4 gpurun_out/tests_$tag.log
