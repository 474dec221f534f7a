mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_conv_tc" --tb=line -p no:cacheprovider 2>&1 | tail -4
LFDM_CONV_PAIR=1 timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_conv_tc" --tb=line -p no:cacheprovider 2>&1 | tail -6
