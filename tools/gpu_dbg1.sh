mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_r2.py "tests/test_gpu_kernels.py::test_conv_tc" -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -15
