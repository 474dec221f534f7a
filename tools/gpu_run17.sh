mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_conv_tc" --tb=short -p no:cacheprovider -x 2>&1 | tail -12
timeout 120 python tools/profile_kernels.py --only conv3x3_c 2> gpurun_out/mb_c.err | cut -c1-120; tail -2 gpurun_out/mb_c.err
