mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_conv_tc" --tb=line -p no:cacheprovider 2>&1 | tail -3
timeout 120 python tools/profile_kernels.py --only conv3x3_c64 2> gpurun_out/mb_c.err | cut -c1-120; tail -2 gpurun_out/mb_c.err
LFDM_CONV_NA2=1 timeout 120 python tools/profile_kernels.py --only conv3x3_c64 2> gpurun_out/mb_c.err | cut -c1-120
