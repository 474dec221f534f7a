mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_conv_tc" --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/profile_kernels.py --only conv3x3_c 2> gpurun_out/mb_c.err | cut -c1-120; tail -2 gpurun_out/mb_c.err
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
