mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gn or norm" --tb=short -p no:cacheprovider 2>&1 | tail -4
timeout 120 python tools/profile_kernels.py --only gn_apply 2>/dev/null | cut -c1-160
LFDM_GN_GENERIC=1 timeout 120 python tools/profile_kernels.py --only gn_apply 2>/dev/null | cut -c1-160
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | cut -c1-200; tail -2 gpurun_out/bench.err
