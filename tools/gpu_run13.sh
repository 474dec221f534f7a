mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "linear_attention_block_fused" --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 200 python tools/profile_kernels.py --only lblock_fused 2> gpurun_out/mb_l.err; tail -3 gpurun_out/mb_l.err
