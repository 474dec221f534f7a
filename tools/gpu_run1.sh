mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused" --tb=short -p no:cacheprovider -x > gpurun_out/t_fused.log 2>&1; tail -30 gpurun_out/t_fused.log
timeout 200 python tools/profile_kernels.py --only tblock > gpurun_out/mb_tblock.jsonl 2> gpurun_out/mb_tblock.err; cat gpurun_out/mb_tblock.jsonl; tail -5 gpurun_out/mb_tblock.err
