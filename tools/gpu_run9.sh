mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused" --tb=short -p no:cacheprovider -x > gpurun_out/t_fused.log 2>&1; tail -8 gpurun_out/t_fused.log
timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; sed -n 1p gpurun_out/attn_trace.txt | cut -c1-360; sed -n 9,22p gpurun_out/attn_trace.txt | cut -c1-360
timeout 200 python tools/profile_kernels.py --only tblock_fused > gpurun_out/mb_tblock.jsonl 2> gpurun_out/mb_tblock.err; cat gpurun_out/mb_tblock.jsonl; tail -3 gpurun_out/mb_tblock.err
