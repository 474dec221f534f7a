mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_errors.jsonl
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused" --tb=short -p no:cacheprovider -x > gpurun_out/t_fused.log 2>&1; tail -5 gpurun_out/t_fused.log
timeout 200 python tools/profile_kernels.py --only tblock_fused > gpurun_out/mb_tblock.jsonl 2> gpurun_out/mb_tblock.err; cat gpurun_out/mb_tblock.jsonl; tail -3 gpurun_out/mb_tblock.err
timeout 300 python tools/profile_kernels.py --only warp > gpurun_out/mb_warp.jsonl 2> gpurun_out/mb_warp.err; cat gpurun_out/mb_warp.jsonl; tail -3 gpurun_out/mb_warp.err
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all.log 2>&1; tail -25 gpurun_out/t_all.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_temporal_fused --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_attn_fused_v4 python tools/profile_kernels.py --only tblock_fused_32 --iters 1 > gpurun_out/ncu_fused.log 2>&1; tail -2 gpurun_out/ncu_fused.log
