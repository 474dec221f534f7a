mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_temporal_fused --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_attn_fused_v4 python tools/profile_kernels.py --only tblock_fused_32 --iters 1 > gpurun_out/ncu_fused.log 2>&1; tail -2 gpurun_out/ncu_fused.log
ls -la gpurun_out/*.ncu-rep
