mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linattn_apply --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_linattn_apply python tools/profile_kernels.py --only lblock_fused_32 --iters 1 > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linattn_ctx --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_linattn_ctx python tools/profile_kernels.py --only lblock_fused_32 --iters 1 > gpurun_out/ncu_l2.log 2>&1; tail -1 gpurun_out/ncu_l2.log
