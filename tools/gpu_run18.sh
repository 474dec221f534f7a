mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_conv3x3_c64 python tools/profile_kernels.py --only conv3x3_c64_32_nogn --iters 1 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_conv3x3_c512 python tools/profile_kernels.py --only conv3x3_c512_4 --iters 1 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
