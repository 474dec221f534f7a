mkdir -p gpurun_out
timeout 120 python tools/attn_trace.py --cold > gpurun_out/attn_trace_cold.txt 2>&1; sed -n 1,2p gpurun_out/attn_trace_cold.txt | cut -c1-360; sed -n 10,20p gpurun_out/attn_trace_cold.txt | cut -c1-360
timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace.txt 2>&1; sed -n 1p gpurun_out/attn_trace.txt
timeout 200 python tools/profile_kernels.py --only tblock_fused_32 --noflush 2>/dev/null
timeout 200 python tools/profile_kernels.py --only tblock_fused_32 2>/dev/null
