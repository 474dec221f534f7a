# smoke() of the driver entry point + compute-sanitizer memcheck of the fused kernels' small test cases
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "linear_attention_block_fused and (3-128 or 5-256) or temporal_attention_block_fused and 2-10" --tb=line > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -10
