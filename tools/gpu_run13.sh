mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "linear_attention_block_fused" --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/profile_kernels.py --only lblock_fused 2> gpurun_out/mb_l.err; tail -3 gpurun_out/mb_l.err
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/eval_breakdown.py > /dev/null 2>&1; wc -l gpurun_out/launches_eval.csv
