mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/t.log | head
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
timeout 200 python tools/profile_kernels.py > gpurun_out/microbench.jsonl 2>>gpurun_out/mb.err
timeout 200 python tools/eval_breakdown.py > gpurun_out/breakdown.md 2>/dev/null; head -3 gpurun_out/breakdown.md
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/eval_breakdown.py > /dev/null 2>&1; wc -l gpurun_out/launches_eval.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 188 --launch-count 10 -f -o gpurun_out/r01_conv_eval python tools/eval_breakdown.py > /dev/null 2>&1; ls -la gpurun_out/*.ncu-rep | tail -2
