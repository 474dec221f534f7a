# round-end measurement set (1 GPU): full GPU tests, bench line, per-kernel microbench, per-layer conv breakdown, ncu launch list
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_errors.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/t_all.log | head -20
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 400 python tools/profile_kernels.py > gpurun_out/microbench.jsonl 2> gpurun_out/mb.err; tail -2 gpurun_out/mb.err
timeout 200 python tools/eval_breakdown.py > gpurun_out/breakdown.md 2> gpurun_out/breakdown.err; head -3 gpurun_out/breakdown.md
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/eval_breakdown.py > /dev/null 2>&1; wc -l gpurun_out/launches_eval.csv
