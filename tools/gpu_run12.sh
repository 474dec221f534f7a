mkdir -p gpurun_out
timeout 200 python tools/profile_kernels.py --only lblock 2> gpurun_out/mb_l.err; tail -3 gpurun_out/mb_l.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lin_launches.csv python tools/profile_kernels.py --only lblock_fused --iters 1 > /dev/null 2>&1; grep -E "linattn" gpurun_out/lin_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
rm -f gpurun_out/r02_parity_errors.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/t_all.log | head -20
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
