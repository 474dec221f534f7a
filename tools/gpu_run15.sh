mkdir -p gpurun_out
timeout 600 python bench.py --config natops128 --steps 3 --warmup 3 > gpurun_out/bench_natops128.json 2> gpurun_out/bench_natops128.err; cut -c1-300 gpurun_out/bench_natops128.json; tail -2 gpurun_out/bench_natops128.err
timeout 900 python bench.py --config mug256 --steps 3 --warmup 3 > gpurun_out/bench_mug256.json 2> gpurun_out/bench_mug256.err; cut -c1-300 gpurun_out/bench_mug256.json; tail -2 gpurun_out/bench_mug256.err
