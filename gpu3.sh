set -x
timeout 300 python bench.py --workload tiny --steps 2 --warmup 1 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; tail -3 gpurun_out/bench_tiny.err
timeout 600 python bench.py --workload cfg2_small --steps 3 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -3 gpurun_out/bench_small.err
timeout 1500 python bench.py --workload cfg2 --steps 3 --warmup 1 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err
