timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f2_tests.log 2>&1; tail -25 gpurun_out/f2_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 --workload cfg4_small --no-cpu > gpurun_out/f2_bench_cfg4s.json 2> gpurun_out/f2_bench_cfg4s.err; tail -3 gpurun_out/f2_bench_cfg4s.err; cut -c1-600 gpurun_out/f2_bench_cfg4s.json
