timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f1_tests.log 2>&1; tail -15 gpurun_out/f1_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; tail -3 gpurun_out/f1_bench.err; cat gpurun_out/f1_bench.json | cut -c1-1500
