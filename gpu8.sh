set -x
python -m pytest tests/test_gpu_extend.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -25
timeout 900 python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err
