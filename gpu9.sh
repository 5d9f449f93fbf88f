set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
