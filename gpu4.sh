set -x
python prof1.py cfg2 > gpurun_out/prof1.log 2>&1
python bench.py --workload cfg2_small --steps 2 --warmup 1 --no-pipeline --no-cpu > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_small.csv python bench.py --workload cfg2_small --steps 2 --warmup 1 --no-pipeline --no-cpu > gpurun_out/ncu_small.log 2>&1
