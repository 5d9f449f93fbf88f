set -x
python bench.py --workload cfg2 --steps 1 --warmup 1 --no-pipeline --no-cpu > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"big_|sim_hash|sim_big" -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --workload cfg2 --steps 1 --warmup 1 --no-pipeline --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
