set -x
python -m pytest tests -x -q -m gpu 2>&1 | grep -E "passed|failed" | tail -3
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-pipeline --no-cpu"
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sim_|big_|_kernel" -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sim_big_accum_kernel|big_eval_kernel|sim_warp_gmem_kernel|sim_warp_smem_kernel|big_select_kernel" -c 12 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_r1_full.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
