set -x
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-pipeline --no-cpu"
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"xmap" -c 1500 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sim_big_accum_kernel|big_eval_kernel|sim_warp_gmem_kernel|sim_warp_smem_kernel|big_select_kernel" -c 12 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_r1_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
