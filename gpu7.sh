set -x
CMD="python bench.py --workload cfg2 --steps 1 --warmup 1 --no-pipeline --no-cpu"
$CMD > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"big_eval_kernel" -s 0 -c 6 -o gpurun_out/prof_eval $CMD > gpurun_out/ncu_eval.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sim_hash_kernel|sim_big_accum" -s 0 -c 3 -o gpurun_out/prof_hash $CMD > gpurun_out/ncu_hash.log 2>&1
ls -la gpurun_out/*.ncu-rep
