timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_recsim.py tests/test_gpu_pipeline.py tests/test_gpu_sim.py -q > gpurun_out/x7_tests.log 2>&1; tail -30 gpurun_out/x7_tests.log
rm -f gpurun_out/x7_time.log
for cfg in "cta 12 20" "cta 11 20" "cta 12 19" "warp 9 17"; do
  set -- $cfg
  echo "== mode=$1 cells_lg=$2 unit_lg=$3" >> gpurun_out/x7_time.log
  XMAP_XSIM_MODE=$1 XMAP_XSIM_CTA_CELLS_LG=$2 XMAP_XSIM_CTA_UNIT_LG=$3 timeout 300 python tools/xsim_time.py cfg2 2>&1 | grep -v "^lib" | tail -3 >> gpurun_out/x7_time.log
done
cat gpurun_out/x7_time.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-pipeline --no-cpu > gpurun_out/x7_bench.json 2> gpurun_out/x7_bench.err; tail -2 gpurun_out/x7_bench.err; python -c "
import json; p=json.loads(open('gpurun_out/x7_bench.json').read().strip().splitlines()[-1]); print(p['ms_per_step'], p['roofline']['per_kernel_ms_per_step'], p['e2e']['ms_per_step'])"
