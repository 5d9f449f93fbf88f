timeout 600 python -m pytest tests/test_gpu_extend.py tests/test_gpu_sim.py -x -q > gpurun_out/x6_tests.log 2>&1; tail -15 gpurun_out/x6_tests.log
rm -f gpurun_out/x6_time.log
for cfg in "9 20 32 19" "9 20 16 19" "9 20 64 19" "9 16 32 19" "10 10 32 19" "9 20 32 18"; do
  set -- $cfg
  echo "== cells_lg=$1 warps=$2 max_passes=$3 unit_lg=$4" >> gpurun_out/x6_time.log
  XMAP_XSIM_CELLS_LG=$1 XMAP_XSIM_WARPS=$2 XMAP_XSIM_MAX_PASSES=$3 XMAP_XSIM_UNIT_LG=$4 timeout 300 python tools/xsim_time.py cfg2 2>&1 | grep -v "^lib" | tail -3 >> gpurun_out/x6_time.log
done
cat gpurun_out/x6_time.log
