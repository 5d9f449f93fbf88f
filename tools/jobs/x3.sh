set -x
timeout 300 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x3_tests.log 2>&1; tail -15 gpurun_out/x3_tests.log
for cfg in "9 16" "9 20" "10 8" "10 10" "11 5"; do
  set -- $cfg
  echo "== cells_lg=$1 warps=$2" >> gpurun_out/x3_time.log
  XMAP_XSIM_CELLS_LG=$1 XMAP_XSIM_WARPS=$2 timeout 300 python tools/xsim_time.py cfg2 >> gpurun_out/x3_time.log 2>&1
done
cat gpurun_out/x3_time.log | grep -v "^lib"
