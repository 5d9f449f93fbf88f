# hybrid X-SIM scheduling with the two kernels on two streams
timeout 600 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x19_tests.log 2>&1; tail -4 gpurun_out/x19_tests.log
for hp in 524288 2097152; do
  echo "== hot_paths=$hp"; XMAP_XSIM_HOT_PATHS=$hp timeout 300 python tools/xsim_share_time.py cfg2 8 2>&1 | grep -v "^lib" | tail -11
done > gpurun_out/x19_shares.log 2>&1; cat gpurun_out/x19_shares.log
