# record-list X-SIM kernel, second version (chunk pre-combine + O(g) reducer): parity tests, then a bounded sweep at cfg2
timeout 600 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x13_tests.log 2>&1; tail -6 gpurun_out/x13_tests.log
timeout 200 python tools/xsim_sweep.py cfg2_small "warp 1 9 17" "ll 1 12 17 1.25 0.62 16 11" "ll 1 11 17 1.25 0.62 8 10" > gpurun_out/x13_small.log 2>&1; grep -v "^lib" gpurun_out/x13_small.log
timeout 420 python tools/xsim_sweep.py cfg2 \
  "warp 1 9 17" \
  "ll 1 12 17 1.25 0.62 16 11" "ll 1 11 17 1.25 0.62 8 10" "ll 1 12 17 1.25 0.62 16 10" "ll 1 11 17 1.25 0.62 8 11" \
  "ll 1 12 17 1.6 0.72 16 11" "ll 1 13 17 1.25 0.62 16 9" "ll 0 12 17 1.25 0.62 16 11" "ll 1 12 18 1.25 0.62 16 11" "ll 1 12 17 1.25 0.62 8 11" \
  > gpurun_out/x13_sweep.log 2>&1; grep -v "^lib" gpurun_out/x13_sweep.log
