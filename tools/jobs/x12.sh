# new record-list X-SIM kernel + fused bridge lists: parity tests first, then the sweep at cfg2
timeout 900 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x12_tests.log 2>&1; tail -12 gpurun_out/x12_tests.log
timeout 900 python tools/xsim_sweep.py cfg2 \
  "cta 0 12 17" "cta 1 12 17" "warp 1 9 17" \
  "ll 1 12 17 1.25 0.62 16 11" "ll 0 12 17 1.25 0.62 16 11" "ll 1 12 17 1.25 0.62 16 12" "ll 1 12 17 1.25 0.62 16 10" \
  "ll 1 11 17 1.25 0.62 8 10" "ll 1 11 17 1.25 0.62 8 11" "ll 1 12 17 1.25 0.62 8 11" "ll 1 13 17 1.25 0.62 16 9" \
  "ll 1 12 17 1.6 0.72 16 11" "ll 1 12 18 1.25 0.62 16 11" "ll 1 12 16 1.25 0.62 16 11" \
  > gpurun_out/x12_sweep.log 2>&1; cat gpurun_out/x12_sweep.log | grep -v "^lib"
