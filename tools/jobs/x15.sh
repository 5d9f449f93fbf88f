# X-SIM: parity tests, then warp kernel register budgets / cta at cfg2
timeout 600 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x15_tests.log 2>&1; tail -6 gpurun_out/x15_tests.log
timeout 420 python tools/xsim_sweep.py cfg2 "warp 1 9 17" "warp 1 9 17 1.25 0.62 16" "warp 1 9 17 1.25 0.62 18" "warp 1 9 17 1.25 0.62 12" "cta 1 12 17" > gpurun_out/x15_sweep.log 2>&1; grep -v "^lib" gpurun_out/x15_sweep.log
