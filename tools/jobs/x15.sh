# pair descriptors + segmented run reduction: parity tests, then warp / cta at cfg2
timeout 600 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/x15_tests.log 2>&1; tail -6 gpurun_out/x15_tests.log
timeout 420 python tools/xsim_sweep.py cfg2 "warp 1 9 17" "warp 0 9 17" "cta 1 12 17" "warp 1 10 17 1.25 0.62 10" "warp 1 9 16" "warp 1 9 17 1.6 0.72" > gpurun_out/x15_sweep.log 2>&1; grep -v "^lib" gpurun_out/x15_sweep.log
XMAP_XSIM_MODE=warp timeout 400 ncu --set full --clock-control none --import-source on -k regex:xsim_warp -c 1 -o gpurun_out/x15_warp python tools/prof_xsim.py cfg2_small > gpurun_out/x15_ncu_warp.log 2>&1; tail -2 gpurun_out/x15_ncu_warp.log
