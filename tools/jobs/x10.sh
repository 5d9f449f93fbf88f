timeout 600 python -m pytest tests/test_gpu_sim.py -q -x -k mid_size > gpurun_out/x10_tests.log 2>&1; grep -n "AssertionError" gpurun_out/x10_tests.log | head -3; tail -3 gpurun_out/x10_tests.log
timeout 600 python tools/xsim_modes.py cfg2 > gpurun_out/x10_modes.log 2>&1; tail -12 gpurun_out/x10_modes.log
bash tools/jobs/x9.sh
