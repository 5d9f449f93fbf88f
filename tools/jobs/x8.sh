timeout 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_recsim.py -q -x > gpurun_out/x8_tests.log 2>&1; tail -8 gpurun_out/x8_tests.log
timeout 600 python tools/xsim_modes.py cfg2 > gpurun_out/x8_modes.log 2>&1; tail -20 gpurun_out/x8_modes.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-pipeline --no-cpu > gpurun_out/x8_bench.json 2> gpurun_out/x8_bench.err; tail -2 gpurun_out/x8_bench.err; python -c "
import json; p=json.loads(open('gpurun_out/x8_bench.json').read().strip().splitlines()[-1]); print(p['ms_per_step'], p['roofline']['per_kernel_ms_per_step'], p['e2e']['ms_per_step'])"
