timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -s > gpurun_out/f3_tests.log 2>&1; tail -12 gpurun_out/f3_tests.log
timeout 400 python tools/dropin_cfg1.py > gpurun_out/f3_dropin.json 2> gpurun_out/f3_dropin.err; tail -2 gpurun_out/f3_dropin.err; cut -c1-1200 gpurun_out/f3_dropin.json
timeout 400 python tools/cfg3_leg.py cfg2_small > gpurun_out/f3_cfg3.json 2> gpurun_out/f3_cfg3.err; tail -2 gpurun_out/f3_cfg3.err; cut -c1-1200 gpurun_out/f3_cfg3.json
timeout 400 python tools/dense_ab.py cfg2 1024 > gpurun_out/f3_dense.json 2> gpurun_out/f3_dense.err; tail -2 gpurun_out/f3_dense.err; cut -c1-1500 gpurun_out/f3_dense.json
