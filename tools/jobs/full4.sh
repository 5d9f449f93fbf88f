timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_recsim.py -q -s > gpurun_out/f4_tests.log 2>&1; tail -12 gpurun_out/f4_tests.log
