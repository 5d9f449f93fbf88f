set -x
python -m pytest tests/test_gpu_extend.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | grep -vE "Warning|utcfrom|^\s*$" | tail -12
python prof5.py > gpurun_out/prof5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"xsim" -c 200 --csv --log-file gpurun_out/launches_xsim.csv python prof5.py > gpurun_out/ncu_xsim.log 2>&1
