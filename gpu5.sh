set -x
python -m pytest tests/test_gpu_sim.py -x -q -m gpu 2>&1 | tail -15
python prof3.py > gpurun_out/prof3.log 2>&1
python prof1.py cfg2 > gpurun_out/prof2.log 2>&1
