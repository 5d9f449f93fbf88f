set -x
python -m pytest tests/test_gpu_sim.py -x -q -m gpu 2>&1 | tail -15
python prof1.py cfg2 > gpurun_out/prof2.log 2>&1
