set -x
python -m pytest tests/test_gpu_sim.py tests/test_gpu_extend.py -x -q -m gpu 2>&1 | tail -60
