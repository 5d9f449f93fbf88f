set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-pipeline --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -5 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; tail -5 gpurun_out/multi_check.log
