# cfg2 only at N GPUs
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/mg${N}_bench_cfg2.json 2> gpurun_out/mg${N}_bench_cfg2.err; tail -3 gpurun_out/mg${N}_bench_cfg2.err; python -c "
import json; p=json.loads(open('gpurun_out/mg${N}_bench_cfg2.json').read().strip().splitlines()[-1]); print(p['value'], p['ms_per_step'], p['e2e']['ms_per_step'], p.get('multi_parity'), p['pipeline'])"
