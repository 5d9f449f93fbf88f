# round-2 evidence: GPU tests, 1-GPU bench line, launch list, full captures of the two dominant kernels
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e1_tests.log 2>&1; tail -6 gpurun_out/e1_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/e1_bench.json 2> gpurun_out/e1_bench.err; tail -3 gpurun_out/e1_bench.err; cut -c1-2500 gpurun_out/e1_bench.json
bash tools/jobs/ncu_final.sh
