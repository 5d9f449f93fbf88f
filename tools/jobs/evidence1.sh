# round-2 evidence: GPU tests, 1-GPU bench line, the other configs' legs, launch list, full captures of the dominant kernels
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e1_tests.log 2>&1; tail -6 gpurun_out/e1_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/e1_bench.json 2> gpurun_out/e1_bench.err; tail -3 gpurun_out/e1_bench.err; cut -c1-600 gpurun_out/e1_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e1_smoke.log 2>&1; tail -4 gpurun_out/e1_smoke.log
timeout 400 python tools/dropin_cfg1.py > gpurun_out/e1_dropin.json 2> gpurun_out/e1_dropin.err; tail -2 gpurun_out/e1_dropin.err; cut -c1-400 gpurun_out/e1_dropin.json
timeout 400 python tools/cfg3_leg.py cfg2_small > gpurun_out/e1_cfg3.json 2> gpurun_out/e1_cfg3.err; tail -2 gpurun_out/e1_cfg3.err; cut -c1-600 gpurun_out/e1_cfg3.json
bash tools/jobs/ncu_final.sh
