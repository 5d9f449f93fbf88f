export XMAP_XSIM_MODE=cta
timeout 200 python tools/prof_xsim.py cfg2_small > gpurun_out/x9_plain.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:xsim_cta -c 1 -o gpurun_out/x9_prof python tools/prof_xsim.py cfg2_small > gpurun_out/x9_ncu.log 2>&1
tail -2 gpurun_out/x9_plain.log; tail -2 gpurun_out/x9_ncu.log
