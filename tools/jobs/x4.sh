timeout 200 python tools/prof_xsim.py cfg2_small > gpurun_out/x4_plain.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:xsim_warp -c 1 -o gpurun_out/x4_prof python tools/prof_xsim.py cfg2_small > gpurun_out/x4_ncu.log 2>&1
tail -2 gpurun_out/x4_plain.log; tail -2 gpurun_out/x4_ncu.log
