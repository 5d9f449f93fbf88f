timeout 300 python tools/prof_xsim.py cfg2 > gpurun_out/x5_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:xsim_warp -c 1 -o gpurun_out/x5_prof python tools/prof_xsim.py cfg2 > gpurun_out/x5_ncu.log 2>&1
tail -2 gpurun_out/x5_plain.log; tail -2 gpurun_out/x5_ncu.log
