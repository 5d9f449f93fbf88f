# warp-kernel sweep with fused lists + ncu full captures of warp and ll on cfg2_small
timeout 420 python tools/xsim_sweep.py cfg2 \
  "warp 1 9 17" "warp 1 10 17 1.25 0.62" "warp 1 9 16" "warp 1 9 18" "warp 1 9 17 1.6 0.72" "warp 1 9 17 1.25 0.75" "warp 1 8 17" "cta 1 12 17" "cta 1 13 17" \
  > gpurun_out/x14_sweep.log 2>&1; grep -v "^lib" gpurun_out/x14_sweep.log
XMAP_XSIM_WARPS=10 timeout 300 python tools/xsim_sweep.py cfg2 "warp 1 10 17" "warp 1 10 17 1.6 0.72" > gpurun_out/x14_sweep2.log 2>&1; grep -v "^lib\|^plan" gpurun_out/x14_sweep2.log
XMAP_XSIM_MODE=warp timeout 400 ncu --set full --clock-control none --import-source on -k regex:xsim_warp -c 1 -o gpurun_out/x14_warp python tools/prof_xsim.py cfg2_small > gpurun_out/x14_ncu_warp.log 2>&1; tail -2 gpurun_out/x14_ncu_warp.log
XMAP_XSIM_MODE=ll timeout 400 ncu --set full --clock-control none --import-source on -k regex:xsim_ll -c 1 -o gpurun_out/x14_ll python tools/prof_xsim.py cfg2_small > gpurun_out/x14_ncu_ll.log 2>&1; tail -2 gpurun_out/x14_ncu_ll.log
