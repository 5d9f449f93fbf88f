# launch list of one bench run + full captures of the dominant kernels (similarity: tri_cta; pipeline: xsim_warp + xsim_cta)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/n_plain.json 2> gpurun_out/n_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/n_ncu1.log 2>&1
tail -2 gpurun_out/n_plain.err; tail -2 gpurun_out/n_ncu1.log
timeout 300 python tools/prof_sim.py > gpurun_out/n_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tri_cta -c 14 -o gpurun_out/r2_sim_full python tools/prof_sim.py > gpurun_out/n_ncu2.log 2>&1
tail -2 gpurun_out/n_plain2.log; tail -2 gpurun_out/n_ncu2.log
timeout 300 python tools/prof_xsim.py cfg2 > gpurun_out/n_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xsim_ -c 3 -o gpurun_out/r2_xsim_full python tools/prof_xsim.py cfg2 > gpurun_out/n_ncu3.log 2>&1
tail -2 gpurun_out/n_plain3.log; tail -2 gpurun_out/n_ncu3.log
