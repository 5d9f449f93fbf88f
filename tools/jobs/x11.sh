rm -f gpurun_out/x11_time.log
for cfg in "1.25 0.62 20" "1.25 0.62 17" "1.6 0.62 17" "2.0 0.62 17" "2.5 0.62 17" "1.6 0.72 17" "3.0 0.72 17"; do
  set -- $cfg
  echo "== rho=$1 load=$2 unit_lg=$3" >> gpurun_out/x11_time.log
  XMAP_XSIM_RHO=$1 XMAP_XSIM_LOAD=$2 XMAP_XSIM_CTA_UNIT_LG=$3 timeout 300 python tools/xsim_time.py cfg2 2>&1 | grep -v "^lib" | tail -3 >> gpurun_out/x11_time.log
done
cat gpurun_out/x11_time.log
