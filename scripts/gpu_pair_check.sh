#!/bin/bash
# Experimental Linear kernels vs the default tcgen05 kernel through the C ABI (no torch):  bash scripts/gpu_pair_check.sh [modes]
# modes: 2 = CTA pair, 3 = TMA-fed, 4 = TMA-fed with raw heads (sb_set_tensor_cores values); tc_probe checks the raw-head premise
timeout 30 ./scripts/build/tc_probe > gpurun_out/tc_probe.log 2>&1; echo "tc_probe rc=$?"; cat gpurun_out/tc_probe.log
for m in ${@:-3 4}; do
  timeout 60 ./scripts/build/pair_check $m > gpurun_out/pair_check_mode$m.log 2>&1; echo "pair_check $m rc=$?"
  tail -14 gpurun_out/pair_check_mode$m.log
done
