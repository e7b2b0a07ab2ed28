#!/bin/bash
# ncu --set full of the TMA-fed Linear kernel at the phi size (mode 3 on the prologue+statistics case, mode 4 on the plain case)
timeout 45 ncu --set full --import-source on --clock-control none -k regex:linear_tc_tma_kernel -s 1 -c 1 -f -o gpurun_out/prof_linear_tma3 ./scripts/build/pair_check 3 5 > gpurun_out/ncu_tma3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_tma3.log
timeout 45 ncu --set full --import-source on --clock-control none -k regex:linear_tc_tma_kernel -s 1 -c 1 -f -o gpurun_out/prof_linear_tma4 ./scripts/build/pair_check 4 6 > gpurun_out/ncu_tma4.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_tma4.log
