#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 12 -c 1 -f -o gpurun_out/prof_linear_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_lin.log 2>&1; echo "ncu-lin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 1 -f -o gpurun_out/prof_wgrad_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_wg.log 2>&1; echo "ncu-wg rc=$?"
