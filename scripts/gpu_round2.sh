#!/bin/bash
# bench + ncu full captures of the two tcgen05 kernels at phi size (perf_probe = phi stack only)
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 12 -c 2 -f -o gpurun_out/prof_linear_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_lin.log 2>&1; echo "ncu-lin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_wgrad_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_wg.log 2>&1; echo "ncu-wg rc=$?"
tail -30 gpurun_out/ncu_wg.log
