#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launch rc=$?"
