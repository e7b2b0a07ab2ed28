#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fast2_bwd -s 3 -c 1 -f -o gpurun_out/prof_att_bwd \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_att.log 2>&1; echo "ncu-att rc=$?"
