#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 300 python scripts/agg_bench.py 1024 128 20 2>&1 | tee gpurun_out/agg_bench.log
timeout 300 python scripts/agg_bench.py 1024 64 20 2>&1 | tee -a gpurun_out/agg_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_agg_tma -s 6 -c 2 -f -o gpurun_out/prof_gin_agg_fwd \
    python scripts/agg_bench.py 1024 128 4 > gpurun_out/ncu_full_fwd.log 2>&1; echo "ncu-fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_agg_tma -s 12 -c 2 -f -o gpurun_out/prof_gin_agg_bwd \
    python scripts/agg_bench.py 1024 128 4 > gpurun_out/ncu_full_bwd.log 2>&1; echo "ncu-bwd rc=$?"
