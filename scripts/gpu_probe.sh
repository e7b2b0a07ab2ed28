#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python scripts/perf_probe.py 1024 128 8 2>&1 | tee gpurun_out/perf_probe.log
