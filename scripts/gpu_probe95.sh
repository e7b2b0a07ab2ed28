#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest.log
timeout 300 python scripts/perf_probe.py 256 95 4 2>&1 | tail -22 | tee gpurun_out/perf_probe95.log
timeout 300 python scripts/perf_probe.py 128 64 8 2>&1 | tail -22 | tee gpurun_out/perf_probe64.log
timeout 600 python scripts/perf_probe.py 1024 128 8 2>&1 | grep -E "phi fwd|wgrad\[N=128,K=128|linear_fwd\[K=128,N=128" | tee gpurun_out/perf_probe.log
