#!/bin/bash
timeout 600 python scripts/perf_probe.py 1024 128 8 2>&1 | grep -E "phi fwd|wgrad\[N=128,K=128|linear_fwd\[K=128,N=128|affine|bn_bwd" | tee gpurun_out/perf_probe.log
