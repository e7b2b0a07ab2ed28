#!/bin/bash
timeout 300 python scripts/e2e_probe.py 2>&1 | tee gpurun_out/e2e_probe.log
