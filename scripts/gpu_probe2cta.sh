#!/bin/bash
timeout 60 ./scripts/build/tc_probe_2cta_pipe 2>&1 | tee gpurun_out/tc_probe_2cta_pipe.log; echo "rc=${PIPESTATUS[0]}"
