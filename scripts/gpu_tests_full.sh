#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_full.log
