#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_ops.log
