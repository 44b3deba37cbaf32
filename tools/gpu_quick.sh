#!/bin/bash
# quick loop: GPU parity tests + kernel micro-benchmark of the default build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7 | tee gpurun_out/kbench.log
