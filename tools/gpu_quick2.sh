#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7 | tee gpurun_out/kbench.log
for v in gpvecchia_b200/variants/lib_*.so; do GPV_LIB_PATH=$PWD/$v timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7; done
