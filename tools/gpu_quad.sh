#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7
  GPV_KERNEL_FAMILY=fold timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7
  timeout 300 python tools/kbench.py 1000000 31 2 2>&1 | tail -7
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/quad.log
