#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -6
  timeout 300 python tools/kbench.py 1000000 20 2 2>&1 | tail -6
  timeout 300 python tools/kbench.py 1000000 25 2 2>&1 | tail -6
  timeout 300 python tools/kbench.py 1000000 40 3 2>&1 | tail -6
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/band.log
