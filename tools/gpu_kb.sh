#!/bin/bash
# usage: gpu_kb.sh [m] -- kernel micro-benchmark of the default build (quad where it applies) with parity check
mkdir -p gpurun_out
timeout 300 python tools/kbench.py 1000000 ${1:-30} 2 2>&1 | tail -7 | tee gpurun_out/kb.log
