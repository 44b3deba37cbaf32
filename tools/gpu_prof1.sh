#!/bin/bash
# usage: gpu_prof1.sh <out-name> [m] [kernel-regex] [skip]
mkdir -p gpurun_out
KBENCH_CHECK=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:${3:-u_band} -s ${4:-10} -c 1 -f -o gpurun_out/$1 \
   python tools/kbench.py 1000000 ${2:-30} 2 > gpurun_out/$1.log 2>&1
ls -la gpurun_out | grep $1
