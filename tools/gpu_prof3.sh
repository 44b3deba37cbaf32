#!/bin/bash
mkdir -p gpurun_out
KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_sets -s ${1:-26} -c 1 -f -o gpurun_out/prof_${2:-gen} python tools/kbench.py 1000000 30 2 > gpurun_out/prof_${2:-gen}.log 2>&1
tail -3 gpurun_out/prof_${2:-gen}.log
