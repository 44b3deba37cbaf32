#!/bin/bash
mkdir -p gpurun_out
# kbench order: nu0.5 x8, nu1.5 x8 ... -> skip 10 launches of the set kernel = third nu=1.5 launch
KBENCH_CHECK=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:u_quad -s 10 -c 1 -f -o gpurun_out/prof_quad31 \
   python tools/kbench.py 1000000 30 2 > gpurun_out/prof_quad31.log 2>&1
KBENCH_CHECK=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:u_quad -s 10 -c 1 -f -o gpurun_out/prof_quad32 \
   python tools/kbench.py 1000000 31 2 > gpurun_out/prof_quad32.log 2>&1
ls -la gpurun_out | grep prof_quad
