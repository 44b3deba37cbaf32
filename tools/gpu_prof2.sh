#!/bin/bash
mkdir -p gpurun_out
V=${1:-base}
GPV_LIB_PATH=$PWD/gpvecchia_b200/variants/lib_$V.so KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_sets -s 10 -c 1 -f -o gpurun_out/prof_$V python tools/kbench.py 1000000 30 2 > gpurun_out/prof_$V.log 2>&1
tail -3 gpurun_out/prof_$V.log
