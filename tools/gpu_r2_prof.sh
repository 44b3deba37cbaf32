#!/bin/bash
# ncu --set full captures of the shipped kernels (closed forms, general nu) at n = 1e6, m = 30
mkdir -p gpurun_out
prof() {  # name n m d skip
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_band -s $5 -c 1 -f -o gpurun_out/$1 \
     python tools/kbench.py $2 $3 $4 > gpurun_out/$1.log 2>&1
}
prof r02c_u_band_general_P31_D2_nu08 1000000 30 2 26
prof r02c_u_band_closed_P31_D2_nu15 1000000 30 2 10
ls -la gpurun_out | grep r02c
