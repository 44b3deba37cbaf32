#!/bin/bash
mkdir -p gpurun_out
for v in gpvecchia_b200/variants/lib_*.so; do
  GPV_LIB_PATH=$PWD/$v timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | tail -7
done | tee gpurun_out/variants.log
