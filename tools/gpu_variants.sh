#!/bin/bash
# A/B timing of the builds under gpvecchia_b200/variants (tools/build_variant.sh); KB_M selects m
mkdir -p gpurun_out
for v in gpvecchia_b200/variants/lib_*.so; do
  GPV_LIB_PATH=$PWD/$v timeout 300 python tools/kbench.py 1000000 ${KB_M:-30} 2 2>&1 | tail -6 | grep -v "nu0.5\|nu2.5"
done | tee gpurun_out/variants.log
