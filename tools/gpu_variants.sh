#!/bin/bash
# A/B timing of the builds under gpvecchia_b200/variants (tools/build_variant.sh); KB_CFG = "m d" list
mkdir -p gpurun_out
for v in gpvecchia_b200/variants/lib_*.so; do
  for cfg in ${KB_CFGS:-30_2}; do
    GPV_LIB_PATH=$PWD/$v KBENCH_CHECK=${KBENCH_CHECK:-1} timeout 300 python tools/kbench.py 1000000 ${cfg%_*} ${cfg#*_} 2>&1 | tail -6 | grep -v "nu0.5\|nu2.5"
  done
done | tee gpurun_out/variants.log
