#!/bin/bash
# A/B timing of the builds under gpvecchia_b200/variants (tools/build_variant.sh) next to the default build
mkdir -p gpurun_out
( timeout 300 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6
  for v in gpvecchia_b200/variants/lib_*.so; do
    for cfg in ${KB_CFGS:-30_2}; do
      GPV_LIB_PATH=$PWD/$v KBENCH_CHECK=${KBENCH_CHECK:-1} timeout 300 python -u tools/kbench.py 1000000 ${cfg%_*} ${cfg#*_} 2>&1 | tail -6
    done
  done ) | tee gpurun_out/variants.log
