#!/bin/bash
# GPU test-suite + kernel timings (all covariances, oracle parity sample) for the current build and any variants
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
for cfg in ${KB_CFGS:-30_2 20_2 40_3}; do
  timeout 300 python -u tools/kbench.py 1000000 ${cfg%_*} ${cfg#*_} 2>&1 | tail -6
done | tee gpurun_out/r2_kbench.log
for v in gpvecchia_b200/variants/lib_*.so; do
  [ -f $v ] && GPV_LIB_PATH=$PWD/$v timeout 300 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6
done | tee gpurun_out/r2_kbench_variants.log
