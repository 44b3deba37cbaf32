#!/bin/bash
# timing-only experiments (GPV_HACK_* builds give wrong values): where does the closed-form kernel's time go?
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
KBENCH_CHECK=0 timeout 120 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6
for v in gpvecchia_b200/variants/lib_*.so; do
  GPV_LIB_PATH=$PWD/$v KBENCH_CHECK=0 timeout 120 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6
done | tee gpurun_out/r2_hacks.log
