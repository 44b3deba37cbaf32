#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6 | tee gpurun_out/r2_kbench.log
for loc in 0 1; do
  echo "GPV_LOCALITY=$loc n=8e6"
  GPV_LOCALITY=$loc KBENCH_CHECK=1 timeout 600 python -u tools/kbench.py 8000000 30 2 2>&1 | tail -6
done | tee gpurun_out/r2_kbench_8e6.log
GPV_LOCALITY=1 timeout 300 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6 | tee gpurun_out/r2_kbench_loc1e6.log
