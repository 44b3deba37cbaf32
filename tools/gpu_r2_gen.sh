#!/bin/bash
# tests + kbench + one ncu --set full capture of the general-nu band kernel
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -u tools/kbench.py 1000000 30 2 2>&1 | tail -6 | tee gpurun_out/r2_kbench.log
KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_band -s 26 -c 1 -f -o gpurun_out/r02b_u_band_general_P31_D2_nu08 \
     python tools/kbench.py 1000000 30 2 > gpurun_out/r02b_gen.log 2>&1
ls -la gpurun_out | grep r02b
