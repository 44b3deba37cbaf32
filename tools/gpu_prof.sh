#!/bin/bash
# profile pass: launch list + one full capture of the set kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:u_sets -s 2 -c 1 -f -o gpurun_out/prof_u_sets \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu2.log 2>&1
ls -la gpurun_out
