#!/bin/bash
# round-end evidence: launch list of the bench command + full captures of the closed-form and general-nu set kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:u_(sets|band)' -s 4 -c 1 -f -o gpurun_out/prof_closed \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ncu2.log 2>&1
KBENCH_CHECK=0 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:u_(sets|band)' -s 26 -c 1 -f -o gpurun_out/prof_general \
   python tools/kbench.py 1000000 30 2 > gpurun_out/bench_ncu3.log 2>&1
ls -la gpurun_out | grep -E "prof_|launches"
