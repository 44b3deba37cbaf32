#!/bin/bash
# round-end evidence: launch list of the bench command + full captures of the set kernels that ship
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-north-star > gpurun_out/bench_ncu1.log 2>&1
prof() {  # name n m d skip
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_band -s $5 -c 1 -f -o gpurun_out/$1 \
     python tools/kbench.py $2 $3 $4 > gpurun_out/$1.log 2>&1
}
prof r02f_u_band_closed_P31_D2_nu15 1000000 30 2 10
prof r02f_u_band_general_P31_D2_nu08 1000000 30 2 26
prof r02f_u_band_closed_P41_D3_nu15 1000000 40 3 10
prof r02f_u_band_closed_P31_D2_nu15_n8e6 8000000 30 2 10
ls -la gpurun_out | grep -E "r02f_|r02_launches"
