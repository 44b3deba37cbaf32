#!/bin/bash
# Round 2, first GPU call: (1) first hardware run of the warp-specialised experiment next to the default kernel,
# (2) the EARLY_RCP variant, (3) fresh ncu --set full captures of the kernels that ship (closed, general, P=41),
# (4) the n = 8e6 capture whose DRAM traffic DESIGN.md 5 quotes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_gpu.txt
bash tools/gpu_ws.sh
GPV_LIB_PATH=$PWD/gpvecchia_b200/variants/lib_earlyrcp.so timeout 120 python -u tools/kbench.py 1000000 30 2 > gpurun_out/earlyrcp_kbench.log 2>&1
tail -7 gpurun_out/earlyrcp_kbench.log
prof() {  # name n m d skip
  KBENCH_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:u_band -s $5 -c 1 -f -o gpurun_out/$1 \
     python tools/kbench.py $2 $3 $4 > gpurun_out/$1.log 2>&1
}
prof r02_u_band_closed_P31_D2_nu15 1000000 30 2 10
prof r02_u_band_general_P31_D2_nu08 1000000 30 2 26
prof r02_u_band_closed_P41_D3_nu15 1000000 40 3 10
prof r02_u_band_closed_P31_D2_nu15_n8e6 8000000 30 2 10
ls -la gpurun_out | grep r02_
