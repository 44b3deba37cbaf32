#!/bin/bash
mkdir -p gpurun_out
( cd tools/microbench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o i2f i2f.cu && ./i2f ) 2>&1 | tee gpurun_out/microbench_i2f.log
bash tools/gpu_variants.sh
