#!/bin/bash
# N-GPU bench launched exactly as the driver does
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; tail -1 gpurun_out/bench_n$N.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
