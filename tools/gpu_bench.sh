#!/bin/bash
# tests + both bench arms at N=1 (what the driver runs at round end)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
