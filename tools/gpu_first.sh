#!/bin/bash
# first GPU call: smoke, parity tests, short bench
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep "Model name"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
