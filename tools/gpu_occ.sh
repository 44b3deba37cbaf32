#!/bin/bash
for b in 1 2 3 4; do echo "blocks/SM=$b"; GPV_BLOCKS_PER_SM=$b KBENCH_CHECK=0 timeout 300 python tools/kbench.py 1000000 30 2 2>&1 | grep -E "nu1.5|gen0.8"; done | tee gpurun_out/occ.log
