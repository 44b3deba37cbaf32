#!/bin/bash
# First GPU contact of the warp-specialised experiment (u_band_ws.cuh, GPV_KERNEL_FAMILY=ws): kernel time and
# oracle parity next to the default kernel.  Every run sits under its own `timeout`: the experiment hands slots
# between warps through spin-waits in shared memory and has only run on the host emulation so far.
mkdir -p gpurun_out
for fam in default ws; do
  if [ $fam = ws ]; then export GPV_KERNEL_FAMILY=ws; else unset GPV_KERNEL_FAMILY; fi
  timeout 90 python -u tools/kbench.py ${1:-1000000} 30 2 > gpurun_out/ws_kbench_$fam.log 2>&1
  echo "rc=$?" >> gpurun_out/ws_kbench_$fam.log
  tail -8 gpurun_out/ws_kbench_$fam.log
done
