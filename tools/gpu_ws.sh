#!/bin/bash
# First GPU contact of the warp-specialised experiment (u_band_ws.cuh, GPV_KERNEL_FAMILY=ws): kernel time and
# oracle parity next to the default kernel.  Every run sits under its own `timeout`: the experiment hands slots
# between warps through spin-waits in shared memory and has only run on the host emulation so far.
# Build the second variant first (here, on the CPU):  tools/build_variant.sh wsfin -DGPV_WS_FINISH_IN_PRODUCERS=1
mkdir -p gpurun_out
run() {   # tag, library ("" = default build), family ("" = default)
  if [ -n "$3" ]; then export GPV_KERNEL_FAMILY=$3; else unset GPV_KERNEL_FAMILY; fi
  if [ -n "$2" ]; then export GPV_LIB_PATH=$2; else unset GPV_LIB_PATH; fi
  timeout 90 python -u tools/kbench.py ${N:-1000000} 30 2 > gpurun_out/ws_kbench_$1.log 2>&1
  echo "rc=$?" >> gpurun_out/ws_kbench_$1.log
  tail -8 gpurun_out/ws_kbench_$1.log
}
run default "" ""
run ws "" ws
[ -f gpvecchia_b200/variants/lib_wsfin.so ] && run wsfin $PWD/gpvecchia_b200/variants/lib_wsfin.so ws
# parity on the GPU: bit-identical to the default kernel (opt-in test file)
unset GPV_LIB_PATH GPV_KERNEL_FAMILY
GPV_TEST_EXPERIMENTS=1 timeout 300 python -u -m pytest tests/test_experiments_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/ws_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/ws_pytest.log; tail -5 gpurun_out/ws_pytest.log
