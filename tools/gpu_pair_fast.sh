#!/bin/bash
# One-call check of a variant build (tools/build_variant.sh <name> ...): kernel time + oracle parity for the
# five covariances at n = 1e6, m = 30, then the GPU test-suite against the same build.
v=${1:-fast}
mkdir -p gpurun_out
export GPV_LIB_PATH=$PWD/gpvecchia_b200/variants/lib_$v.so
timeout 60 python -u tools/kbench.py 1000000 30 2 > gpurun_out/${v}_kbench.log 2>&1
echo "kbench rc=$?" >> gpurun_out/${v}_kbench.log
timeout 100 python -u -m pytest tests -m gpu -x -v -p no:cacheprovider > gpurun_out/${v}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${v}_pytest.log
tail -7 gpurun_out/${v}_kbench.log; tail -3 gpurun_out/${v}_pytest.log
