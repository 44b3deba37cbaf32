#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG=..]...   -> gpvecchia_b200/variants/lib_<name>.so
# Development helper: builds the C-ABI library with extra compile-time options for A/B timing.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/gpvecchia_b200/csrc
out=$root/gpvecchia_b200/variants
bd=/tmp/gpv_variant_$name
mkdir -p $out $bd
objs=""
for f in gpv_capi gpv_multi nn_search host_specify u_inst_B8_21 u_inst_B8_26 u_inst_B8_31 u_inst_B8_32 u_inst_B16_41 u_inst_P31 u_inst_P32 u_inst_P21 u_inst_P41 u_inst_P4 u_inst_P8 u_inst_P11 u_inst_P16 u_inst_P26 u_inst_P51 u_inst_P64; do
  ( /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
      -Xcompiler -fPIC,-O2 -I$root/include "$@" -c $src/$f.cu -o $bd/$f.o 2> $bd/$f.log || (cat $bd/$f.log; exit 1) ) &
  objs="$objs $bd/$f.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $out/lib_$name.so $objs -lcudart_static -lrt -lpthread -ldl
echo built $out/lib_$name.so
