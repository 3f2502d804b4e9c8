#!/bin/bash
# build_variant.sh NAME "<extra nvcc flags>": libpna variant with the f32 aggregation TUs recompiled, other objects reused
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
d=variants/$name; mkdir -p $d
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC"
for tu in pna_aggregate_f32_vec pna_aggregate_f32_fsplit; do
  nvcc $FL "$@" -c pna_b200/csrc/$tu.cu -o $d/$tu.o &
done
wait
objs=""
for o in pna_b200/csrc/build/*.o; do b=$(basename $o); if [ -f $d/$b ]; then objs="$objs $d/$b"; else objs="$objs $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libpna_sm100.so $objs
echo built $d/libpna_sm100.so
