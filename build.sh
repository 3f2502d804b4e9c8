#!/bin/bash
# Build libpna_sm100.so in-tree (sm_100a only).  __graft_entry__.build() runs the same command.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o pna_b200/libpna_sm100.so pna_b200/csrc/pna_aggregate.cu pna_b200/csrc/pna_csr.cu pna_b200/csrc/pna_misc.cu
