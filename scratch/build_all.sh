#!/bin/bash
# Build the product library and the phase-profiling variant of the pair kernel (tuning aid).
set -e
cd /root/repo/graph-physics_b200
bash build.sh 2>&1 | grep -v "deprecated\|177-D\|declared but never\|\^\|^$\|Remark" | tail -3
mkdir -p build_prof
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGP_FWD2_PROF -c csrc/edge_fwd2.cu -o build_prof/edge_fwd2.o 2>&1 | grep "error" || true
nvcc -shared -o graphphysics_b200/lib/libgp_b200_prof.so build_prof/edge_fwd2.o $(ls build/*.o | grep -v edge_fwd2) 2>&1 | grep -v deprecated || true
ls -la graphphysics_b200/lib/
