#!/bin/bash
# Build the product library and the phase-profiling variant of the pair kernel (tuning aid).
set -e
cd /root/repo/graph-physics_b200
bash build.sh 2>&1 | grep -v "deprecated\|177-D\|declared but never\|\^\|^$\|Remark" | tail -3
mkdir -p build_prof
for f in edge_fwd2 mlp_fwd mlp_bwd; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGP_FWD2_PROF -DGP_MLP_PROF -c csrc/$f.cu -o build_prof/$f.o 2>&1 | grep "error" || true
done
nvcc -shared -o graphphysics_b200/lib/libgp_b200_prof.so build_prof/*.o $(ls build/*.o | grep -v "edge_fwd2\|mlp_fwd\|mlp_bwd") 2>&1 | grep -v deprecated || true
ls -la graphphysics_b200/lib/
