#!/bin/bash
# A/B builds while tuning: scratch/build_variant.sh NAME FILE "FLAGS"  ->  lib/libgp_b200_NAME.so with csrc/FILE.cu rebuilt under FLAGS
# (run with GP_B200_LIB=.../libgp_b200_NAME.so)
set -e
cd /root/repo/graph-physics_b200
mkdir -p build_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $3 -c csrc/$2.cu -o build_var/$2_$1.o
nvcc -shared -o graphphysics_b200/lib/libgp_b200_$1.so build_var/$2_$1.o $(ls build/*.o | grep -v "/$2.o")
echo built graphphysics_b200/lib/libgp_b200_$1.so
