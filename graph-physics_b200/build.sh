#!/bin/bash
# Build libgp_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
mkdir -p graphphysics_b200/lib build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${GP_EXTRA_FLAGS:-}"
pids=()
for f in csrc/*.cu; do
  o=build/$(basename "$f" .cu).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find csrc include ../include -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o" 2>/dev/null | head -1)" ]; then
    nvcc $FLAGS ${GP_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -o graphphysics_b200/lib/libgp_b200.so build/*.o
echo "built graphphysics_b200/lib/libgp_b200.so"
