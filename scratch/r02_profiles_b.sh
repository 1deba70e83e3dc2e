#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_bwd_kernel<.int.128, .int.1>" -s 20 -c 1 -f -o $O/r02_edge_bwd_B $B > $O/r02_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_bwd_kernel<.int.128, .int.2>" -s 20 -c 1 -f -o $O/r02_edge_bwd_A $B >> $O/r02_ncu_b.log 2>&1
tail -5 $O/r02_ncu_b.log
ls -la $O/r02_edge_bwd*
