#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): launch lists + one --set full capture per hot kernel.
# Outputs land in gpurun_out/ and are summarised into profiles/ by profiles/summarize.py.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_final_launches.csv $B > $O/r02_final_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/r02_tfm_launches.csv python scratch/tfm_step.py 2 > $O/r02_tfm_launches.log 2>&1
# steady-state launches of the three edge kernels inside a real training step (skip the first layers' launches)
ncu --set full --clock-control none --import-source on -k regex:edge_fwd2_kernel -s 20 -c 1 -f -o $O/r02_edge_fwd $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_bwd_kernel<.int.128, .int.1>" -s 20 -c 1 -f -o $O/r02_edge_bwd_B $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mlp_bwd_kernel<.int.128, .int.2>" -s 20 -c 1 -f -o $O/r02_edge_bwd_A $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 6 -c 1 -f -o $O/r02_attn_fwd python scratch/tfm_step.py 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_rows_kernel -s 3 -c 1 -f -o $O/r02_attn_bwd_rows python scratch/tfm_step.py 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_cols_kernel -s 3 -c 1 -f -o $O/r02_attn_bwd_cols python scratch/tfm_step.py 1 > /dev/null 2>&1
ls -la $O/r02_*
