#!/bin/bash
# Final round-2 evidence after the two-threads-per-row forward kernel: launch list, full capture of the forward edge kernel,
# default bench line (run under gpurun, one GPU).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_final_launches.csv $B > $O/r02_final_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edge_fwd2_kernel -s 20 -c 1 -f -o $O/r02_edge_fwd $B > /dev/null 2>&1
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
tail -c 600 $O/r02_bench_1gpu.json
ls -la $O/r02_edge_fwd.ncu-rep $O/r02_final_launches.csv
