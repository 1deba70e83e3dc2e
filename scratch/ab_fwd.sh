#!/bin/bash
# time the forward edge kernel inside the benchmark step for several library builds
cd /root/repo
for lib in "$@"; do
  L=/root/repo/graph-physics_b200/graphphysics_b200/lib/libgp_b200$lib.so
  GP_B200_LIB=$L timeout 300 python bench.py --no-secondary --no-partition --no-cpu-baseline --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels_ms_per_step']; print('$lib', round(d['ms_per_step'],3), {a: round(b/15*1000,1) for a,b in k.items()})"
done
