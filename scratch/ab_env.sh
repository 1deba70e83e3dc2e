#!/bin/bash
# scratch/ab_env.sh VAR v1 v2 ...: headline bench line under different values of an environment switch
cd /root/repo
V=$1; shift
for val in "$@"; do
  env $V=$val timeout 300 python bench.py --no-secondary --no-partition --no-cpu-baseline --steps 20 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels_ms_per_step']; print('$V=$val', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'], {a: round(b/15*1000,1) for a,b in k.items()})"
done
