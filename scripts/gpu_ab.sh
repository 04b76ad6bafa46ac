#!/bin/bash
# A/B of two builds of the library on one box: scripts/gpu_ab.sh <lib A> <lib B> [bench args]
A=$1; B=$2; shift 2
for rep in 1 2; do
for L in $A $B; do
  UZ_LIB_PATH=$L timeout 300 python bench.py --no-cpu-baseline --no-places --no-extras "$@" 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$L', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'knn2', r['knn2_ms_per_launch'], 'solve', r['solve_ms_per_launch'])"
done
done
