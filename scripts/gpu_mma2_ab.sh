#!/bin/bash
# A/B of the match kernels on one box: one-CTA tensor-core kernel (UZ_MATCH_MMA=1) against the CTA-pair kernel (=2), both configurations
for rep in 1 2; do
for V in "1 0" "2 0" "7 0"; do
  set -- $V
  UZ_MATCH_MMA=$1 UZ_MMA2_CFG=$2 timeout 300 python bench.py --no-cpu-baseline --no-places --no-extras --steps 6 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('MATCH_MMA=$1 CFG=$2', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['pageable'], 'knn2', r['knn2_ms_per_launch'], 'solve', r['solve_ms_per_launch'])"
done
done
