#!/bin/bash
# solve time against resident solve CTAs per SM (UZ_SOLVE_SMEM_PAD pads the CTA's shared memory: 0 -> 5 CTAs, 12000 -> 4, 31000 -> 3, 68000 -> 2)
for P in 0 12000 31000 68000 0; do
  UZ_SOLVE_SMEM_PAD=$P timeout 300 python bench.py --no-cpu-baseline --no-places --no-extras --steps 6 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pad $P', 'ms', d['ms_per_step'], 'knn2', r['knn2_ms_per_launch'], 'solve', r['solve_ms_per_launch'])"
done
