#!/bin/bash
# end-to-end leg (host buffers in, host records out) under a few settings of the upload path; one line per setting
# usage (via gpurun): bash scripts/gpu_e2e_sweep.sh "UZ_LAYOUT_APART=0 UZ_COPY_CTAS=24" "UZ_LAYOUT_APART=1 UZ_COPY_CTAS=32" ...
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-places --no-extras --steps 5 2>/dev/null | \
    python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('$cfg', 'value', d['value'], 'e2e pinned', d['e2e']['pinned'], 'pageable', d['e2e']['pageable'], 'ok_frac', d['sanity']['ok_frac'])"
done
