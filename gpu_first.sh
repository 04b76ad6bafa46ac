set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np
from uzliti_slam_b200 import EdgeEstimator, synthetic as S
est = EdgeEstimator(0)
for op, name in enumerate(['POPC', 'LOP3', 'IMAD', 'VIMNMX']):
    print('microbench', name, est.microbench(op), 'Gop/s')
PY
python __graft_entry__.py smoke
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
