import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uzliti_slam_b200 import EdgeEstimator, synthetic as S
est = EdgeEstimator(0)
kfs, pairs, _ = S.make_map(500, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=20, cross_cluster=4, seed=4)
h = est.add_keyframes(kfs)
sel = pairs[:4000]
est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
est.set_debug(True)
res = est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
ph = np.array([est.debug_phases(i) for i in range(0, 4000, 10)])
d = np.diff(ph, axis=1)
names = ['ratio+keys', 'sort', 'count+gather', 'hyp solve', 'score+scan', 'compact+refit', 'final+mse']
same = (sel[::10, 0] // 25) == (sel[::10, 1] // 25)
for nm, col in zip(names, d.T):
    print(f'{nm:15s} true-pairs {col[same].mean():9.0f} cyc   false-pairs {col[~same].mean():9.0f} cyc')
print('total', (ph[:, 7] - ph[:, 0])[same].mean(), (ph[:, 7] - ph[:, 0])[~same].mean())
est.set_debug(False)
est.enable_timers(True); est.reset_timers()
for _ in range(3): est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
print(est.get_timers())
