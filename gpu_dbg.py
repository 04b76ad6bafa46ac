import sys
sys.path.insert(0, '.')
import numpy as np
from uzliti_slam_b200 import EdgeEstimator, synthetic as S
from oracle import binding as O
est = EdgeEstimator(0)
f, t, _ = S.make_pair(500, seed=1)
est.set_debug(True)
r = est.estimateEdgeDirect([f], [t])
o = O.estimate_edge([f], [t])
m, mask = est.debug_pair(0, r["n_matches"])
print('matches equal', np.array_equal(m, o['matches']), 'mask equal', np.array_equal(mask, o['inlier_mask']))
if not np.array_equal(m, o['matches']):
    bad = np.flatnonzero((m != o['matches']).any(1)); print(bad[:20]); print(m[bad[:10]]); print(o['matches'][bad[:10]])
c = est.debug_counts(0)
print('counts equal', np.array_equal(c, o['counts']))
bad = np.flatnonzero(c != o['counts']); print(bad, c[bad], o['counts'][bad])
print('best', r['best_iteration'], o['best_iteration'])
# direct mode on oracle's P,Q
P = t['pos'][o['matches'][:, 0]]; Q = f['pos'][o['matches'][:, 1]]
g = est.estimateSVD(P, Q, 0.1, 100, 0.6)
oo = O.estimate_svd(P, Q, 0.1, 100, 0.6)
print('direct best', g['best_iteration'], oo['best_iteration'], np.array_equal(g['T'], oo['T']))
sl = est.sample_list(len(P), 100); print('samples', sl[[14, 46]])
for h in bad[:5]:
    s = sl[h]
    print(h, 'oracle T\n', O.pose_svd(P[s], Q[s]))
