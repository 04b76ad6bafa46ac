import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from uzliti_slam_b200 import EdgeEstimator, synthetic as S
est = EdgeEstimator(0)
for n in (500, 1000):
    f, t, _ = S.make_pair(n, seed=1)
    h = est.add_keyframes([f, t])
    for _ in range(20): est.estimateEdges(h[:1], h[1:2])
    lat = []
    for _ in range(200):
        t0 = time.perf_counter(); est.estimateEdges(h[:1], h[1:2]); lat.append(time.perf_counter() - t0)
    est.enable_timers(True); est.reset_timers()
    for _ in range(50): est.estimateEdges(h[:1], h[1:2])
    tm = est.get_timers(); est.enable_timers(False)
    print(n, "wall us", round(np.median(lat) * 1e6, 1), "match us", round(tm["match_ms"] / 50 * 1e3, 1), "solve us", round(tm["solve_ms"] / 50 * 1e3, 1))
