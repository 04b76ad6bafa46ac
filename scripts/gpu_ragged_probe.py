"""Measurement helper: match-kernel throughput per CTA shape (UZ_KNN_CFG) on uniform and ragged camera sizes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uzliti_slam_b200 import EdgeEstimator, synthetic as S

for n_feat in (1000, 400, 300, 700):
    kfs, pairs, _ = S.make_map(300, n_features=n_feat, cluster=25, pool=n_feat, n_shared=int(0.6 * n_feat), k_candidates=20, cross_cluster=4, seed=4)
    sel = np.concatenate([pairs] * 3)[:18000]
    for cfg in ("auto", "3", "4", "2", "5"):
        os.environ.pop("UZ_KNN_CFG", None)
        if cfg != "auto":
            os.environ["UZ_KNN_CFG"] = cfg
        os.environ["UZ_STREAM_SOLVE"] = "0"
        est = EdgeEstimator(0)
        h = est.add_keyframes(kfs)
        est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
        est.enable_timers(True); est.reset_timers()
        for _ in range(3):
            est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
        t = est.get_timers()
        print(f"N={n_feat:5d} cfg {cfg:>4s}: match {t['match_ms'] / 3:7.3f} ms  {t['compares'] / t['match_ms'] * 1e-6:7.1f} Gcmp/s (useful)  solve {t['solve_ms'] / 3:.3f} ms", flush=True)
        est.close()
