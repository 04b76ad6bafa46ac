"""GPU probe: knn2_wide_kernel (64-byte BRISK / FREAK rows) on a loop-closure batch, both CTA shapes, against the POPC
ceiling measured in the same process.  Usage (GPU box): python scripts/gpu_wide_probe.py [n_keyframes]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from uzliti_slam_b200 import EdgeEstimator, synthetic as S  # noqa: E402

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 400
kfs, pairs, _ = S.make_map(n_kf, n_features=1000, k_candidates=20, seed=77, desc_bytes=64)
out = {}
for cfg in ("auto", "0", "1"):
    if cfg == "auto":
        os.environ.pop("UZ_KNN_WIDE_CFG", None)
    else:
        os.environ["UZ_KNN_WIDE_CFG"] = cfg
    for stream in ("0", "1"):
        os.environ["UZ_STREAM_SOLVE"] = stream
        est = EdgeEstimator(0)
        h = est.add_keyframes(kfs)
        est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        est.enable_timers(True)
        est.reset_timers()
        t0 = time.perf_counter()
        for _ in range(3):
            r = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        dt = (time.perf_counter() - t0) / 3
        tm = est.get_timers()
        popc = est.microbench(0)
        g = tm["compares"] / (tm["match_ms"] * 1e-3) * 1e-9
        out[f"cfg_{cfg}_stream_{stream}"] = dict(pairs=len(pairs), edges_per_s=round(len(pairs) / dt, 1), knn2_wide_ms=round(tm["match_ms"] / 3, 3),
                                                 solve_ms=round(tm["solve_ms"] / 3, 3), gcmp512_per_s=round(g, 1),
                                                 frac_of_popc_ceiling=round(g * 8 / popc, 4), ok=int(r["ok"].sum()),
                                                 median_consensus=float(np.median(r["consensus"])))
        est.close()
print(json.dumps(out, indent=1))
