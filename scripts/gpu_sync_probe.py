import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
kfs, pairs, _ = bench.build_map(10000)
for env in ({"UZ_PIPELINE_CALLS": "0"}, {}):
    est = bench._new_estimator(0, **env)
    h = est.add_keyframes(kfs)
    f, t = h[pairs[:25000, 0]], h[pairs[:25000, 1]]
    est.estimateEdges(f, t); est.estimateEdges(f, t)
    t0 = time.perf_counter()
    for _ in range(5): r = est.estimateEdges(f, t)
    dt = (time.perf_counter() - t0) / 5
    print(env, "synchronous uz_estimate_edges, 25000 pairs: %.3f ms = %.0f k edges/s" % (dt * 1e3, 25000 / dt / 1e3), flush=True)
    est.close()
