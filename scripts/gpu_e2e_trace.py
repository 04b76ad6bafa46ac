"""host stage times of the e2e leg (UZ_TRACE=1)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch
desc = np.empty((10000, 1000, 32), np.uint8); pos = np.empty((10000, 1000, 3), np.float64); valid = np.empty((10000, 1000), np.uint8)
kfs, pairs, _ = bench.build_map(10000, out=(desc, pos, valid))
td, tp, tv = torch.from_numpy(desc).pin_memory(), torch.from_numpy(pos).pin_memory(), torch.from_numpy(valid).pin_memory()
pinned = [dict(kf, desc=td.numpy()[i], pos=tp.numpy()[i], valid=tv.numpy()[i]) for i, kf in enumerate(kfs)]
my = pairs[:25000]
os.environ["UZ_TRACE"] = os.environ.get("UZ_TRACE", "1")
from uzliti_slam_b200 import EdgeEstimator
est = EdgeEstimator(0)
for name, src in (("pinned", pinned), ("pageable", kfs)):
    prep = est.prepare_host_pairs([([src[a]], [src[b]]) for a, b in my])
    for _ in range(4):
        t0 = time.perf_counter(); est.estimateEdgesHostPrepared(prep); print(name, (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr)
