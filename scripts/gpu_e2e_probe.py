"""e2e leg (uz_estimate_edges_host from pinned / pageable buffers) under different knobs; host stage times with UZ_TRACE=1.
usage: python scripts/gpu_e2e_probe.py [pairs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
    import torch
    desc = np.empty((10000, 1000, 32), np.uint8); pos = np.empty((10000, 1000, 3), np.float64); valid = np.empty((10000, 1000), np.uint8)
    kfs, pairs, _ = bench.build_map(10000, out=(desc, pos, valid))
    td, tp, tv = torch.from_numpy(desc).pin_memory(), torch.from_numpy(pos).pin_memory(), torch.from_numpy(valid).pin_memory()
    pinned = [dict(kf, desc=td.numpy()[i], pos=tp.numpy()[i], valid=tv.numpy()[i]) for i, kf in enumerate(kfs)]
    my = pairs[:n_pairs]
    knobs = [{}, {"UZ_HOST_SLOTS": "2"}, {"UZ_HOST_SLOTS": "3"}, {"UZ_HOST_SLOTS": "4"}, {"UZ_COPY_CTAS": "32"}, {"UZ_COPY_CTAS": "148"},
             {"UZ_HOST_CHUNKS": "8"}, {"UZ_HOST_CHUNKS": "32"}, {"UZ_ALT_CHUNKS": "0"}, {"UZ_TRACE": "2"}]
    if os.environ.get("UZ_PROBE_KNOBS"):
        import json
        knobs = json.loads(os.environ["UZ_PROBE_KNOBS"])
    for env in knobs:
        est = bench._new_estimator(0, **env)
        os.environ.update(env)
        for name, src in (("pinned", pinned), ("pageable", kfs)):
            prep = est.prepare_host_pairs([([src[a]], [src[b]]) for a, b in my])
            est.estimateEdgesHostPrepared(prep); est.estimateEdgesHostPrepared(prep)
            t0 = time.perf_counter()
            for _ in range(4):
                est.estimateEdgesHostPrepared(prep)
            dt = (time.perf_counter() - t0) / 4
            print(f"{env} {name}: {dt * 1e3:.2f} ms/step = {n_pairs / dt / 1e3:.0f} k edges/s", flush=True)
        for k in env:
            os.environ.pop(k, None)
        est.close()


if __name__ == "__main__":
    main()
