"""GPU probe: cost of the cross-check (uz_params.cross_check = 1) in its two forms - fused into the forward match kernel
(column minima beside the row top-2, the default) and as a second, reversed matching (UZ_XCHECK_FUSED=0) - against the
plain path, on 256-bit and 512-bit rows.  Both forms must return identical records.
Usage (GPU box): python scripts/gpu_xcheck_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from uzliti_slam_b200 import EdgeEstimator, synthetic as S  # noqa: E402

out = {}
for nb in (32, 64):
    kfs, pairs, _ = S.make_map(300, n_features=1000, k_candidates=20, seed=5, desc_bytes=nb)
    sel = np.concatenate([pairs] * 3)[:18000] if nb == 32 else pairs
    recs = {}
    for name, cross, fused in (("plain", 0, 1), ("fused", 1, 1), ("reversed", 1, 0)):
        os.environ["UZ_XCHECK_FUSED"] = str(fused)
        os.environ["UZ_STREAM_SOLVE"] = "0"
        est = EdgeEstimator(0)
        est.setConfig(cross_check=cross)
        h = est.add_keyframes(kfs)
        recs[name] = est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
        est.enable_timers(True)
        est.reset_timers()
        for _ in range(3):
            est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
        t = est.get_timers()
        out[f"{nb * 8}bit_{name}"] = dict(pairs=len(sel), match_ms=round(t["match_ms"] / 3, 3), solve_ms=round(t["solve_ms"] / 3, 3),
                                          mean_ratio_matches=float(recs[name]["n_ratio_matches"].mean()))
        est.close()
    out[f"{nb * 8}bit_fused_equals_reversed"] = bool(recs["fused"].tobytes() == recs["reversed"].tobytes())
    base = out[f"{nb * 8}bit_plain"]["match_ms"]
    for name in ("fused", "reversed"):
        out[f"{nb * 8}bit_{name}"]["match_cost_vs_plain"] = round(out[f"{nb * 8}bit_{name}"]["match_ms"] / base, 3)
print(json.dumps(out, indent=1))
