"""Measurement helper: how the streaming solve and the match kernel share the SMs.
UZ_STREAM_PROBE (measurement only): 1 = streaming grid resident but idle, 2 = stop after the gather, 3 = stop after
hypothesis scoring.  Usage: python scripts/gpu_stream_probe.py [forms...], a form is "<ctas per SM>[:probe]"."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uzliti_slam_b200 import EdgeEstimator, synthetic as S

kfs, pairs, _ = S.make_map(1250, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=20, cross_cluster=4, seed=4)
sel = pairs[:25000]
for form in (sys.argv[1:] or ["0", "1", "2"]):
    n, _, probe = form.partition(":")
    for k in ("UZ_STREAM_SOLVE", "UZ_STREAM_PROBE"):
        os.environ.pop(k, None)
    os.environ["UZ_STREAM_SOLVE"] = n
    if probe:
        os.environ["UZ_STREAM_PROBE"] = probe
    est = EdgeEstimator(0)
    h = est.add_keyframes(kfs)
    for _ in range(2):
        est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
    est.enable_timers(True); est.reset_timers()
    for _ in range(4):
        est.estimateEdges(h[sel[:, 0]], h[sel[:, 1]])
    t = est.get_timers()
    print(f"stream ctas/SM {n} probe {probe or '-'}: match {t['match_ms'] / 4:.2f} ms  solve span {t['solve_ms'] / 4:.2f} ms", flush=True)
    est.close()
