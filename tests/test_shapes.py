"""Every CTA shape of the match kernel (UZ_KNN_CFG: 0 = 256x4, 1 = 128x4, 2 = 64x2, 3 = 256x2, 4 = 128x2, 5 = 32x2 queries
per CTA; normally chosen per batch by the padding cost model) must return exactly the oracle's neighbours on sizes that
leave ragged tiles, partial 128-row key blocks and partial train stages."""
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SIZES = [(1000, 1000), (400, 300), (65, 129), (1, 5), (513, 4096), (130, 127), (64, 2)]


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5])
def test_every_cta_shape_matches_the_oracle(oracle, cfg):
    from uzliti_slam_b200 import EdgeEstimator
    old = os.environ.get("UZ_KNN_CFG")
    os.environ["UZ_KNN_CFG"] = str(cfg)
    try:
        est = EdgeEstimator(0)
    finally:
        if old is None:
            os.environ.pop("UZ_KNN_CFG", None)
        else:
            os.environ["UZ_KNN_CFG"] = old
    try:
        for nq, nt in SIZES:
            rng = np.random.default_rng(1000 * cfg + nq + nt)
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
            if nt > 200:                       # ties across key blocks and stages: duplicates of query rows far apart
                t[5] = q[0]; t[nt - 3] = q[0]; t[130] = q[nq // 2]
            idx, dist = est.knnMatch(q, t)
            oi, od = oracle.knn2(q, t)
            assert np.array_equal(idx, oi), (cfg, nq, nt)
            assert np.array_equal(dist, od), (cfg, nq, nt)
    finally:
        est.close()


@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_cross_check_fused_and_reversed_forms_agree_with_the_oracle(oracle, cfg):
    """uz_params.cross_check in both forms: column minima tracked by the forward match kernel (default) and a second,
    reversed matching (UZ_XCHECK_FUSED=0), for every two-queries-per-thread CTA shape, on ragged tiles (clamped duplicate
    queries in the last tile), tie-heavy rows (lowest query index must win a column), a rig and both descriptor widths."""
    from uzliti_slam_b200 import EdgeEstimator
    ests = []
    for fused in (1, 0):
        os.environ["UZ_KNN_CFG"] = str(cfg)
        os.environ["UZ_KNN_WIDE_CFG"] = "0" if cfg in (3, 4) else "1"
        os.environ["UZ_XCHECK_FUSED"] = str(fused)
        os.environ["UZ_STREAM_SOLVE"] = "0"
        try:
            e = EdgeEstimator(0)
            e.setConfig(cross_check=1)
            ests.append(e)
        finally:
            for k in ("UZ_KNN_CFG", "UZ_KNN_WIDE_CFG", "UZ_XCHECK_FUSED", "UZ_STREAM_SOLVE"):
                os.environ.pop(k, None)
    try:
        cases = []
        for nb in (32, 64):
            for kw in (dict(n_from=1000, seed=1), dict(n_from=700, n_to=333, seed=2), dict(n_from=129, n_to=577, seed=3),
                       dict(n_from=300, seed=4, tie_stress=True), dict(n_from=65, n_to=31, seed=5, tie_stress=True),
                       dict(n_from=513, n_to=1025, seed=6)):
                f, t, _ = S.make_pair(desc_bytes=nb, **kw)
                cases.append(([f], [t]))
        fa, ta, _ = S.make_pair(400, 300, seed=7, sensor_frame=0, tie_stress=True)
        fb, tb, _ = S.make_pair(350, 450, seed=8, sensor_frame=1)
        cases.append(([fa, fb], [ta, tb]))
        got = [e.estimateEdgesHost(cases) for e in ests]
        assert got[0].tobytes() == got[1].tobytes()
        dropped = 0
        for r, (cf, ct) in zip(got[0], cases):
            o = oracle.estimate_edge(cf, ct, cross_check=True)
            assert r["n_ratio_matches"] == o["n_ratio_matches"] and r["n_matches"] == o["n_matches"]
            assert r["consensus"] == o["consensus"] and r["cam_from"] == o["cam_from"]
            assert np.abs(r["T"].reshape(4, 4) - o["T"]).max() < 1e-5
            dropped += oracle.estimate_edge(cf, ct)["n_ratio_matches"] - o["n_ratio_matches"]
        assert dropped > 500
    finally:
        for e in ests:
            e.close()


@pytest.mark.parametrize("nb", [32, 64])
def test_random_shapes_fuzz(est, oracle, nb):
    """seeded random (nq, nt) around the tile / key-block / stage boundaries, with and without the fused cross-check:
    neighbours through uz_match_knn2, ratio survivors and match counts through the whole path"""
    rng = np.random.default_rng(1234 + nb)
    edges = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 767, 1023, 1024, 1025]
    try:
        for it in range(36):
            nq = int(rng.choice(edges)) if it % 3 else int(rng.integers(1, 1400))
            nt = int(rng.choice(edges)) if it % 2 else int(rng.integers(1, 1400))
            keep = int(rng.choice([1, 2, nb]))
            q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, nb), dtype=np.uint8)
            q[:, keep:] = 0
            t[:, keep:] = 0
            idx, dist = est.knnMatch(q, t)
            oi, od = oracle.knn2(q, t)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nb, nq, nt, keep)
            if nq >= 7 and nt >= 7:
                cam = lambda d: dict(desc=d, pos=rng.normal(size=(len(d), 3)) + [0, 0, 3], valid=np.ones(len(d), np.uint8),
                                     feature_type=2, sensor_frame=0)
                cf, ct = cam(t), cam(q)
                for cross in (0, 1):
                    est.setConfig(cross_check=cross)
                    r = est.estimateEdgeDirect([cf], [ct])
                    o = oracle.estimate_edge([cf], [ct], cross_check=bool(cross))
                    assert r["n_ratio_matches"] == o["n_ratio_matches"] and r["n_matches"] == o["n_matches"], (nb, nq, nt, keep, cross)
                    assert r["consensus"] == o["consensus"] and bool(r["ok"]) == o["ok"], (nb, nq, nt, keep, cross)
    finally:
        est.setConfig(cross_check=0)
