"""Every CTA shape of the match kernel (UZ_KNN_CFG: 0 = 256x4, 1 = 128x4, 2 = 64x2, 3 = 256x2, 4 = 128x2, 5 = 32x2 queries
per CTA; normally chosen per batch by the padding cost model) must return exactly the oracle's neighbours on sizes that
leave ragged tiles, partial 128-row key blocks and partial train stages."""
import os

import numpy as np
import pytest

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
