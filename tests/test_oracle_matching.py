"""CPU: pins the oracle's matching stage (K1, K2 predicate) — against the live OpenCV matcher the reference
calls (cv2.BFMatcher(NORM_HAMMING).knnMatch, feature_transformation_estimator.cpp:38,58) and against the
literal float/double ratio expression of :67."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

cv2 = pytest.importorskip("cv2")


def _cv2_knn(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1, np.int32)
    for i, row in enumerate(m):
        for j, d in enumerate(row):
            idx[i, j] = d.trainIdx
            dist[i, j] = int(d.distance)
    return idx, dist


@pytest.mark.parametrize("nq,nt", [(500, 500), (1000, 1000), (257, 511), (64, 2000), (3, 2), (5, 1), (1, 1)])
def test_oracle_knn2_matches_cv2(oracle, nq, nt):
    rng = np.random.default_rng(nq + 13 * nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    oi, od = oracle.knn2(q, t)
    ci, cd = _cv2_knn(q, t)
    assert np.array_equal(oi, ci) and np.array_equal(od, cd)


@pytest.mark.parametrize("keep", [1, 2, 4])
def test_oracle_tie_rule_is_lowest_train_index(oracle, keep):
    """Hamming distances are small integers, ties are the common case: OpenCV orders by (distance, trainIdx)."""
    rng = np.random.default_rng(keep)
    q = rng.integers(0, 256, (800, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1200, 32), dtype=np.uint8)
    q[:, keep:] = 0
    t[:, keep:] = 0
    t[700] = t[5]                                       # exact duplicates
    oi, od = oracle.knn2(q, t)
    ci, cd = _cv2_knn(q, t)
    assert (od[:, 0] == od[:, 1]).sum() > 100
    assert np.array_equal(oi, ci) and np.array_equal(od, cd)


@pytest.mark.parametrize("nq,nt,keep", [(500, 500, 64), (1000, 700, 64), (129, 65, 64), (300, 400, 33), (400, 300, 2),
                                        (3, 2, 64), (1, 1, 64)])
def test_oracle_knn2_matches_cv2_on_512_bit_rows(oracle, nq, nt, keep):
    """BRISK / FREAK rows are 64 bytes (feature_extraction_core.cpp:69-77); same matcher, same tie rule, distances to 512."""
    rng = np.random.default_rng(nq + 17 * nt + keep)
    q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 64), dtype=np.uint8)
    q[:, keep:] = 0
    t[:, keep:] = 0
    if nt > 1:
        t[-1] = 255 - q[0]                            # one exact complement: distance 512 when keep == 64
    oi, od = oracle.knn2(q, t)
    ci, cd = _cv2_knn(q, t)
    assert np.array_equal(oi, ci) and np.array_equal(od, cd)
    assert od.max() <= 512


def test_oracle_synthetic_pair_matches_cv2(oracle):
    f, t, _ = S.make_pair(1000, seed=5)
    oi, od = oracle.knn2(t["desc"], f["desc"])
    ci, cd = _cv2_knn(t["desc"], f["desc"])
    assert np.array_equal(oi, ci) and np.array_equal(od, cd)


def test_ratio_predicate_equals_integer_form(oracle):
    """d0 < 0.99*d1 in float/double (reference :67) == 100*d0 < 99*d1 for every reachable distance pair."""
    for d1 in range(0, 513):                             # 512: BRISK / FREAK rows
        for d0 in range(0, d1 + 1):
            assert oracle.ratio_pass(d0, d1) == (100 * d0 < 99 * d1), (d0, d1)


def _cv2_cross(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    idx = np.full(len(q), -1, np.int32)
    dist = np.full(len(q), -1, np.int32)
    for d in m:
        idx[d.queryIdx] = d.trainIdx
        dist[d.queryIdx] = int(d.distance)
    return idx, dist


@pytest.mark.parametrize("nq,nt,keep", [(500, 500, 32), (300, 700, 32), (700, 300, 32), (400, 600, 2), (600, 400, 1),
                                        (5, 1, 32), (1, 5, 32), (300, 500, 64), (500, 300, 3)])
def test_oracle_cross_check_matches_cv2(oracle, nq, nt, keep):
    """The opt-in cross-check (uz_params.cross_check) is OpenCV's crossCheck matcher, tie rules included."""
    rng = np.random.default_rng(7 * nq + nt + keep)
    nb = 64 if keep in (64, 3) else 32
    q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, nb), dtype=np.uint8)
    q[:, keep:] = 0
    t[:, keep:] = 0
    oi, od = oracle.cross_match(q, t)
    ci, cd = _cv2_cross(q, t)
    assert np.array_equal(oi, ci) and np.array_equal(od, cd)


def test_oracle_cross_check_edge_is_a_subset_filter(oracle):
    """edge-level definition: a ratio survivor (q, t) stays iff the crossCheck matcher returns (q, t)"""
    f, t, _ = S.make_pair(600, seed=77, tie_stress=True)
    a = oracle.estimate_edge([f], [t])
    b = oracle.estimate_edge([f], [t], cross_check=True)
    ci, _ = _cv2_cross(t["desc"], f["desc"])
    keep = [tuple(m) for m in a["matches"] if ci[m[0]] == m[1]]
    assert 0 < len(keep) < len(a["matches"])
    assert sorted(keep) == sorted(tuple(m) for m in b["matches"])
