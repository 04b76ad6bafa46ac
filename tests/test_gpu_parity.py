"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle, bit-exact for integer work
(match indices/distances, sorted match lists, inlier masks, consensus) and within 1e-5 m / 1e-5 rad for
transforms (BASELINE.json north_star) — in practice the transforms are bit-identical too.
"""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu

TOL_T = 1e-5      # metres   (north_star)
TOL_R = 1e-5      # radians  (north_star), measured with the atan2 metric


def _check_edge(r, o, tag=""):
    assert bool(r["ok"]) == o["ok"], tag
    assert r["cam_from"] == o["cam_from"] and r["cam_to"] == o["cam_to"], tag
    assert r["n_ratio_matches"] == o["n_ratio_matches"], tag
    assert r["n_matches"] == o["n_matches"], tag
    assert r["consensus"] == o["consensus"], tag
    assert r["best_iteration"] == o["best_iteration"], tag
    assert r["iterations_run"] == o["iterations_run"], tag
    T = r["T"].reshape(4, 4)
    assert np.linalg.norm(T[:3, 3] - o["T"][:3, 3]) <= TOL_T, tag
    assert S.rot_angle(T[:3, :3], o["T"][:3, :3]) <= TOL_R, tag
    if np.isnan(o["mse"]):
        assert np.isnan(r["mse"]), tag
    else:
        assert abs(r["mse"] - o["mse"]) <= 1e-12 * max(1.0, abs(o["mse"])), tag
        assert abs(r["info_scale"] - o["info_scale"]) <= 1e-9 * max(1.0, abs(o["info_scale"])), tag


# ------------------------------------------------------------------------------------------------
# K1: kNN-2 matching — bit-exact with the oracle (which is pinned to cv2.BFMatcher)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt", [(500, 500), (1000, 1000), (2000, 2000), (1, 1), (1, 2), (7, 3), (33, 1),
                                   (257, 511), (1000, 513), (1025, 1), (64, 4097), (3, 1000)])
def test_knn2_random(est, oracle, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)


def test_knn2_empty_train(est):
    q = np.zeros((5, 32), np.uint8)
    idx, dist = est.knnMatch(q, np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (dist == -1).all()


@pytest.mark.parametrize("keep_bytes", [1, 2, 4])
def test_knn2_tie_stress(est, oracle, keep_bytes):
    """low-entropy descriptors: most rows tie on distance; lowest train index must win for both neighbours"""
    rng = np.random.default_rng(keep_bytes)
    q = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    q[:, keep_bytes:] = 0
    t[:, keep_bytes:] = 0
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(q, t)
    assert (od[:, 0] == od[:, 1]).sum() > 100
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_knn2_duplicates_and_extremes(est, oracle):
    rng = np.random.default_rng(5)
    t = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    t[100] = t[3]; t[599] = t[3]; t[0] = 0; t[1] = 255       # duplicate rows, all-zero, all-one rows
    q = np.concatenate([t[[3, 100, 0, 1]], 255 - t[:50], rng.integers(0, 256, (40, 32), dtype=np.uint8)])
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    assert idx[0, 0] == 3 and idx[0, 1] == 100 and dist[0, 0] == 0 and dist[0, 1] == 0
    assert dist.max() <= 256


def test_knn2_strided_rows(est, oracle):
    """cv::Mat rows may be padded: row stride 48 bytes"""
    rng = np.random.default_rng(9)
    qb = rng.integers(0, 256, (300, 48), dtype=np.uint8)
    tb = rng.integers(0, 256, (400, 40), dtype=np.uint8)
    q, t = qb[:, :32], tb[:, :32]
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(np.ascontiguousarray(q), np.ascontiguousarray(t))
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_knn2_matches_cv2(est):
    cv2 = pytest.importorskip("cv2")
    f, t, _ = S.make_pair(1000, seed=11)
    idx, dist = est.knnMatch(t["desc"], f["desc"])
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(t["desc"], f["desc"], k=2)
    ci = np.array([[a.trainIdx, b.trainIdx] for a, b in m])
    cd = np.array([[a.distance, b.distance] for a, b in m]).astype(np.int32)
    assert np.array_equal(idx, ci) and np.array_equal(dist, cd)


# ------------------------------------------------------------------------------------------------
# K3..K5: estimateSVD / consensus3D on raw point sets
# ------------------------------------------------------------------------------------------------
def _point_problem(rng, M, inlier_frac, noise=0.01):
    T = S._rand_pose(rng, 30.0, 1.5)
    P = S._landmarks(rng, M)
    Q = P @ T[:3, :3].T + T[:3, 3] + rng.normal(size=(M, 3)) * noise
    out = rng.random(M) > inlier_frac
    Q[out] = S._landmarks(rng, int(out.sum()))
    return P, Q, T


@pytest.mark.parametrize("M,frac,iters,bp,prosac", [
    (300, 0.5, 100, 0.6, True), (300, 0.8, 100, 0.6, True), (57, 0.4, 100, 0.6, True), (3, 1.0, 100, 0.6, True),
    (4, 0.5, 10, 0.6, True), (1000, 0.3, 1000, 0.6, True), (100, 0.7, 200, 1.0, False), (100, 0.05, 50, 0.6, True),
    (2000, 0.6, 300, 0.9, True), (129, 0.5, 129, 0.55, False)])
def test_estimate_svd(est, oracle, M, frac, iters, bp, prosac):
    rng = np.random.default_rng(M * 31 + iters)
    P, Q, _ = _point_problem(rng, M, frac)
    g = est.estimateSVD(P, Q, 0.1, iters, bp, prosac)
    o = oracle.estimate_svd(P, Q, 0.1, iters, bp, prosac)
    assert g["consensus"] == o["consensus"]
    assert g["best_iteration"] == o["best_iteration"] and g["iterations_run"] == o["iterations_run"]
    assert np.array_equal(g["mask"], o["mask"])
    assert np.array_equal(g["T"], o["T"])                      # bit-identical, stronger than the 1e-5 bar
    assert g["mse"] == o["mse"] or (np.isnan(g["mse"]) and np.isnan(o["mse"]))


def test_estimate_svd_custom_samples_and_transformation_filter_call(est, oracle):
    """TransformationFilter::calcValidEdges calls estimateSVD(P,Q,...,0.3,200,1.0,false) on <=100 points
    (transformation_filter.cpp:272)"""
    rng = np.random.default_rng(3)
    P, Q, _ = _point_problem(rng, 100, 0.7, noise=0.05)
    g = est.estimateSVD(P, Q, 0.3, 200, 1.0, False)
    o = oracle.estimate_svd(P, Q, 0.3, 200, 1.0, False)
    assert g["consensus"] == o["consensus"] and np.array_equal(g["mask"], o["mask"]) and np.array_equal(g["T"], o["T"])
    samples = rng.integers(0, 100, (64, 3)).astype(np.int32)
    g = est.estimateSVD(P, Q, 0.3, 64, 0.6, True, samples=samples)
    o = oracle.estimate_svd(P, Q, 0.3, 64, 0.6, True, samples=samples)
    assert g["consensus"] == o["consensus"] and np.array_equal(g["mask"], o["mask"]) and np.array_equal(g["T"], o["T"])


def test_estimate_svd_degenerate(est, oracle):
    rng = np.random.default_rng(4)
    # all points identical / collinear / fewer than 3
    P = np.tile(rng.normal(size=(1, 3)), (20, 1)); Q = P + 1.0
    for Pi, Qi in ((P, Q), (np.outer(np.arange(30.0), [1, 2, 3]), np.outer(np.arange(30.0), [3, 2, 1]))):
        g = est.estimateSVD(Pi, Qi, 0.1, 50, 0.6)
        o = oracle.estimate_svd(Pi, Qi, 0.1, 50, 0.6)
        assert g["consensus"] == o["consensus"] and np.array_equal(g["mask"], o["mask"])
        assert np.array_equal(g["T"], o["T"])
    g = est.estimateSVD(P[:2], Q[:2], 0.1, 50, 0.6)
    assert g["consensus"] == 0 and np.array_equal(g["T"], np.eye(4))


def test_sample_list_matches_real_random_shuffle(est, oracle):
    for M, I, pros in [(3, 100, True), (300, 100, True), (557, 100, True), (100, 200, False), (1000, 1000, True)]:
        assert np.array_equal(est.sample_list(M, I, pros), oracle.sample_list(M, I, pros))


def test_consensus3d(est, oracle):
    rng = np.random.default_rng(8)
    P, Q, T = _point_problem(rng, 777, 0.5)
    for thr in (0.1, 0.03, 1e-9, 10.0):
        c, m = est.consensus3D(P, Q, T, thr)
        oc, om = oracle.consensus3d(P, Q, T, thr)
        assert c == oc and np.array_equal(m, om)


def test_consensus_threshold_is_exact_at_the_boundary(est, oracle):
    """residuals that equal the threshold up to 1 ulp: sqrt(s) < thr must agree with the oracle's sqrt form"""
    P = np.zeros((64, 3)); T = np.eye(4)
    thr = 0.1
    d = np.array([np.nextafter(thr, 0), thr, np.nextafter(thr, 1)] + list(thr * (1 + np.arange(-30, 31) * 2.2e-16)))
    Q = np.zeros((64, 3)); Q[:, 0] = d
    c, m = est.consensus3D(P, Q, T, thr)
    oc, om = oracle.consensus3d(P, Q, T, thr)
    assert c == oc and np.array_equal(m, om)


# ------------------------------------------------------------------------------------------------
# full edge estimates: store path, host path, debug taps
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,seed,kw", [
    (500, 1, {}), (1000, 2, {}), (1000, 3, dict(tie_stress=True)), (300, 4, dict(rho=0.9)),
    (400, 5, dict(rho=0.05)), (2000, 6, {}), (1000, 7, dict(invalid_frac=0.9)), (64, 8, dict(rho=0.8)),
    (1000, 9, dict(gross_outlier_frac=0.5)), (777, 10, dict(noise=False))])
def test_edge_host_path(est, oracle, n, seed, kw):
    f, t, Tgt = S.make_pair(n, seed=seed, **kw)
    est.set_debug(True)
    r = est.estimateEdgeDirect([f], [t])
    o = oracle.estimate_edge([f], [t])
    _check_edge(r, o, f"n={n} seed={seed}")
    m, mask = est.debug_pair(0, r["n_matches"])
    assert np.array_equal(m, o["matches"])                     # sorted final_matches (q, t, d)
    if o["ok"]:
        assert np.array_equal(mask, o["inlier_mask"])
        assert np.array_equal(est.debug_counts(0), o["counts"])   # consensus of EVERY hypothesis
    est.set_debug(False)


def test_edge_maximum_size(est, oracle):
    """UZ_MAX_FEATURES = 4096 keypoints per camera: the largest pair the on-chip solve holds (cap 4096, one CTA per SM),
    with and without depth gaps; one more keypoint is refused with an error, not truncated."""
    for seed, kw in ((70, dict(rho=0.9, invalid_frac=0.0)), (71, dict(rho=0.6))):
        f, t, _ = S.make_pair(4096, seed=seed, **kw)
        est.set_debug(True)
        r = est.estimateEdgeDirect([f], [t])
        o = oracle.estimate_edge([f], [t])
        _check_edge(r, o, f"max size seed={seed}")
        assert r["n_matches"] > 2000
        m, mask = est.debug_pair(0, r["n_matches"])
        assert np.array_equal(m, o["matches"])
        assert np.array_equal(mask, o["inlier_mask"])
        assert np.array_equal(est.debug_counts(0), o["counts"])
        est.set_debug(False)
    f, t, _ = S.make_pair(4097, seed=72)
    with pytest.raises(Exception, match="feature count out of range"):
        est.estimateEdgeDirect([f], [t])
    # the context is still usable after the refusal
    f, t, _ = S.make_pair(100, seed=73, rho=0.8)
    _check_edge(est.estimateEdgeDirect([f], [t]), oracle.estimate_edge([f], [t]), "after refusal")


def test_edge_recovers_ground_truth(est):
    f, t, Tgt = S.make_pair(1000, seed=21)
    r = est.estimateEdgeDirect([f], [t])
    T = r["T"].reshape(4, 4)
    assert r["ok"] and r["consensus"] > 100
    assert np.linalg.norm(T[:3, 3] - Tgt[:3, 3]) < 0.05 and S.rot_angle(T[:3, :3], Tgt[:3, :3]) < 0.02


def test_edge_store_batch(est, oracle):
    kfs, pairs, _ = S.make_map(40, n_features=600, cluster=8, pool=600, n_shared=350, k_candidates=6,
                               cross_cluster=2, seed=3)
    est.clear()
    h = est.add_keyframes(kfs)
    assert est.store_size() == 40
    est.set_debug(True)
    res = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    for i, (a, b) in enumerate(pairs):
        o = oracle.estimate_edge([kfs[a]], [kfs[b]])
        _check_edge(res[i], o, f"pair {i} ({a},{b})")
        m, mask = est.debug_pair(i, res[i]["n_matches"])
        assert np.array_equal(m, o["matches"])
        if o["ok"]:
            assert np.array_equal(mask, o["inlier_mask"])
    est.set_debug(False)
    # removing a keyframe invalidates its handle, not the others
    est.remove_keyframe(int(h[0]))
    with pytest.raises(Exception):
        est.estimateEdges([h[0]], [h[1]])
    r = est.estimateEdges([h[1]], [h[2]])
    _check_edge(r[0], oracle.estimate_edge([kfs[1]], [kfs[2]]))
    est.clear()


def test_edge_failure_conventions(est, oracle):
    """score 0 / ok False cases of estimateEdgeDirect (:47-49, :93-95, :118) and transformation_estimator.cpp:53-55"""
    f, t, _ = S.make_pair(500, seed=30)
    few = {k: (v[:6] if isinstance(v, np.ndarray) else v) for k, v in f.items()}             # < 7 keypoints
    other_frame = dict(t, sensor_frame=5)
    other_type = dict(t, feature_type=4)
    surf = (dict(f, feature_type=5), dict(t, feature_type=5))                                  # non-binary type
    no_depth = dict(t, valid=np.zeros_like(t["valid"]))
    for cf, ct in (([few], [t]), ([f], [other_frame]), ([f], [other_type]), ([surf[0]], [surf[1]]),
                   ([f], [no_depth]), ([], [t]), ([f], [])):
        r = est.estimateEdgeDirect(cf, ct)
        o = oracle.estimate_edge(cf, ct)
        _check_edge(r, o)
        assert not r["ok"] and r["consensus"] == 0 and np.array_equal(r["T"].reshape(4, 4), np.eye(4))


def test_edge_multi_camera_best_pair(est, oracle):
    """rig keyframes: several FeatureData per node; only same-frame pairs are matched and the pair with
    the strictly greatest number of ratio survivors wins (first wins ties)  (:40-49, :81-86)"""
    cams_f, cams_t = [], []
    for cam in range(3):
        f, t, _ = S.make_pair(400 + 50 * cam, seed=40 + cam, rho=0.2 + 0.3 * cam, sensor_frame=cam)
        cams_f.append(f); cams_t.append(t)
    cams_t = cams_t[::-1]                                   # frames appear in a different order on the to-side
    r = est.estimateEdgeDirect(cams_f, cams_t)
    o = oracle.estimate_edge(cams_f, cams_t)
    _check_edge(r, o)
    assert r["ok"]
    # two identical cameras on the same frame: the first must win the tie
    f, t, _ = S.make_pair(300, seed=50)
    r = est.estimateEdgeDirect([f, dict(f)], [t])
    o = oracle.estimate_edge([f, dict(f)], [t])
    _check_edge(r, o)
    assert r["cam_from"] == 0


@pytest.mark.parametrize("params", [dict(ransac_iterations=1000), dict(ransac_iterations=1), dict(ransac_threshold=0.02),
                                    dict(break_percentage=0.2), dict(do_prosac=0), dict(ransac_iterations=300, break_percentage=1.0)])
def test_edge_params(est, oracle, params):
    f, t, _ = S.make_pair(800, seed=60)
    try:
        est.setConfig(**params)
        p = est.get_params()
        r = est.estimateEdgeDirect([f], [t])
        o = oracle.estimate_edge([f], [t], thr=p.ransac_threshold, iterations=p.ransac_iterations,
                                 bp=p.break_percentage, do_prosac=bool(p.do_prosac))
        _check_edge(r, o, str(params))
    finally:
        est.setConfig(ransac_threshold=0.1, break_percentage=0.6, ransac_iterations=100, do_prosac=1)


@pytest.mark.parametrize("n,seed,kw", [(500, 70, {}), (1000, 71, dict(tie_stress=True)), (300, 72, dict(rho=0.9)),
                                       (1500, 73, dict(gross_outlier_frac=0.4))])
def test_edge_cross_check(est, oracle, n, seed, kw):
    """opt-in mutual filter (uz_params.cross_check): column minima tracked by the match kernel (tests/test_shapes.py compares the
    fused form with the reversed matching)"""
    f, t, _ = S.make_pair(n, seed=seed, **kw)
    try:
        est.setConfig(cross_check=1)
        est.set_debug(True)
        r = est.estimateEdgeDirect([f], [t])
        o = oracle.estimate_edge([f], [t], cross_check=True)
        _check_edge(r, o, f"cross n={n}")
        m, mask = est.debug_pair(0, r["n_matches"])
        assert np.array_equal(m, o["matches"])
        if o["ok"]:
            assert np.array_equal(mask, o["inlier_mask"])
        plain = oracle.estimate_edge([f], [t])
        assert o["n_ratio_matches"] <= plain["n_ratio_matches"]
        # rigs + store path with the filter on
        cams_f, cams_t = [], []
        for cam in range(2):
            a, b, _ = S.make_pair(300 + 100 * cam, seed=seed + 5 + cam, sensor_frame=cam)
            cams_f.append(a); cams_t.append(b)
        est.clear()
        h = est.add_keyframes([cams_f, cams_t])
        rr = est.estimateEdges([h[0]], [h[1]])[0]
        _check_edge(rr, oracle.estimate_edge(cams_f, cams_t, cross_check=True), "cross rig")
    finally:
        est.set_debug(False)
        est.setConfig(cross_check=0)
        est.clear()


def test_batch_is_order_independent_and_deterministic(est):
    kfs, pairs, _ = S.make_map(30, n_features=500, cluster=6, pool=500, n_shared=300, k_candidates=4,
                               cross_cluster=1, seed=9)
    est.clear()
    h = est.add_keyframes(kfs)
    a = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    perm = np.random.default_rng(0).permutation(len(pairs))
    b = est.estimateEdges(h[pairs[perm, 0]], h[pairs[perm, 1]])
    assert a[perm].tobytes() == b.tobytes()
    c = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    assert a.tobytes() == c.tobytes()
    # host path == store path
    d = est.estimateEdgesHost([([kfs[i]], [kfs[j]]) for i, j in pairs[:10]])
    assert d.tobytes() == a[:10].tobytes()
    est.clear()


def test_full_size_properties(est, oracle):
    """BASELINE config C3 shape (1 query vs many candidates, N=1000): too slow to oracle-check every pair on
    CPU, so check size-independent properties on all pairs and the oracle on a sample."""
    kfs, pairs, poses = S.make_map(200, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=20,
                                   cross_cluster=4, seed=1)
    est.clear()
    h = est.add_keyframes(kfs)
    res = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    same = (pairs[:, 0] // 25) == (pairs[:, 1] // 25)
    assert (res["ok"] == 1).all()
    assert (res["consensus"] <= res["n_matches"]).all() and (res["n_matches"] <= res["n_ratio_matches"]).all()
    assert (res["consensus"][same] >= 50).all()            # true loop closures are found
    assert (res["consensus"][~same] < 20).all()            # false candidates get no support
    for i in np.flatnonzero(same)[:50]:
        T = res[i]["T"].reshape(4, 4)
        Tgt = S.gt_transform(poses, pairs[i, 0], pairs[i, 1])
        assert np.linalg.norm(T[:3, 3] - Tgt[:3, 3]) < 0.05 and S.rot_angle(T[:3, :3], Tgt[:3, :3]) < 0.02
        R = T[:3, :3]
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(R) - 1) < 1e-5
    for i in np.random.default_rng(0).choice(len(pairs), 40, replace=False):
        o = oracle.estimate_edge([kfs[pairs[i, 0]]], [kfs[pairs[i, 1]]])
        _check_edge(res[i], o, f"pair {i}")
    est.clear()


def test_estimate_svd_fuzz_with_awkward_inputs(est, oracle):
    """seeded fuzz of estimateSVD on raw point sets: random sizes / thresholds / iteration counts, huge and tiny
    coordinates, exact duplicates, points at the threshold, and non-finite coordinates (a NaN or an infinity is outside for
    every hypothesis in the reference's double evaluation; a sample that touches one yields a non-finite transform that
    simply never wins) - inlier sets, consensus and the winner must equal the oracle's, finite transforms bit for bit"""
    rng = np.random.default_rng(2024)
    for it in range(40):
        M = int(rng.choice([3, 4, 5, 17, 63, 64, 65, 130, 400, 1100]))
        iters = int(rng.choice([1, 7, 50, 100, 129, 300]))
        thr = float(rng.choice([1e-3, 0.05, 0.1, 0.3, 5.0]))
        bp = float(rng.choice([0.3, 0.6, 1.0]))
        prosac = bool(rng.integers(0, 2))
        scale = float(rng.choice([1.0, 1.0, 1e-3, 1e3]))
        P, Q, _ = _point_problem(rng, M, float(rng.uniform(0.2, 0.95)), noise=0.01 * scale)
        P *= scale; Q *= scale
        kind = it % 5
        if kind == 1 and M > 8:                                   # duplicates of one correspondence
            P[M // 2:M // 2 + 4] = P[0]; Q[M // 2:M // 2 + 4] = Q[0]
        if kind == 2 and M > 8:                                   # non-finite coordinates
            P[3, 1] = np.nan; Q[5, 2] = np.inf; P[M - 1] = -np.inf
        if kind == 3:                                             # far from the origin: the float32 pre-screen's margin grows
            P += 5e4; Q += 5e4
        g = est.estimateSVD(P, Q, thr * scale, iters, bp, prosac)
        o = oracle.estimate_svd(P, Q, thr * scale, iters, bp, prosac)
        tag = (it, M, iters, thr, bp, prosac, scale, kind)
        assert g["consensus"] == o["consensus"] and np.array_equal(g["mask"], o["mask"]), tag
        assert g["best_iteration"] == o["best_iteration"] and g["iterations_run"] == o["iterations_run"], tag
        assert np.array_equal(g["T"], o["T"], equal_nan=True), tag
        assert (np.isnan(g["mse"]) and np.isnan(o["mse"])) or g["mse"] == o["mse"], tag


def test_edge_fuzz_with_awkward_keyframes(est, oracle):
    """whole path on keyframes a front-end can actually produce: few or no depth-valid keypoints, NaN / inf positions behind
    a set valid flag, identical descriptors everywhere (every distance 0: the ratio test drops every match), one camera far
    larger than the other - records must equal the oracle's"""
    rng = np.random.default_rng(77)
    cases = []
    for it in range(24):
        nf, nt = int(rng.choice([7, 8, 40, 300, 1000])), int(rng.choice([7, 9, 64, 333, 900]))
        f, t, _ = S.make_pair(nf, nt, seed=4000 + it, invalid_frac=float(rng.choice([0.0, 0.15, 0.9, 1.0])))
        kind = it % 6
        if kind == 1:
            f["pos"][::3] = np.nan                                    # NaN behind valid flags
        if kind == 2:
            t["pos"][1::4, 0] = np.inf
        if kind == 3:
            f["desc"][:] = 7; t["desc"][:] = 7                        # all distances 0
        if kind == 4:
            t["desc"][:] = t["desc"][0]                               # every query row identical
        if kind == 5:
            f["valid"][:] = 1; t["valid"][:] = 1; f["pos"][:] = f["pos"][0]      # all from-points identical
        cases.append(([f], [t]))
    got = est.estimateEdgesHost(cases)
    for i, (r, (cf, ct)) in enumerate(zip(got, cases)):
        o = oracle.estimate_edge(cf, ct)
        tag = (i, len(cf[0]["desc"]), len(ct[0]["desc"]))
        assert bool(r["ok"]) == o["ok"] and r["n_ratio_matches"] == o["n_ratio_matches"] and r["n_matches"] == o["n_matches"], tag
        assert r["consensus"] == o["consensus"] and r["best_iteration"] == o["best_iteration"], tag
        assert r["iterations_run"] == o["iterations_run"], tag
        assert np.array_equal(r["T"].reshape(4, 4), o["T"], equal_nan=True), tag
        assert (np.isnan(r["mse"]) and np.isnan(o["mse"])) or r["mse"] == o["mse"], tag
