"""512-bit rows: the reference matches BRISK and FREAK descriptors (64 bytes) through the same BFMatcher call as ORB
(feature_transformation_estimator.cpp:54-58; cv::BRISK is even FeatureExtractionCore's default,
feature_extraction_core.cpp:46-49).  knn2_wide_kernel, the 513-bin counting sort of the solve kernel, the two-launch
batch (256-bit tiles + 512-bit tiles), the store / ingestion formats and the place recogniser's row stride are checked
here against the oracle (pinned to cv2 for 64-byte rows in tests/test_oracle_matching.py and tests/golden/edge_wide*.npz).
"""
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S
from test_gpu_parity import _check_edge
from test_ingest import encode_features, _depth_image

pytestmark = pytest.mark.gpu


def _estimator(**env):
    from uzliti_slam_b200 import EdgeEstimator
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return EdgeEstimator(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ---- K1 on 64-byte rows ---------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt", [(500, 500), (1000, 1000), (2000, 1500), (1, 1), (1, 2), (7, 3), (33, 1), (257, 511),
                                   (1000, 513), (513, 63), (64, 65), (129, 64), (3, 1000), (600, 257), (4096, 4096)])
def test_knn2_wide_random(est, oracle, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt + 1)
    q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 64), dtype=np.uint8)
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)


@pytest.mark.parametrize("other", [0, 2])
def test_knn2_wide_integer_pipe_kernel_equals_tensor_core_kernel(est, oracle, other):
    """the default for 64-byte rows is knn2_mmaf_kernel<true> (tensor cores, 4-bit operands, K = 512); UZ_MATCH_MMA_WIDE=0 keeps
    knn2_wide_kernel on the integer pipes, UZ_MATCH_MMA_WIDE=2 knn2_mmaw_kernel on two int8 planes: same neighbours on random,
    tie-heavy and extreme rows, on the 240-row tile boundaries of the 4-bit kernel, and byte-identical edge records"""
    e = _estimator(UZ_MATCH_MMA_WIDE=other)
    try:
        for nq, nt in [(500, 500), (1000, 1000), (257, 511), (129, 64), (3, 1000), (1, 1), (4096, 4096), (1000, 127), (130, 129),
                       (256, 240), (300, 241), (300, 239), (128, 480), (200, 368), (260, 369), (31, 304), (40, 113), (7, 2), (65, 193)]:
            for mode in range(3):
                rng = np.random.default_rng(nq * 31 + nt + mode)
                q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
                t = rng.integers(0, 256, (nt, 64), dtype=np.uint8)
                if mode == 1:
                    q[:, 1:] = 0; t[:, 1:] = 0                   # ties everywhere
                if mode == 2:
                    q[0] = 0; t[0] = 255; t[nt - 1] = 0; q[nq - 1] = 255      # distances 0 and 512
                a = est.knnMatch(q, t)
                b = e.knnMatch(q, t)
                o = oracle.knn2(q, t)
                assert np.array_equal(a[0], o[0]) and np.array_equal(a[1], o[1]), (nq, nt, mode)
                assert np.array_equal(b[0], o[0]) and np.array_equal(b[1], o[1]), (nq, nt, mode)
        kfs, pairs, _ = S.make_map(60, n_features=700, cluster=6, pool=700, n_shared=400, k_candidates=6, cross_cluster=2, seed=8, desc_bytes=64)
        for cross in (0, 1):
            recs = []
            for x in (est, e):
                x.setConfig(cross_check=cross)
                x.clear()
                h = x.add_keyframes(kfs)
                recs.append(x.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]]))
            assert recs[0].tobytes() == recs[1].tobytes()
            assert (recs[0]["ok"] == 1).any()
        est.setConfig(cross_check=0)
        est.clear()
    finally:
        e.close()


@pytest.mark.parametrize("cfg", [0, 1])
def test_knn2_wide_both_shapes(oracle, cfg):
    """256 x 2 and 64 x 2 query CTAs (UZ_KNN_WIDE_CFG forces one)"""
    e = _estimator(UZ_KNN_WIDE_CFG=cfg, UZ_MATCH_MMA_WIDE=0)
    try:
        for nq, nt in [(700, 900), (130, 70), (1000, 1000)]:
            rng = np.random.default_rng(nq + nt + cfg)
            q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, 64), dtype=np.uint8)
            idx, dist = e.knnMatch(q, t)
            oi, od = oracle.knn2(q, t)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    finally:
        e.close()


@pytest.mark.parametrize("keep_bytes", [1, 2, 33, 40])
def test_knn2_wide_tie_stress(est, oracle, keep_bytes):
    rng = np.random.default_rng(keep_bytes)
    q = rng.integers(0, 256, (1000, 64), dtype=np.uint8)
    t = rng.integers(0, 256, (1500, 64), dtype=np.uint8)
    q[:, keep_bytes:] = 0
    t[:, keep_bytes:] = 0
    if keep_bytes > 2:                                  # low entropy in the SECOND half only
        q[:, :32] = 0
        t[:, :32] = 0
        q[:, 34:] = 0
        t[:, 34:] = 0
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(q, t)
    assert (od[:, 0] == od[:, 1]).sum() > 100
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_knn2_wide_extremes_and_strides(est, oracle):
    """distances up to 512 (complement rows), duplicate rows, padded cv::Mat rows (stride 80 / 72 bytes)"""
    rng = np.random.default_rng(5)
    tb = rng.integers(0, 256, (600, 80), dtype=np.uint8)
    t = tb[:, :64]
    t[100] = t[3]; t[599] = t[3]; t[0] = 0; t[1] = 255
    qb = np.zeros((94, 72), np.uint8)
    q = qb[:, :64]
    q[:] = np.concatenate([t[[3, 100, 0, 1]], 255 - t[:50], rng.integers(0, 256, (40, 64), dtype=np.uint8)])
    idx, dist = est.knnMatch(q, t)
    oi, od = oracle.knn2(np.ascontiguousarray(q), np.ascontiguousarray(t))
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    assert idx[0, 0] == 3 and idx[0, 1] == 100 and dist[0, 0] == 0 and dist[0, 1] == 0
    only = est.knnMatch(255 - t[:1], t[:1])             # one train row: the complement, second neighbour missing
    assert only[0].tolist() == [[0, -1]] and only[1].tolist() == [[512, -1]]


def test_knn2_wide_matches_cv2(est):
    cv2 = pytest.importorskip("cv2")
    f, t, _ = S.make_pair(1000, seed=12, desc_bytes=64)
    idx, dist = est.knnMatch(t["desc"], f["desc"])
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(t["desc"], f["desc"], k=2)
    ci = np.array([[a.trainIdx, b.trainIdx] for a, b in m])
    cd = np.array([[a.distance, b.distance] for a, b in m]).astype(np.int32)
    assert np.array_equal(idx, ci) and np.array_equal(dist, cd)


def test_unsupported_widths_are_refused(est):
    from uzliti_slam_b200 import UzError
    for nb in (16, 48, 128):
        with pytest.raises(UzError):
            est.knnMatch(np.zeros((10, nb), np.uint8), np.zeros((10, nb), np.uint8))
        with pytest.raises(UzError):
            est.add_keyframe([dict(desc=np.zeros((10, nb), np.uint8), pos=np.zeros((10, 3)), valid=np.ones(10, np.uint8))])


# ---- the whole path ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(n_from=500, seed=31), dict(n_from=1000, seed=32), dict(n_from=300, n_to=777, seed=33),
                                dict(n_from=256, seed=34, tie_stress=True), dict(n_from=150, seed=35, invalid_frac=0.95),
                                dict(n_from=2000, seed=36, rho=0.8)])
def test_wide_edge_direct_with_taps(est, oracle, kw):
    f, t, _ = S.make_pair(desc_bytes=64, **kw)
    o = oracle.estimate_edge([f], [t])
    est.set_debug(True)
    try:
        r = est.estimateEdgeDirect([f], [t])
        m, mask = est.debug_pair(0, r["n_matches"])
        counts = est.debug_counts(0)
    finally:
        est.set_debug(False)
    _check_edge(r, o, str(kw))
    assert np.array_equal(m, o["matches"])
    if o["ok"]:
        assert np.array_equal(mask, o["inlier_mask"])
        assert np.array_equal(counts, o["counts"])
        assert np.array_equal(r["T"].reshape(4, 4), o["T"])


def test_wide_cross_check(est, oracle):
    f, t, _ = S.make_pair(600, seed=41, desc_bytes=64, tie_stress=True)
    est.setConfig(cross_check=1)
    try:
        r = est.estimateEdgeDirect([f], [t])
    finally:
        est.setConfig(cross_check=0)
    o = oracle.estimate_edge([f], [t], cross_check=True)
    plain = oracle.estimate_edge([f], [t])
    assert o["n_ratio_matches"] < plain["n_ratio_matches"]
    _check_edge(r, o)


def test_mixed_widths_in_one_batch(oracle):
    """ORB (32 B) and BRISK (64 B) keyframes in one store and one batch: two match launches, one solve; cameras of
    different width are never compared (cv::BFMatcher would throw), so such a pair reports ok = 0 like any other
    incomparable pair."""
    e = _estimator(UZ_STREAM_SOLVE=0)
    try:
        kn, pn, _ = S.make_map(12, n_features=400, cluster=6, pool=400, n_shared=250, k_candidates=3, cross_cluster=1, seed=3)
        kw, pw, _ = S.make_map(12, n_features=500, cluster=6, pool=500, n_shared=300, k_candidates=3, cross_cluster=1, seed=4,
                               desc_bytes=64)
        for k in kw:
            k["feature_type"] = 2                        # same feature_type_ on purpose: the WIDTH must keep them apart
        kfs = kn + kw
        h = e.add_keyframes(kfs)
        pairs = np.concatenate([pn, pw + len(kn), [[0, len(kn)], [len(kn) + 1, 1]]])
        rng = np.random.default_rng(0)
        pairs = pairs[rng.permutation(len(pairs))]
        n0 = e.launch_count()
        res = e.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        assert e.launch_count() - n0 == 3                # narrow match + wide match (both on the tensor cores) + solve
        for r, (a, b) in zip(res, pairs):
            if kfs[a]["desc"].shape[1] != kfs[b]["desc"].shape[1]:
                assert r["ok"] == 0 and r["consensus"] == 0 and r["cam_from"] == -1
            else:
                _check_edge(r, oracle.estimate_edge([kfs[a]], [kfs[b]]), f"{a}->{b}")
        assert (res["consensus"] > 50).sum() >= 20
        # the same batch from host buffers
        host = e.estimateEdgesHost([([kfs[a]], [kfs[b]]) for a, b in pairs])
        assert host.tobytes() == res.tobytes()
    finally:
        e.close()


def test_wide_rig_picks_the_best_camera_pair(est, oracle):
    """two cameras per keyframe (different sensor frames), 64-byte rows: :74-86 on wide tasks"""
    fa, ta, _ = S.make_pair(400, seed=51, desc_bytes=64, sensor_frame=0, rho=0.3)
    fb, tb, _ = S.make_pair(450, seed=52, desc_bytes=64, sensor_frame=1, rho=0.7)
    r = est.estimateEdgeDirect([fa, fb], [ta, tb])
    o = oracle.estimate_edge([fa, fb], [ta, tb])
    _check_edge(r, o)
    assert r["cam_from"] == 1 and r["cam_to"] == 1


def test_wide_streaming_equals_one_cta_per_pair():
    kfs, pairs, _ = S.make_map(80, n_features=600, cluster=10, pool=600, n_shared=350, k_candidates=8, cross_cluster=2, seed=6,
                               desc_bytes=64)
    kfs = [{k: (v[:[600, 520, 300, 64, 5][i % 5]] if isinstance(v, np.ndarray) else v) for k, v in kf.items()}
           for i, kf in enumerate(kfs)]
    plain = _estimator(UZ_STREAM_SOLVE=0, UZ_MATCH_MMA_WIDE=0)      # the streaming solve runs beside the integer-pipe kernels only
    stream = _estimator(UZ_STREAM_SOLVE=1, UZ_STREAM_SOLVE_MIN_PAIRS=1, UZ_MATCH_MMA_WIDE=0)
    try:
        hp = plain.add_keyframes(kfs)
        hs = stream.add_keyframes(kfs)
        a = plain.estimateEdges(hp[pairs[:, 0]], hp[pairs[:, 1]])
        b = stream.estimateEdges(hs[pairs[:, 0]], hs[pairs[:, 1]])
        assert a.tobytes() == b.tobytes()
        assert (a["ok"] == 0).any() and (a["consensus"] > 50).any()
    finally:
        plain.close()
        stream.close()


# ---- store, ingestion formats, place recogniser ---------------------------------------------------------
def test_wide_store_roundtrip_and_wire(est, oracle):
    f, t, _ = S.make_pair(300, seed=61, desc_bytes=64)
    h = est.add_keyframe([f])
    back = est.read_keyframe(h)
    assert back["desc"].shape == (300, 64) and np.array_equal(back["desc"], f["desc"])
    assert np.array_equal(back["pos"], f["pos"]) and np.array_equal(back["valid"], f["valid"])
    rng = np.random.default_rng(2)
    u = rng.integers(0, 640, 300); v = rng.integers(0, 480, 300)
    weird = f["desc"].astype(np.float32)
    weird[0, :6] = [255.9, 256.0, -1.0, 1e10, np.nan, 3.7]
    blob = encode_features(weird, u, v, f["valid"], np.full(300, -1.0), f["pos"])
    assert len(blob) == 4 + 300 * 297
    gd, gp, gv, guv = est.wire_decode(blob)
    od, op, ov, ouv = oracle.wire_decode(blob, cols=64)
    assert gd.shape == (300, 64) and np.array_equal(gd, od) and np.array_equal(gp, op) and np.array_equal(gv, ov)
    assert np.array_equal(guv, ouv)
    blob = encode_features(f["desc"].astype(np.float32), u, v, f["valid"], np.full(300, -1.0), f["pos"])
    hw = est.add_keyframe_wire(blob, feature_type=3)
    ht = est.add_keyframe([t])
    a = est.estimateEdges([h, hw], [ht, ht])
    assert a[0].tobytes() == a[1].tobytes()
    _check_edge(a[0], oracle.estimate_edge([f], [t]))
    for x in (h, hw, ht):
        est.remove_keyframe(x)


@pytest.mark.parametrize("reverse", [False, True])
def test_wide_rgbd_ingest(est, oracle, reverse):
    rng = np.random.default_rng(7)
    d = _depth_image(rng)
    n = 500
    u = rng.integers(0, 640, n).astype(np.int32)
    v = rng.integers(0, 480, n).astype(np.int32)
    desc = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    h = est.add_keyframe_rgbd(desc, u, v, d, reverse=reverse, feature_type=3)
    back = est.read_keyframe(h)
    op, ov = oracle.backproject(u, v, d, reverse=reverse)
    assert np.array_equal(back["desc"], desc[::-1] if reverse else desc)
    assert back["pos"].tobytes() == op.tobytes() and np.array_equal(back["valid"], ov)
    q = rng.integers(0, 256, (100, 64), dtype=np.uint8)
    hq = est.add_keyframe([dict(desc=q, pos=np.zeros((100, 3)), valid=np.ones(100, np.uint8), feature_type=3)])
    r = est.estimateEdges([h], [hq])[0]                 # matching runs on the CSA copy the ingest built
    o = oracle.estimate_edge([dict(desc=back["desc"], pos=back["pos"], valid=back["valid"], feature_type=3)],
                             [dict(desc=q, pos=np.zeros((100, 3)), valid=np.ones(100, np.uint8), feature_type=3)])
    assert r["n_ratio_matches"] == o["n_ratio_matches"] and r["n_matches"] == o["n_matches"]
    est.remove_keyframe(h)
    est.remove_keyframe(hq)


def test_wide_places(oracle):
    """LshSetRecognizer keys on bytes [0, 32) of a row whatever its width (lsh_set_recognizer.cpp:243: 32 - key_width + 1)"""
    kfs, _, _ = S.make_map(60, n_features=300, cluster=6, pool=300, n_shared=200, k_candidates=4, cross_cluster=1, seed=9,
                           desc_bytes=64)
    e = _estimator()
    try:
        h = e.add_keyframes(kfs)
        stamps = (np.arange(len(kfs), dtype=np.int64) * 7 + 1) * 1_000_000_000
        e.setPlaceConfig(T=2.0, k_nearest_neighbors=5)
        got = e.searchAndAddPlaces(h, stamps)
        ref = oracle.Places(T=2.0, k=5)
        want = [ref.search_and_add(int(h[i]), int(stamps[i]), [kfs[i]]) for i in range(len(kfs))]
        want = np.concatenate(want + [np.zeros((0, 2), np.int64)])
        assert len(want) > 30 and np.array_equal(got.astype(np.int64), want)
        for filt in (False, True):
            assert np.array_equal(e.place_votes(int(h[-1]), 0, filt), ref.votes(kfs[-1], len(kfs), filt))
    finally:
        e.close()
