"""8f-1 candidate generation: the oracle's restatement of LshSetRecognizer / PlaceRecognizer
(place_recognition/src/lsh_set_recognizer.cpp, place_recognizer.cpp) checked on hand-made cases (CPU), and the device
recogniser against it (GPU): raw bucket votes, ranked/filtered candidate pairs, incremental batches, add/search/remove."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

SEC = 1_000_000_000


def _cam(desc):
    n = len(desc)
    return dict(desc=np.ascontiguousarray(desc, np.uint8), pos=np.zeros((n, 3)), valid=np.ones(n, np.uint8),
                feature_type=2, sensor_frame=0)


def _brute_votes(q, places, filtered_q, filtered_p):
    """sum over tables of equal 32-bit words, with the popcount>12 filter where the reference applies it"""
    qw = q.view("<u4").reshape(len(q), 8)
    out = []
    for p, fp in zip(places, filtered_p):
        pw = p.view("<u4").reshape(len(p), 8)
        v = 0
        for k in range(8):
            a, b = qw[:, k], pw[:, k]
            if filtered_q:
                a = a[np.array([bin(int(x)).count("1") > 12 for x in a], bool)] if len(a) else a
            if fp:
                b = b[np.array([bin(int(x)).count("1") > 12 for x in b], bool)] if len(b) else b
            if len(a) and len(b):
                ua, ca = np.unique(a, return_counts=True)
                ub, cb = np.unique(b, return_counts=True)
                common, ia, ib = np.intersect1d(ua, ub, return_indices=True)
                v += int((ca[ia] * cb[ib]).sum())
        out.append(v)
    return np.array(out, np.int32)


def test_oracle_votes_are_equal_word_counts(oracle):
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (400, 32), dtype=np.uint8)
    places = []
    for i in range(6):
        d = base.copy()
        flip = rng.random(d.shape) < 0.02 * (i + 1)
        d[flip] ^= rng.integers(1, 256, d.shape, dtype=np.uint8)[flip]
        d[:5, :4] = 0                                     # low-popcount keys: dropped by matchAndAdd, kept by add
        places.append(d)
    P = oracle.Places(T=0.0, k=100)
    for i, d in enumerate(places[:3]):
        P.search_and_add(i, i * 10 * SEC, [_cam(d)])      # filtered inserts
    for i, d in enumerate(places[3:], 3):
        P.add(i, i * 10 * SEC, [_cam(d)])                 # unfiltered inserts
    q = places[0]
    got = P.votes(_cam(q), 6, filtered=False)
    assert np.array_equal(got, _brute_votes(q, places, False, [True] * 3 + [False] * 3))
    got = P.votes(_cam(q), 6, filtered=True)
    assert np.array_equal(got, _brute_votes(q, places, True, [True] * 3 + [False] * 3))


def test_oracle_filters(oracle):
    """ranking by votes, T threshold, |dt| > 5 s, first k, checked_, small keyframes are matched but not inserted"""
    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, (300, 32), dtype=np.uint8)

    def variant(frac):
        d = base.copy()
        rows = rng.random(len(d)) < frac
        d[rows] = rng.integers(0, 256, (int(rows.sum()), 32), dtype=np.uint8)
        return d
    P = oracle.Places(T=1.0, k=2)
    assert len(P.search_and_add(10, 0, [_cam(variant(0.0))])) == 0
    assert len(P.search_and_add(11, 1 * SEC, [_cam(variant(0.1))])) == 0            # 1 s apart: dropped
    r = P.search_and_add(12, 20 * SEC, [_cam(variant(0.2))])
    assert r.tolist() == [[10, 12], [11, 12]]                                       # more shared rows first
    r = P.search_and_add(13, 40 * SEC, [_cam(variant(0.5))])
    assert len(r) == 2 and set(r[:, 1]) == {13}                                     # k = 2
    assert len(P.search_and_add(13, 41 * SEC, [_cam(variant(0.5))])) == 0           # existing place
    small = _cam(base[:100])
    r = P.search_and_add(14, 60 * SEC, [small])                                     # <= 150 rows: match only
    assert len(r) == 2
    r2 = P.search(15, 80 * SEC, [_cam(variant(0.0))])
    assert 14 not in r2[:, 0]                                                       # ... and was never inserted
    again = P.search(15, 80 * SEC, [_cam(variant(0.0))])
    assert len(again) == 0                                                          # checked_ remembers the pairs
    P.remove(10)
    r3 = P.search(16, 100 * SEC, [_cam(variant(0.0))])
    assert 10 not in r3[:, 0]


def _map(n_kf, seed, n_features=400):
    kfs, pairs, _ = S.make_map(n_kf, n_features=n_features, cluster=6, pool=n_features, n_shared=int(0.6 * n_features),
                               k_candidates=4, cross_cluster=1, seed=seed)
    rng = np.random.default_rng(seed)
    order = rng.permutation(n_kf)                        # neighbours must not be adjacent in time
    stamps = (np.arange(n_kf) * 3 * SEC).astype(np.int64)
    return [kfs[i] for i in order], stamps


@pytest.mark.gpu
def test_gpu_votes_match_oracle(est, oracle):
    kfs, stamps = _map(30, 5)
    est.clear()
    h = est.add_keyframes(kfs)
    est.setPlaceConfig(T=1.0, k_nearest_neighbors=5)
    P = oracle.Places(T=1.0, k=5)
    est.searchAndAddPlaces(h[:20], stamps[:20])
    est.addPlaces(h[20:], stamps[20:])
    for i in range(20):
        P.search_and_add(int(h[i]), int(stamps[i]), [kfs[i]])
    for i in range(20, 30):
        P.add(int(h[i]), int(stamps[i]), [kfs[i]])
    assert est.place_count() == 30
    for i in (0, 7, 29):
        for filt in (False, True):
            assert np.array_equal(est.place_votes(h[i], 0, filt), P.votes(kfs[i], 30, filt)), (i, filt)
    est.clear()


@pytest.mark.gpu
@pytest.mark.parametrize("T,k,batch", [(1.0, 20, 60), (2.0, 5, 7), (0.0, 3, 1), (4.0, 20, 60)])
def test_gpu_search_and_add_equals_sequential_oracle(est, oracle, T, k, batch):
    """a batch call == the reference's keyframe-by-keyframe searchAndAddPlace, for any batch split"""
    kfs, stamps = _map(60, 11)
    kfs[7] = {key: (v[:120] if isinstance(v, np.ndarray) else v) for key, v in kfs[7].items()}   # <= 150 rows: match only
    est.clear()
    h = est.add_keyframes(kfs)
    est.setPlaceConfig(T=T, k_nearest_neighbors=k)
    P = oracle.Places(T=T, k=k)
    want = [P.search_and_add(int(h[i]), int(stamps[i]), [kfs[i]]) for i in range(60)]
    want = np.concatenate([w for w in want if len(w)]) if any(len(w) for w in want) else np.zeros((0, 2), np.int64)
    got = [est.searchAndAddPlaces(h[i:i + batch], stamps[i:i + batch]) for i in range(0, 60, batch)]
    got = np.concatenate(got)
    assert len(want) > 0 or T >= 4.0
    assert np.array_equal(got.astype(np.int64), want)
    est.clear()


@pytest.mark.gpu
def test_gpu_places_add_search_remove(est, oracle):
    kfs, stamps = _map(40, 21)
    est.clear()
    h = est.add_keyframes(kfs)
    est.setPlaceConfig(T=1.0, k_nearest_neighbors=6)
    P = oracle.Places(T=1.0, k=6)
    est.addPlaces(h[:30], stamps[:30])                                   # resume path: addPlace for every stored node
    for i in range(30):
        P.add(int(h[i]), int(stamps[i]), [kfs[i]])
    got = est.searchPlaces(h[30:], stamps[30:])
    want = np.concatenate([P.search(int(h[i]), int(stamps[i]), [kfs[i]]) for i in range(30, 40)])
    assert len(want) > 0 and np.array_equal(got.astype(np.int64), want)
    assert len(est.searchPlaces(h[30:], stamps[30:])) == 0               # checked_
    victim = int(want[0, 0])
    est.removePlace(victim)
    P.remove(victim)
    got = est.searchAndAddPlaces(h[30:], stamps[30:])
    want = np.concatenate([P.search_and_add(int(h[i]), int(stamps[i]), [kfs[i]]) for i in range(30, 40)] + [np.zeros((0, 2), np.int64)])
    assert np.array_equal(got.astype(np.int64), want)
    assert victim not in got[:, 0]
    with pytest.raises(Exception):
        est.removePlace(victim)
    est.clearPlaces()
    assert est.place_count() == 0
    est.clear()


@pytest.mark.gpu
def test_gpu_places_rig_and_table_growth(est, oracle):
    """two cameras per keyframe (lists are concatenated per sensor), and enough entries to force the slot table and
    the node array to grow and re-link"""
    rng = np.random.default_rng(3)
    kfs, stamps = _map(24, 31, n_features=300)
    rig = []
    for i in range(0, 24, 2):
        rig.append([kfs[i], dict(kfs[i + 1], sensor_frame=1)])
    st = (np.arange(len(rig)) * 7 * SEC).astype(np.int64)
    est.clear()
    h = est.add_keyframes(rig)
    est.setPlaceConfig(T=0.5, k_nearest_neighbors=4)
    P = oracle.Places(T=0.5, k=4)
    want = [P.search_and_add(int(h[i]), int(st[i]), rig[i]) for i in range(len(rig))]
    want = np.concatenate(want + [np.zeros((0, 2), np.int64)])
    got = est.searchAndAddPlaces(h, st)
    assert np.array_equal(got.astype(np.int64), want)
    # growth: 70 x 2000 x 8 = 1.12 M entries > the initial 2^20 nodes and > half the initial 2^20 slots
    big = [dict(desc=rng.integers(0, 256, (2000, 32), dtype=np.uint8), pos=np.zeros((2000, 3)), valid=np.ones(2000, np.uint8),
                feature_type=2, sensor_frame=0) for _ in range(70)]
    for j in range(1, 70):
        big[j]["desc"][:500] = big[0]["desc"][:500]
    est.clear()
    hb = est.add_keyframes(big)
    sb = (np.arange(70) * 10 * SEC).astype(np.int64)
    P = oracle.Places(T=0.5, k=4)
    want = np.concatenate([P.search_and_add(int(hb[i]), int(sb[i]), [big[i]]) for i in range(70)])
    got = np.concatenate([est.searchAndAddPlaces(hb[i:i + 8], sb[i:i + 8]) for i in range(0, 70, 8)])
    assert np.array_equal(got.astype(np.int64), want)
    est.clear()
