"""The device-resident keyframe store: one arena range per keyframe, given back by uz_store_remove (the reference removes and
merges nodes for as long as it runs, graph_slam_node.cpp:641,665-777,1050); pageable and pinned host buffers take different
roads to the device (pinned ring + gather kernel / direct gather) and must land the same bytes."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def test_add_remove_ten_times_the_capacity_keeps_store_bytes_flat(est):
    est.clear()
    rng = np.random.default_rng(3)
    live = []
    peak = 0
    for i in range(40):                                   # resident set of ~40 keyframes
        f, _, _ = S.make_pair(int(rng.integers(300, 1000)), seed=1000 + i)
        live.append(est.add_keyframe([f]))
    base = est.store_bytes()
    for i in range(400):                                  # 10 x the capacity goes through
        est.remove_keyframe(live.pop(int(rng.integers(0, len(live)))))
        f, _, _ = S.make_pair(int(rng.integers(300, 1000)), seed=2000 + i)
        live.append(est.add_keyframe([f]))
        peak = max(peak, est.store_bytes())
    assert est.store_size() == 40
    assert peak < 1.6 * base, (peak, base)                # sizes vary by 3x; a leak would grow by 10x
    for h in live:
        est.remove_keyframe(h)
    assert est.store_bytes() == 0 and est.store_size() == 0
    est.clear()


def test_removed_handle_is_recycled_without_stale_place_or_data(est, oracle):
    """a recycled handle must not inherit the place, the checked_ pairs or the rows of its predecessor"""
    est.clear(); est.clearPlaces()
    kfs, pairs, _ = S.make_map(12, n_features=400, cluster=4, pool=400, n_shared=250, k_candidates=3, cross_cluster=0, seed=5)
    h = est.add_keyframes(kfs)
    stamps = np.arange(len(h), dtype=np.int64) * 10_000_000_000
    est.setPlaceConfig(T=1.0, k_nearest_neighbors=5)
    first = est.searchAndAddPlaces(h, stamps)
    assert len(first) > 0
    victim = int(h[1])
    est.remove_keyframe(victim)
    f2, _, _ = S.make_pair(400, seed=991)                 # an unrelated keyframe takes the recycled handle
    h2 = est.add_keyframe([f2])
    assert h2 == victim
    got = est.searchAndAddPlaces(np.array([h2], np.int32), np.array([10 ** 12], np.int64))
    # the new keyframe IS searched and inserted (not skipped as "existing place"), and nothing votes for the dead rows
    assert est.place_count() == len(h) + 1
    assert not any(int(a) == victim for a, b in got)
    r = est.estimateEdges(np.array([h[0]], np.int32), np.array([h2], np.int32))[0]
    o = oracle.estimate_edge([kfs[0]], [f2])
    assert r["n_matches"] == o["n_matches"] and r["consensus"] == o["consensus"]
    est.clear()


def test_pageable_and_pinned_sources_give_identical_records(est):
    import torch
    kfs, pairs, _ = S.make_map(120, n_features=600, cluster=6, pool=600, n_shared=350, k_candidates=8, cross_cluster=2, seed=9)
    pageable = est.estimateEdgesHost([([kfs[a]], [kfs[b]]) for a, b in pairs])
    pinned_kfs = []
    keep = []
    for k in kfs:
        d = torch.from_numpy(k["desc"].copy()).pin_memory(); p = torch.from_numpy(k["pos"].copy()).pin_memory()
        v = torch.from_numpy(k["valid"].copy()).pin_memory()
        keep += [d, p, v]
        pinned_kfs.append(dict(k, desc=d.numpy(), pos=p.numpy(), valid=v.numpy()))
    pinned = est.estimateEdgesHost([([pinned_kfs[a]], [pinned_kfs[b]]) for a, b in pairs])
    assert pageable.tobytes() == pinned.tobytes()
    # strided descriptor rows (a cv::Mat with padded rows) and a tiny staging ring that wraps many times
    wide = np.zeros((600, 48), np.uint8)
    strided = []
    for k in kfs[:20]:
        w = wide.copy(); w[:, :32] = k["desc"]
        strided.append(dict(k, desc=w[:, :32]))
    a = est.estimateEdgesHost([([strided[i]], [strided[i + 1]]) for i in range(19)])
    b = est.estimateEdgesHost([([kfs[i]], [kfs[i + 1]]) for i in range(19)])
    assert a.tobytes() == b.tobytes()


def test_small_ring_wraps(oracle):
    import os
    from uzliti_slam_b200 import EdgeEstimator
    os.environ["UZ_RING_MB"] = "1"
    try:
        e = EdgeEstimator(0)
    finally:
        os.environ.pop("UZ_RING_MB", None)
    try:
        kfs, pairs, _ = S.make_map(150, n_features=900, cluster=5, pool=900, n_shared=500, k_candidates=6, cross_cluster=1, seed=19)
        h = e.add_keyframes(kfs)                              # ~8 MB of pageable arrays through a 2 x 1 MB ring
        for i in (0, 77, 149):
            back = e.read_keyframe(int(h[i]))
            assert np.array_equal(back["desc"], kfs[i]["desc"]) and np.array_equal(back["pos"], kfs[i]["pos"])
            assert np.array_equal(back["valid"], kfs[i]["valid"])
        r = e.estimateEdges(h[pairs[:30, 0]], h[pairs[:30, 1]])
        for rec, (a, b) in list(zip(r, pairs))[:6]:
            o = oracle.estimate_edge([kfs[a]], [kfs[b]])
            assert rec["consensus"] == o["consensus"] and rec["n_matches"] == o["n_matches"]
    finally:
        e.close()


@pytest.mark.gpu
def test_pipelined_synchronous_call_equals_one_launch_pair(built):
    """uz_estimate_edges cuts a large store-resident batch into chunks that double in size (the host prepares chunk c + 1
    while chunk c runs); UZ_PIPELINE_CALLS=0 keeps one launch pair: byte-identical records, more launches"""
    import os
    from uzliti_slam_b200 import EdgeEstimator
    kfs, pairs, _ = S.make_map(420, n_features=300, cluster=10, pool=300, n_shared=180, k_candidates=20, cross_cluster=4, seed=71)
    assert len(pairs) >= 8 * 5 * 148
    os.environ["UZ_PIPELINE_CALLS"] = "0"
    try:
        one = EdgeEstimator(0)
    finally:
        os.environ.pop("UZ_PIPELINE_CALLS", None)
    piped = EdgeEstimator(0)
    try:
        ho, hp = one.add_keyframes(kfs), piped.add_keyframes(kfs)
        n0, n1 = one.launch_count(), piped.launch_count()
        a = one.estimateEdges(ho[pairs[:, 0]], ho[pairs[:, 1]])
        b = piped.estimateEdges(hp[pairs[:, 0]], hp[pairs[:, 1]])
        assert a.tobytes() == b.tobytes() and (a["ok"] == 1).sum() > 1000
        assert one.launch_count() - n0 == 2 and piped.launch_count() - n1 >= 6
    finally:
        one.close(); piped.close()
