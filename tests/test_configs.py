"""The remaining BASELINE.json configs as parity cases (bench.py measures C4; C1/C3 are covered by test_gpu_parity.py):
C2 sequential visual odometry (consecutive-pair edges along a chain of 1000-feature frames) and C5 (rigs of 4 cameras x
2000 features per keyframe, 1000 RANSAC hypotheses per pair)."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S


def _same(r, o, tag):
    assert bool(r["ok"]) == o["ok"], tag
    for k in ("cam_from", "cam_to", "n_ratio_matches", "n_matches", "consensus", "best_iteration", "iterations_run"):
        assert int(r[k]) == int(o[k]), (tag, k, int(r[k]), int(o[k]))
    assert np.array_equal(r["T"].reshape(4, 4), o["T"]), tag
    assert r["mse"] == o["mse"] or (np.isnan(r["mse"]) and np.isnan(o["mse"])), tag


@pytest.mark.gpu
def test_c2_visual_odometry_chain(est, oracle):
    """frames of one trajectory segment share a landmark pool; edges are estimated between consecutive frames, queued
    the way the reference's odometry thread does (graph_slam_node.cpp:266) but delivered as one batch"""
    kfs, _, poses = S.make_map(60, n_features=1000, cluster=60, pool=1400, n_shared=700, k_candidates=1, cross_cluster=0, seed=12)
    est.clear()
    h = est.add_keyframes(kfs)
    frm, to = np.arange(0, 59), np.arange(1, 60)
    res = est.estimateEdges(h[frm], h[to])
    assert (res["ok"] == 1).all() and (res["consensus"] >= 50).all()
    for i in range(59):
        _same(res[i], oracle.estimate_edge([kfs[i]], [kfs[i + 1]]), f"vo pair {i}")
        T = res[i]["T"].reshape(4, 4)
        Tgt = S.gt_transform(poses, i, i + 1)
        assert np.linalg.norm(T[:3, 3] - Tgt[:3, 3]) < 0.05 and S.rot_angle(T[:3, :3], Tgt[:3, :3]) < 0.02
    # composing the chain's edges reproduces the end-to-end motion (drift bounded by the per-edge error)
    acc = np.eye(4)
    for i in range(59):
        acc = acc @ res[i]["T"].reshape(4, 4)
    Tgt = S.gt_transform(poses, 0, 59)
    assert np.linalg.norm(acc[:3, 3] - Tgt[:3, 3]) < 0.5 and S.rot_angle(acc[:3, :3], Tgt[:3, :3]) < 0.1
    est.clear()


@pytest.mark.gpu
def test_c5_rig_pairs_1000_hypotheses(est, oracle):
    """4 cameras x 2000 features per keyframe: 4 same-frame matchings of 2000 x 2000 per pair, the best camera pair goes
    through RANSAC with 1000 hypotheses (cfg/FeatureLinkEstimation.cfg:11 maximum)"""
    rigs_f, rigs_t = [], []
    for p in range(3):
        cf, ct = [], []
        for cam in range(4):
            f, t, _ = S.make_pair(2000, seed=500 + 10 * p + cam, rho=0.15 + 0.1 * ((cam + p) % 4), sensor_frame=cam)
            cf.append(f); ct.append(t)
        rigs_f.append(cf); rigs_t.append(ct)
    try:
        est.setConfig(ransac_iterations=1000)
        est.clear()
        hf = est.add_keyframes(rigs_f)
        ht = est.add_keyframes(rigs_t)
        res = est.estimateEdges(hf, ht)
        host = est.estimateEdgesHost(list(zip(rigs_f, rigs_t)))
        assert res.tobytes() == host.tobytes()
        for p in range(3):
            o = oracle.estimate_edge(rigs_f[p], rigs_t[p], iterations=1000)
            _same(res[p], o, f"rig pair {p}")
            assert res[p]["ok"] and res[p]["cam_from"] == res[p]["cam_to"] == int(np.argmax([(0.15 + 0.1 * ((c + p) % 4)) for c in range(4)]))
    finally:
        est.setConfig(ransac_iterations=100)
        est.clear()
