"""8f-2, the step after the path: TransformationFilter::calcValidEdges' cluster RANSAC (transformation_filter.cpp:216-285)
as one batched launch, and GraphSlamNode::newEdgeCallback's numeric gate (graph_slam_node.cpp:798-804)."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S
from uzliti_slam_b200.binding import RESULT_DTYPE


def _rot(axis, deg):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    a = np.deg2rad(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


def test_oracle_gate_angles(oracle):
    """the written-out Eigen quaternion/angle-axis path returns the planted angle, on every Shoemake branch"""
    for axis, deg in (((0, 0, 1), 10.0), ((1, 0, 0), 29.999), ((0, 1, 0), 179.0), ((1, 1, 1), 120.0), ((1, 2, 3), 150.0),
                      ((0, 0, 1), 0.0), ((1, 0, 0), 180.0), ((0, 0, 1), 180.0), ((0, 1, 0), 180.0)):
        T = np.eye(4)
        T[:3, :3] = _rot(axis, deg)
        T[:3, 3] = (0.3, -0.4, 1.2)
        acc, tn, rot = oracle.gate_edge(T, True, 50)
        assert abs(rot - deg) < 1e-6 and abs(tn - 1.3) < 1e-12
        assert acc == (deg <= 30.0)
    T = np.eye(4)
    assert oracle.gate_edge(T, True, 19)[0] is False and oracle.gate_edge(T, True, 20)[0] is True
    assert oracle.gate_edge(T, False, 100)[0] is False                   # failed estimate: score 0
    T[:3, 3] = (1.5, 0, 0)
    assert oracle.gate_edge(T, True, 50)[0] is True
    T[:3, 3] = (1.5000001, 0, 0)
    assert oracle.gate_edge(T, True, 50)[0] is False


@pytest.mark.gpu
def test_gpu_gate_matches_oracle(est, oracle):
    rng = np.random.default_rng(5)
    n = 4000
    res = np.zeros(n, RESULT_DTYPE)
    for i in range(n):
        T = np.eye(4)
        deg = [0.0, 180.0, rng.uniform(0, 60), rng.uniform(170, 180), rng.uniform(0, 1e-5)][i % 5]
        T[:3, :3] = _rot(rng.normal(size=3), deg).astype(np.float32)      # the path returns float32-valued rotations
        T[:3, 3] = rng.normal(size=3) * rng.choice([0.1, 0.8, 2.0])
        res["T"][i] = T.reshape(16)
        res["ok"][i] = i % 7 != 0
        res["consensus"][i] = rng.integers(0, 60)
    acc, tn, rot = est.gateEdges(res)
    for i in range(n):
        a, t, r = oracle.gate_edge(res["T"][i], res["ok"][i], res["consensus"][i])
        assert t == tn[i]
        assert abs(r - rot[i]) <= 1e-12 * max(1.0, r)                     # acos: 1 ulp library difference at most
        if abs(r - 30.0) > 1e-9:
            assert a == acc[i], i
    assert 0 < acc.sum() < n
    # straight behind the path: records of a real batch
    f, t, _ = S.make_pair(600, seed=3)
    r = est.estimateEdgesHost([([f], [t])] * 3)
    acc, tn, rot = est.gateEdges(r, min_matching_score=20.0, max_edge_distance_T=10.0, max_edge_distance_R=180.0)
    assert acc.all()
    a0, t0, r0 = oracle.gate_edge(r["T"][0], r["ok"][0], r["consensus"][0], 20.0, 10.0, 180.0)
    assert a0 and t0 == tn[0] and abs(r0 - rot[0]) < 1e-12


@pytest.mark.gpu
def test_gpu_cluster_ransac_batch_matches_oracle(est, oracle):
    """calcValidEdges: estimateSVD(P, Q, T, c, mse, 0.3, 200, 1.0, false) + consensus3D per cluster, all clusters at once"""
    rng = np.random.default_rng(9)
    Ps, Qs = [], []
    for c in range(37):
        M = int(rng.integers(3, 100)) if c % 9 else [0, 1, 2, 3, 100][c // 9]
        R = _rot(rng.normal(size=3), rng.uniform(0, 40))
        t = rng.normal(size=3)
        P = rng.normal(size=(M, 3)) * 3
        Q = P @ R.T + t + rng.normal(size=(M, 3)) * 0.05
        bad = rng.random(M) < 0.3
        Q[bad] += rng.normal(size=(int(bad.sum()), 3)) * 2
        Ps.append(P); Qs.append(Q)
    got = est.estimateSVDBatch(Ps, Qs, 0.3, 200, 1.0, False)
    for c, (P, Q) in enumerate(zip(Ps, Qs)):
        o = oracle.estimate_svd(P, Q, 0.3, 200, 1.0, do_prosac=False)
        g = got[c]
        assert g["consensus"] == o["consensus"], c
        assert np.array_equal(g["T"], o["T"]), c
        assert np.array_equal(g["mask"], o["mask"]), c
        assert (g["mse"] == o["mse"]) or (np.isnan(g["mse"]) and np.isnan(o["mse"]))
        if len(P) >= 3:
            cnt, mask = oracle.consensus3d(P, Q, o["T"], 0.3)             # the second call of calcValidEdges
            assert cnt == g["consensus"] and np.array_equal(mask, g["mask"])
    # and it equals the one-problem entry point
    one = est.estimateSVD(Ps[5], Qs[5], 0.3, 200, 1.0, do_prosac=False)
    assert np.array_equal(one["T"], got[5]["T"]) and one["consensus"] == got[5]["consensus"]
    assert est.estimateSVDBatch([], []) == []
