"""CPU: the oracle's RANSAC stage.  The reference has no tests or golden vectors and PCL/Eigen are not
installable here ("parity unpinned"), so the float32 solve is anchored on (a) an independent
float64 Kabsch/Umeyama solve, (b) planted ground truth, (c) the sequential semantics of prosac()."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S


def _kabsch(P, Q):
    mp, mq = P.mean(0), Q.mean(0)
    H = (Q - mq).T @ (P - mp)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    R = U @ D @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mq - R @ mp
    return T


@pytest.mark.parametrize("k", [3, 4, 10, 300, 1000])
def test_pose_svd_agrees_with_float64_kabsch(oracle, k):
    rng = np.random.default_rng(k)
    for _ in range(20):
        Tgt = S._rand_pose(rng, 30.0, 1.5)
        P = S._landmarks(rng, k)
        Q = P @ Tgt[:3, :3].T + Tgt[:3, 3] + rng.normal(size=(k, 3)) * 0.005
        T = oracle.pose_svd(P, Q)
        Tk = _kabsch(P, Q)
        assert np.linalg.norm(T[:3, 3] - Tk[:3, 3]) < 2e-5          # float32 round-off at ~10 m scale
        assert S.rot_angle(T[:3, :3], Tk[:3, :3]) < 5e-6
        R = T[:3, :3]
        assert np.abs(R @ R.T - np.eye(3)).max() < 5e-6 and abs(np.linalg.det(R) - 1) < 5e-6   # float32 Jacobi


def test_pose_svd_is_a_proper_rotation_on_degenerate_input(oracle):
    """collinear / duplicate samples: the reference has no degeneracy test, the solve must still return a finite
    proper rotation"""
    rng = np.random.default_rng(0)
    a = rng.normal(size=3)
    cases = [np.stack([a, a, a]), np.stack([a, 2 * a, 3 * a]), np.zeros((3, 3)), np.stack([a, a, -a])]
    for P in cases:
        Q = P + 0.5
        T = oracle.pose_svd(P, Q)
        assert np.isfinite(T).all()
        R = T[:3, :3]
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(R) - 1) < 1e-5


def test_consensus_is_strict_and_in_double(oracle):
    P = np.zeros((4, 3)); T = np.eye(4)
    Q = np.zeros((4, 3)); Q[:, 0] = [np.nextafter(0.1, 0), 0.1, np.nextafter(0.1, 1), 0.0]
    c, m = oracle.consensus3d(P, Q, T, 0.1)
    assert c == 2 and m.tolist() == [True, False, False, True]


def test_sample_list_structure(oracle):
    """growing-prefix shuffle of a persistent index vector (reference :214-225)"""
    M, I = 300, 100
    s = oracle.sample_list(M, I, True)
    for i in range(I):
        n = min(int(np.ceil(((i + 3.0) / I) * M)), M)
        assert len(set(s[i])) == 3 and s[i].max() < M
        # an index >= n can only sit in idx[0..2] if it was never inside a shuffled prefix: impossible
        assert (s[i] < max(n, 3)).all()
    # I=1000, M=300: iteration 0 has prefix ceil(0.9)=1 -> no shuffle at all -> the three best matches
    s = oracle.sample_list(300, 1000, True)
    assert s[0].tolist() == [0, 1, 2]
    # do_prosac=false shuffles the whole vector (TransformationFilter's call, transformation_filter.cpp:272)
    s = oracle.sample_list(100, 200, False)
    assert s.max() > 50
    # deterministic: rand() is restarted at seed 1 per call
    assert np.array_equal(oracle.sample_list(77, 50, True), oracle.sample_list(77, 50, True))


def test_prosac_sequential_semantics(oracle):
    """winner = first i with c[i] > max(c[:i]) that also satisfies c[i] >= 3 and c[i] > bp*M, else the first
    argmax (>= 3) — checked through the per-iteration counts the oracle exposes"""
    for seed, kw in [(1, {}), (2, dict(rho=0.9)), (3, dict(rho=0.05)), (4, dict(rho=0.7, gross_outlier_frac=0.0))]:
        f, t, _ = S.make_pair(400, seed=seed, **kw)
        for bp in (0.2, 0.6, 1.0):
            o = oracle.estimate_edge([f], [t], bp=bp)
            c = o["counts"]
            M = o["n_matches"]
            run = o["iterations_run"]
            assert (c[:run] >= 0).all() and (c[run:] == -1).all()
            best, mx, brk = -1, 0, None
            for i in range(run):
                if c[i] > mx:
                    mx, best = c[i], i
                    if mx >= 3 and mx > bp * M:
                        brk = i
                        break
            assert best == o["best_iteration"]
            assert run == (brk + 1 if brk is not None else 100)


def test_edge_recovers_planted_motion(oracle):
    for seed in range(5):
        f, t, Tgt = S.make_pair(1000, seed=100 + seed)
        o = oracle.estimate_edge([f], [t])
        assert o["ok"] and o["consensus"] > 100
        assert np.linalg.norm(o["T"][:3, 3] - Tgt[:3, 3]) < 0.05
        assert S.rot_angle(o["T"][:3, :3], Tgt[:3, :3]) < 0.02
        assert o["info_scale"] == pytest.approx(0.1 * o["consensus"] / o["mse"])


def test_failure_conventions(oracle):
    f, t, _ = S.make_pair(300, seed=9)
    assert not oracle.estimate_edge([], [t])["ok"]
    few = {k: (v[:6] if isinstance(v, np.ndarray) else v) for k, v in f.items()}
    o = oracle.estimate_edge([few], [t])
    assert not o["ok"] and o["consensus"] == 0 and o["cam_from"] == -1
    o = oracle.estimate_edge([f], [dict(t, sensor_frame=3)])
    assert not o["ok"] and o["cam_from"] == -1
    o = oracle.estimate_edge([f], [dict(t, valid=np.zeros_like(t["valid"]))])
    assert not o["ok"] and o["cam_from"] == 0 and o["n_matches"] == 0 and o["n_ratio_matches"] > 0
    assert np.array_equal(o["T"], np.eye(4))
