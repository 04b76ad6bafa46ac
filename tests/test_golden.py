"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from cv2 + the real
std::random_shuffle/rand()).  CPU: the oracle must reproduce them.  GPU: the CUDA path must too."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EDGE_FILES = sorted(glob.glob(os.path.join(HERE, "golden", "edge_*.npz")))


def _kf(d, side):
    return dict(desc=d[f"{side}_desc"], pos=d[f"{side}_pos"], valid=d[f"{side}_valid"], feature_type=2, sensor_frame=0)


def _check(d, r, matches, mask, counts):
    assert int(bool(r["ok"])) == int(d["ok"]) and r["n_ratio_matches"] == d["n_ratio"] and r["n_matches"] == d["n_matches"]
    assert r["consensus"] == d["consensus"] and r["best_iteration"] == d["best_iteration"]
    assert r["iterations_run"] == d["iterations_run"]
    assert np.array_equal(matches, d["matches"])
    if d["ok"]:
        assert np.array_equal(np.asarray(mask, np.uint8), d["inlier_mask"])
        assert np.array_equal(counts, d["counts"])
    T = np.asarray(r["T"]).reshape(4, 4)
    assert np.abs(T - d["T"]).max() <= 1e-5          # north_star tolerance; in practice 0
    if np.isnan(d["mse"]):
        assert np.isnan(r["mse"])
    else:
        assert abs(r["mse"] - d["mse"]) <= 1e-12


def test_fixtures_exist():
    assert len(EDGE_FILES) >= 5 and os.path.exists(os.path.join(HERE, "golden", "sample_lists.npz"))


@pytest.mark.parametrize("path", EDGE_FILES, ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_reproduces_golden(oracle, path):
    d = np.load(path)
    oi, od = oracle.knn2(d["t_desc"], d["f_desc"])
    assert np.array_equal(oi, d["cv2_idx"]) and np.array_equal(od, d["cv2_dist"])       # pinned to OpenCV
    o = oracle.estimate_edge([_kf(d, "f")], [_kf(d, "t")])
    _check(d, o, o["matches"], o["inlier_mask"], o["counts"])


def test_oracle_sample_lists_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "sample_lists.npz"))
    for key in g.files:
        if key.startswith("M"):
            M, I, p = (int(x[1:]) for x in key.split("_"))
            assert np.array_equal(oracle.sample_list(M, I, bool(p)), g[key]), key
    assert np.array_equal(oracle.glibc_rand(64), g["rand_seed1_first64"])
    assert g["rand_seed1_first64"][0] == 1804289383          # the well-known first rand() of glibc


REF_STAGE3 = os.path.join(HERE, "golden", "ref_stage3.npz")


def _run_oracle_case(oracle, kind, P, Q, par, Tin):
    thr, it, bp, prosac = float(par[0]), int(par[1]), float(par[2]), bool(par[3])
    if kind == 0:
        r = oracle.estimate_svd(P, Q, thr, it, bp, prosac)
        return r["T"], r["consensus"], r["mse"], r["mask"]
    if kind == 1:
        return oracle.pose_svd(P, Q), 0, 0.0, np.zeros(len(P), bool)
    c, mask = oracle.consensus3d(P, Q, Tin, thr)
    return Tin, c, 0.0, mask


def test_reference_stage3_cases_run_through_the_oracle(oracle):
    """the Python half of the recipe (oracle/make_ref_golden.py) works here: cases are generated and the oracle solves them"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(os.path.dirname(HERE), "oracle", "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = mod.make_cases()
    assert len(cases) >= 60 and {c[0] for c in cases} == {0, 1, 2}
    for kind, P, Q, thr, it, bp, prosac, T in cases[:20] + cases[-14:]:
        Tout, cons, mse, mask = _run_oracle_case(oracle, kind, P, Q, (thr, it, bp, prosac), T)
        assert np.isfinite(np.asarray(Tout)).all() and 0 <= cons <= len(P) and mask.sum() == (cons if kind != 1 else 0)


def test_oracle_matches_reference_stage3(oracle):
    """Stage 3 pinned to the reference's own binaries - when someone has produced the vectors (oracle/build_ref.sh needs
    Eigen 3 / PCL / Boost headers, which this image does not have).  Until then: parity unpinned."""
    if not os.path.exists(REF_STAGE3):
        pytest.skip("tests/golden/ref_stage3.npz not present: stage 3 (PCL / Eigen arithmetic) is parity-unpinned; "
                    "run oracle/build_ref.sh + oracle/make_ref_golden.py where Eigen 3 / PCL headers exist")
    g = np.load(REF_STAGE3)
    for i in range(int(g["n_cases"])):
        kind = int(g[f"c{i}_kind"])
        Tout, cons, mse, mask = _run_oracle_case(oracle, kind, g[f"c{i}_P"], g[f"c{i}_Q"], g[f"c{i}_par"], g[f"c{i}_Tin"])
        assert np.array_equal(np.asarray(Tout), g[f"c{i}_T"]), f"case {i}: transform differs from the reference's"
        if kind != 1:
            assert cons == int(g[f"c{i}_consensus"]) and np.array_equal(mask.astype(np.uint8), g[f"c{i}_mask"]), f"case {i}"
        if kind == 0:
            ref_mse = float(g[f"c{i}_mse"])
            assert (np.isnan(mse) and np.isnan(ref_mse)) or mse == ref_mse, f"case {i}"


@pytest.mark.gpu
@pytest.mark.parametrize("path", EDGE_FILES, ids=lambda p: os.path.basename(p)[5:-4])
def test_gpu_reproduces_golden(est, path):
    d = np.load(path)
    idx, dist = est.knnMatch(d["t_desc"], d["f_desc"])
    assert np.array_equal(idx, d["cv2_idx"]) and np.array_equal(dist, d["cv2_dist"])    # bit-exact vs OpenCV
    est.set_debug(True)
    r = est.estimateEdgeDirect([_kf(d, "f")], [_kf(d, "t")])
    m, mask = est.debug_pair(0, r["n_matches"])
    counts = est.debug_counts(0) if d["ok"] else None
    est.set_debug(False)
    _check(d, r, m, mask, counts)


@pytest.mark.gpu
def test_gpu_sample_lists_golden(est):
    g = np.load(os.path.join(HERE, "golden", "sample_lists.npz"))
    for key in g.files:
        if key.startswith("M"):
            M, I, p = (int(x[1:]) for x in key.split("_"))
            assert np.array_equal(est.sample_list(M, I, bool(p)), g[key]), key
