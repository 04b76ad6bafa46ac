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
