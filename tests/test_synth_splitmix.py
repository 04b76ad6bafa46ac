"""The benchmark's synthetic map: include/uz_synth.h (C) and uzliti_slam_b200/synth_splitmix.py (numpy) are one generator in two
languages - same bytes - and the scene has the geometry the rest of the suite assumes."""
import ctypes
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synth_splitmix as SM
from uzliti_slam_b200 import synthetic as S


@pytest.fixture(scope="module")
def native():
    SM.build_native()
    return ctypes.CDLL(SM._LIB)


@pytest.mark.parametrize("cfg", [
    dict(n_keyframes=30, n_features=200, cluster=6, pool=200, n_shared=120, k_candidates=5, cross_cluster=2, seed=7),
    dict(n_keyframes=11, n_features=150, cluster=4, pool=300, n_shared=90, k_candidates=3, cross_cluster=1, seed=1, desc_bytes=64),  # ragged last cluster
    dict(n_keyframes=5, n_features=64, cluster=8, pool=64, n_shared=64, k_candidates=6, cross_cluster=2, seed=3),      # one cluster: no cross pairs
    dict(n_keyframes=40, n_features=333, cluster=25, pool=1000, n_shared=0, k_candidates=20, cross_cluster=4, seed=4, invalid_frac=0.5),
])
def test_c_header_and_numpy_mirror_fill_identical_bytes(native, cfg):
    k1, p1, T1 = SM.make_map(native=True, threads=3, **cfg)
    k2, p2, T2 = SM.make_map(native=False, **cfg)
    assert len(k1) == len(k2) == cfg["n_keyframes"]
    for a, b in zip(k1, k2):
        assert a["desc"].tobytes() == b["desc"].tobytes()
        assert a["pos"].tobytes() == b["pos"].tobytes()          # float64 bit for bit: only IEEE + - * / sqrt in a fixed order
        assert a["valid"].tobytes() == b["valid"].tobytes()
        assert a["feature_type"] == b["feature_type"]
    assert p1.tobytes() == p2.tobytes() and T1.tobytes() == T2.tobytes()
    native.uz_synth_checksum_c.restype = ctypes.c_ulonglong
    native.uz_synth_checksum_c.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    for arr in (np.stack([k["desc"] for k in k2]), np.stack([k["pos"] for k in k2]), p2, np.arange(13, dtype=np.uint8)):
        arr = np.ascontiguousarray(arr)
        assert native.uz_synth_checksum_c(arr.ctypes.data, arr.nbytes) == SM.checksum(arr)


def test_out_arrays_receive_the_map():
    desc = np.empty((6, 100, 32), np.uint8); pos = np.empty((6, 100, 3)); valid = np.empty((6, 100), np.uint8)
    kfs, pairs, _ = SM.make_map(6, 100, 3, 100, 60, 2, 1, seed=2, out=(desc, pos, valid), native=False)
    ref, pr, _ = SM.make_map(6, 100, 3, 100, 60, 2, 1, seed=2, native=False)
    for i, k in enumerate(ref):
        assert (desc[i] == k["desc"]).all() and pos[i].tobytes() == k["pos"].tobytes() and (valid[i] == k["valid"]).all()
        assert kfs[i]["desc"].base is not None
    assert (pairs == pr).all()
    with pytest.raises(ValueError):
        SM.keyframe(0, n_features=10, pool=5, n_shared=8)


def test_scene_statistics_and_geometry():
    from oracle import binding as O
    kfs, pairs, poses = SM.make_map(50, native=False)
    v = np.concatenate([k["valid"] for k in kfs])
    assert abs(v.mean() - 0.85) < 0.01
    z = np.concatenate([k["pos"][k["valid"] != 0, 2] for k in kfs])
    assert z.min() > 0.0 and z.max() < 9.0
    assert all((k["pos"][k["valid"] == 0] == (0.0, 0.0, -1.0)).all() for k in kfs)
    assert len(pairs) == 50 * 20 and (pairs[:, 0] != pairs[:, 1]).all()
    same = pairs[(pairs[:, 0] // 25) == (pairs[:, 1] // 25)]
    cross = pairs[(pairs[:, 0] // 25) != (pairs[:, 1] // 25)]
    assert len(same) == 50 * 16 and len(cross) == 50 * 4
    # planted loop closures are found with the planted motion; cross-cluster pairs are not edges
    good = 0
    for a, b in same[:12]:
        r = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        assert r["ok"]
        gt = S.gt_transform(poses, a, b)
        assert S.rot_angle(gt[:3, :3], np.asarray(r["T"])[:3, :3]) < 0.02 and np.abs(gt[:3, 3] - np.asarray(r["T"])[:3, 3]).max() < 0.05
        good += r["consensus"] >= 100
        assert r["consensus"] < 0.6 * r["n_matches"]            # below ransac_break_percentage: all 100 hypotheses run, as on C4
    assert good == 12
    for a, b in cross[:6]:
        r = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        assert (not r["ok"]) or r["consensus"] < 15
