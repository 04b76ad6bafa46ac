"""CPU: the PRODUCT's arithmetic spec (csrc/uz_arith.cuh) and sample-list generator (csrc/uz_samples.h),
compiled for the host by tests/host_shim, against the independently written oracle — bit for bit.  This is
what lets "identical inlier sets" be checked before any GPU time is spent."""
import ctypes as C
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(built):
    lib = C.CDLL(os.path.join(HERE, "host_shim", "libhost_shim.so"))
    lib.hs_residual_sq.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_glibc_rand_reimplementation(shim, oracle):
    for seed, n in ((1, 200000), (12345, 5000), (0, 1000)):
        out = np.empty(n, np.int32)
        shim.hs_glibc_rand(seed, n, _p(out))
        assert np.array_equal(out, oracle.glibc_rand(n, seed=seed if seed else 1) if seed else oracle.glibc_rand(n, seed=0))


@pytest.mark.parametrize("iters,prosac,cap", [(100, 1, 300), (100, 0, 128), (1000, 1, 260), (1, 1, 64), (7, 1, 50),
                                              (200, 0, 100)])
def test_sample_table_equals_real_random_shuffle(shim, oracle, iters, prosac, cap):
    t = np.zeros((cap + 1, iters, 3), np.uint16)
    shim.hs_sample_table(iters, prosac, cap, _p(t))
    for M in sorted(set([3, 4, 5, 17, cap // 2, cap - 1, cap])):
        assert np.array_equal(t[M].astype(np.int32), oracle.sample_list(M, iters, bool(prosac))), M
    assert (t[:3] == 0).all()


def test_pose_bit_exact_with_oracle(shim, oracle):
    rng = np.random.default_rng(0)
    for trial in range(6000):
        k = 3 if trial % 2 == 0 else int(rng.integers(3, 80))
        P = rng.normal(size=(k, 3)) * rng.choice([0.01, 1.0, 5.0])
        if trial % 7 == 0:
            P[2] = P[0] + (P[1] - P[0]) * 0.5                   # collinear sample
        if trial % 11 == 0:
            P[1] = P[0]                                          # duplicate point
        if trial % 13 == 0:
            P[:] = P[0]                                          # rank 0
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        Q = P @ R.T + rng.normal(size=3) + rng.normal(size=(k, 3)) * 0.01
        P = np.ascontiguousarray(P); Q = np.ascontiguousarray(Q)
        T2 = np.empty(16)
        shim.hs_pose(_p(P), _p(Q), k, _p(T2))
        T1 = oracle.pose_svd(P, Q)
        assert np.array_equal(T1.ravel().view(np.uint64), T2.view(np.uint64)), trial


def test_residual_bit_exact_with_oracle(shim, oracle):
    rng = np.random.default_rng(1)
    T = S._rand_pose(rng, 30, 1.5)
    T32 = T.astype(np.float32).astype(np.float64)
    P = S._landmarks(rng, 2000); Q = S._landmarks(rng, 2000)
    for thr in (0.1, 1.0, 3.0):
        _, mask = oracle.consensus3d(P, Q, T32, thr)
        s = np.array([shim.hs_residual_sq(_p(np.ascontiguousarray(T32)), _p(np.ascontiguousarray(P[i])),
                                           _p(np.ascontiguousarray(Q[i]))) for i in range(len(P))])
        assert np.array_equal(np.sqrt(s) < thr, mask)
