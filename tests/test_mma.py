"""The tensor-core match kernel (knn2_mmaf_kernel, uz_knn2_mmaf.cuh: Hamming distance as an exact contraction of 4-bit
operands, tcgen05.mma kind::mxf4) is the default for 256-bit rows; the int8 kernels before it stay selectable.  Each must
return exactly what cv::BFMatcher / the oracle / the integer-pipe kernel return: same neighbours, same tie rule (lowest train
index), on ragged query tiles, ragged train tiles (240-row tiles for the 4-bit kernel, 256 for int8), tiny and maximum sizes."""
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SIZES = [(1000, 1000), (400, 300), (65, 129), (1, 5), (513, 4096), (130, 127), (64, 2), (256, 256), (257, 255), (4096, 4096),
         (7, 1), (129, 257), (1000, 31), (9, 700),
         # the 240-row train tiles of the 4-bit kernel: full tiles, one row either side, both column halves ragged
         (256, 240), (300, 241), (300, 239), (128, 480), (200, 368), (260, 369), (31, 304), (500, 960), (256, 1201), (40, 112), (40, 113)]


def _est(mma, cfg=None):
    from uzliti_slam_b200 import EdgeEstimator
    os.environ["UZ_MATCH_MMA"] = str(mma)
    if cfg is not None:
        os.environ["UZ_MMA2_CFG"] = str(cfg)
    try:
        return EdgeEstimator(0)
    finally:
        os.environ.pop("UZ_MATCH_MMA", None)
        os.environ.pop("UZ_MMA2_CFG", None)


def test_mma_neighbours_equal_oracle_and_popc_kernel(oracle):
    em, ep = _est(1), _est(0)
    try:
        for nq, nt in SIZES:
            for mode in range(3):
                rng = np.random.default_rng(77 * nq + nt + mode)
                q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
                t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
                if mode == 1:                      # tie-heavy: distances collapse onto a few values
                    q[:, 2:] = 0; t[:, 2:] = 0
                if mode == 2 and nt > 4:           # exact duplicates across key blocks, stages and column halves
                    t[nt - 1] = q[0]; t[nt // 2] = q[0]; t[1] = q[nq // 2]; t[min(130, nt - 1)] = q[nq // 2]
                idx, dist = em.knnMatch(q, t)
                oi, od = oracle.knn2(q, t)
                assert np.array_equal(idx, oi), (nq, nt, mode)
                assert np.array_equal(dist, od), (nq, nt, mode)
                pi, pd = ep.knnMatch(q, t)
                assert np.array_equal(idx, pi) and np.array_equal(dist, pd)
        assert em.get_timers is not None
    finally:
        em.close(); ep.close()


def test_mma_all_zero_and_all_one_descriptors(oracle):
    """distance 0 and distance 256 are the ends of the key range (dot = +256 / -256)"""
    em = _est(1)
    try:
        q = np.zeros((300, 32), np.uint8)
        t = np.zeros((520, 32), np.uint8)
        t[::2] = 255
        q[10] = 255
        idx, dist = em.knnMatch(q, t)
        oi, od = oracle.knn2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
        assert dist.min() == 0 and od[0, 0] == 0
        t[:] = 255
        q[:] = 0
        idx, dist = em.knnMatch(q, t)
        assert (dist == 256).all() and (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all()
    finally:
        em.close()


def test_whole_path_records_identical_on_both_match_kernels(oracle):
    """store path, host path, rigs and the cross-check (reversed matching on the tensor cores, fused column minima on the
    integer pipes): byte-identical edge records"""
    em, ep = _est(1), _est(0)
    try:
        kfs, pairs, _ = S.make_map(60, n_features=700, cluster=6, pool=700, n_shared=400, k_candidates=6, cross_cluster=2, seed=21)
        for cross in (0, 1):
            recs = []
            for e in (em, ep):
                e.setConfig(cross_check=cross)
                e.clear()
                h = e.add_keyframes(kfs)
                recs.append(e.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]]))
            assert recs[0].tobytes() == recs[1].tobytes()
            assert (recs[0]["ok"] == 1).any()
            host = em.estimateEdgesHost([([kfs[a]], [kfs[b]]) for a, b in pairs[:40]])
            assert host.tobytes() == recs[0][:40].tobytes()
            for r, (a, b) in list(zip(recs[0], pairs))[:12]:
                o = oracle.estimate_edge([kfs[a]], [kfs[b]], cross_check=bool(cross))
                assert r["n_ratio_matches"] == o["n_ratio_matches"] and r["n_matches"] == o["n_matches"]
                assert r["consensus"] == o["consensus"]
        em.setConfig(cross_check=0); ep.setConfig(cross_check=0)
        # a rig with mixed widths: 256-bit cameras go to the tensor cores, 512-bit ones to knn2_wide_kernel, in one batch
        fa, ta, _ = S.make_pair(400, 300, seed=7, sensor_frame=0)
        fb, tb, _ = S.make_pair(350, 450, seed=8, sensor_frame=1, desc_bytes=64)
        a = em.estimateEdgesHost([([fa, fb], [ta, tb]), ([fa], [ta]), ([fb], [tb])])
        b = ep.estimateEdgesHost([([fa, fb], [ta, tb]), ([fa], [ta]), ([fb], [tb])])
        assert a.tobytes() == b.tobytes()
        o = oracle.estimate_edge([fa, fb], [ta, tb])
        assert a[0]["consensus"] == o["consensus"] and a[0]["cam_from"] == o["cam_from"]
    finally:
        em.close(); ep.close()


@pytest.mark.parametrize("n_features", [230, 480, 700, 950])
def test_many_items_per_cta_with_one_to_four_train_tiles(n_features):
    """the shared-memory rings of knn2_mmaf_kernel turn over many times per CTA: 1, 2, 3 and 4 train tiles per item against
    4 train stages and 2 query-tile buffers (a query tile is released by the stage barrier of its item's last train tile - with
    three tiles per item that stage is refilled by the very next tile).  Records byte-identical to the integer-pipe kernel's."""
    em, ep = _est(1), _est(0)
    try:
        kfs, pairs, _ = S.make_map(240, n_features=n_features, cluster=12, pool=n_features, n_shared=n_features // 2,
                                   k_candidates=12, cross_cluster=2, seed=100 + n_features)
        recs = []
        for e in (em, ep):
            h = e.add_keyframes(kfs)
            recs.append([e.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]]) for _ in range(2)])
        assert recs[0][0].tobytes() == recs[1][0].tobytes() and recs[0][1].tobytes() == recs[1][1].tobytes()
        assert (recs[0][0]["ok"] == 1).any() and len(pairs) > 2000
    finally:
        em.close(); ep.close()


@pytest.mark.parametrize("variant", [4, 7])
def test_alternative_tensor_core_kernels_equal_oracle(oracle, variant):
    """the measured alternatives that stay selectable (UZ_MATCH_MMA=4: int8 operands, keys formed by the MMA, knn2_mmak_kernel;
    7: the IMAD epilogue of knn2_mma_kernel)"""
    e = _est(variant)
    try:
        for nq, nt in SIZES:
            for mode in range(3):
                rng = np.random.default_rng(55 * nq + nt + mode)
                q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
                t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
                if mode == 1:
                    q[:, 2:] = 0; t[:, 2:] = 0
                if mode == 2 and nt > 4:
                    t[nt - 1] = q[0]; t[nt // 2] = q[0]; t[1] = q[nq // 2]; t[min(130, nt - 1)] = q[nq // 2]
                idx, dist = e.knnMatch(q, t)
                oi, od = oracle.knn2(q, t)
                assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt, mode)
    finally:
        e.close()


def test_cta_pair_kernel_neighbours_equal_oracle(oracle, cfg=None):
    """knn2_mma2_kernel (tcgen05.mma.cta_group::2 on CTA pairs; UZ_MATCH_MMA=3 forces it for every launch): the same shapes,
   """
    e2 = _est(3, cfg)
    try:
        for nq, nt in SIZES + [(512, 256), (511, 513), (1024, 1000), (640, 128), (2000, 1500)]:
            for mode in range(3):
                rng = np.random.default_rng(99 * nq + nt + mode)
                q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
                t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
                if mode == 1:
                    q[:, 2:] = 0; t[:, 2:] = 0
                if mode == 2 and nt > 4:
                    t[nt - 1] = q[0]; t[nt // 2] = q[0]; t[1] = q[nq // 2]; t[min(130, nt - 1)] = q[nq // 2]
                idx, dist = e2.knnMatch(q, t)
                oi, od = oracle.knn2(q, t)
                assert np.array_equal(idx, oi), (nq, nt, mode)
                assert np.array_equal(dist, od), (nq, nt, mode)
    finally:
        e2.close()


def test_cta_pair_kernel_whole_batches_identical(oracle, cfg=None):
    """batches large enough to keep every CTA pair busy for many items (ragged sizes, rigs, cross-check): records byte-identical
    to the one-CTA tensor-core kernel's"""
    e2, e1 = _est(2, cfg), _est(1)
    try:
        kfs, pairs, _ = S.make_map(120, n_features=1000, cluster=10, pool=1000, n_shared=600, k_candidates=10, cross_cluster=2, seed=23)
        ragged, rp, _ = S.make_map(60, n_features=700, cluster=6, pool=700, n_shared=400, k_candidates=6, cross_cluster=2, seed=21)
        for cross in (0, 1):
            for ks, ps in ((kfs, pairs), (ragged, rp)):
                recs = []
                for e in (e2, e1):
                    e.setConfig(cross_check=cross)
                    e.clear()
                    h = e.add_keyframes(ks)
                    recs.append(e.estimateEdges(h[ps[:, 0]], h[ps[:, 1]]))
                assert recs[0].tobytes() == recs[1].tobytes()
                assert (recs[0]["ok"] == 1).any()
        host = e2.estimateEdgesHost([([kfs[a]], [kfs[b]]) for a, b in pairs[:300]])
        e1.setConfig(cross_check=1)
        assert host.tobytes() == e1.estimateEdgesHost([([kfs[a]], [kfs[b]]) for a, b in pairs[:300]]).tobytes()
    finally:
        e2.close(); e1.close()


@pytest.mark.parametrize("width", [32, 64])
def test_tensor_core_kernels_fuzz(oracle, width):
    """random shapes and bit densities (sparse / dense rows move the distances towards 0 and towards the maximum), both widths:
    the tensor-core kernels against the oracle"""
    from uzliti_slam_b200 import EdgeEstimator
    est = EdgeEstimator(0)
    rng = np.random.default_rng(20260 + width)
    try:
        for trial in range(48):
            nq = int(rng.integers(1, 1400)) if trial % 6 else int(rng.choice([127, 128, 129, 255, 256, 257, 511, 512, 513, 1024]))
            nt = int(rng.integers(1, 1400)) if trial % 5 else int(rng.choice([127, 128, 129, 255, 256, 257, 383, 384, 385, 1025]))
            p = float(rng.choice([0.02, 0.3, 0.5, 0.5, 0.7, 0.98]))
            q = np.packbits(rng.random((nq, width * 8)) < p, axis=1)
            t = np.packbits(rng.random((nt, width * 8)) < (1.0 - p if trial % 3 == 0 else p), axis=1)
            if trial % 4 == 0 and nt > 8:                          # duplicated train rows: ties on (distance, index)
                t[rng.integers(0, nt, 8)] = t[rng.integers(0, nt)]
            idx, dist = est.knnMatch(q, t)
            oi, od = oracle.knn2(q, t)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od), (width, nq, nt, p)
    finally:
        est.close()
