"""The streaming form of the solve (solve_stream_kernel: a persistent grid beside the match kernel, fed through per-pair
completion counters) must produce exactly the records of the one-CTA-per-pair form, which test_gpu_parity.py pins to the
oracle.  UZ_STREAM_SOLVE / UZ_STREAM_SOLVE_MIN_PAIRS are read by uz_create, so each form gets its own context.
"""
import os

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _estimator(**env):
    from uzliti_slam_b200 import EdgeEstimator
    # the streaming solve and the segmented small launches are forms of the INTEGER-PIPE match kernels (a tensor-core match
    # CTA fills its SM's shared memory, nothing runs beside it): tests of those forms switch the 256-bit rows back to them
    env.setdefault("UZ_MATCH_MMA", 0)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return EdgeEstimator(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ragged_map(seed, n_keyframes=120):
    """keyframes of different sizes (1-3 match tiles per pair), some too small to be matched (< 7 keypoints)"""
    kfs, pairs, _ = S.make_map(n_keyframes, n_features=600, cluster=12, pool=600, n_shared=350, k_candidates=10,
                               cross_cluster=3, seed=seed)
    rng = np.random.default_rng(seed)
    out = []
    for i, kf in enumerate(kfs):
        n = [600, 520, 300, 64, 5][i % 5] if i % 7 else 600
        out.append({k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in kf.items()})
    return out, pairs[rng.permutation(len(pairs))]


@pytest.mark.parametrize("ctas", [1, 2])
def test_streaming_equals_one_cta_per_pair(ctas):
    kfs, pairs = _ragged_map(5)
    plain = _estimator(UZ_STREAM_SOLVE=0)
    stream = _estimator(UZ_STREAM_SOLVE=ctas, UZ_STREAM_SOLVE_MIN_PAIRS=1)
    try:
        hp = plain.add_keyframes(kfs)
        hs = stream.add_keyframes(kfs)
        for cross in (0, 1):
            plain.setConfig(cross_check=cross)
            stream.setConfig(cross_check=cross)
            a = plain.estimateEdges(hp[pairs[:, 0]], hp[pairs[:, 1]])
            n0 = stream.launch_count()
            b = stream.estimateEdges(hs[pairs[:, 0]], hs[pairs[:, 1]])
            assert stream.launch_count() - n0 == 3          # match launch + persistent solve launch + (empty) cleanup launch
            assert a.tobytes() == b.tobytes()
            assert (a["ok"] == 0).any() and (a["consensus"] > 50).any()
            # twice in a row (double-buffered staging, counters re-armed), and a tiny batch
            assert stream.estimateEdges(hs[pairs[:, 0]], hs[pairs[:, 1]]).tobytes() == a.tobytes()
            assert stream.estimateEdges(hs[pairs[:3, 0]], hs[pairs[:3, 1]]).tobytes() == a[:3].tobytes()
    finally:
        plain.close()
        stream.close()


def test_streaming_against_oracle_with_taps(oracle):
    kfs, pairs, _ = S.make_map(24, n_features=500, cluster=6, pool=500, n_shared=300, k_candidates=4, cross_cluster=1, seed=21)
    est = _estimator(UZ_STREAM_SOLVE=1, UZ_STREAM_SOLVE_MIN_PAIRS=1)
    try:
        h = est.add_keyframes(kfs)
        est.set_debug(True)
        res = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        for i, (a, b) in enumerate(pairs):
            o = oracle.estimate_edge([kfs[a]], [kfs[b]])
            assert bool(res[i]["ok"]) == o["ok"] and res[i]["consensus"] == o["consensus"]
            assert res[i]["best_iteration"] == o["best_iteration"] and res[i]["iterations_run"] == o["iterations_run"]
            assert np.abs(res[i]["T"].reshape(4, 4) - o["T"]).max() < 1e-5
            m, mask = est.debug_pair(i, res[i]["n_matches"])
            assert np.array_equal(m, o["matches"])
            if o["ok"]:
                assert np.array_equal(mask, o["inlier_mask"])
    finally:
        est.close()


def test_streaming_host_path_equals_store_path():
    kfs, pairs, _ = S.make_map(60, n_features=500, cluster=10, pool=500, n_shared=300, k_candidates=8, cross_cluster=2, seed=4)
    est = _estimator(UZ_STREAM_SOLVE=1, UZ_STREAM_SOLVE_MIN_PAIRS=1, UZ_HOST_CHUNKS=3)
    try:
        h = est.add_keyframes(kfs)
        a = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        b = est.estimateEdgesHost([([kfs[i]], [kfs[j]]) for i, j in pairs])
        assert a.tobytes() == b.tobytes()
    finally:
        est.close()


def test_streaming_survives_starvation():
    """UZ_STREAM_PROBE=9 holds the match kernel back for 3 x the starvation limit: the streaming CTAs must give up, free
    the SMs, and the cleanup launch behind the match kernel must solve every pair - same records as the plain form."""
    kfs, pairs = _ragged_map(8, n_keyframes=60)
    plain = _estimator(UZ_STREAM_SOLVE=0)
    starved = _estimator(UZ_STREAM_SOLVE=1, UZ_STREAM_SOLVE_MIN_PAIRS=1, UZ_STREAM_PROBE=9)
    try:
        hp = plain.add_keyframes(kfs)
        hs = starved.add_keyframes(kfs)
        a = plain.estimateEdges(hp[pairs[:, 0]], hp[pairs[:, 1]])
        for _ in range(2):
            assert starved.estimateEdges(hs[pairs[:, 0]], hs[pairs[:, 1]]).tobytes() == a.tobytes()
    finally:
        plain.close()
        starved.close()


@pytest.mark.parametrize("chunks,alt,stream,mma", [(2, 1, 1, 0), (5, 1, 1, 0), (7, 0, 1, 0), (5, 1, 0, 0), (16, 1, 1, 0),
                                                   (2, 1, 1, 1), (5, 1, 1, 1), (7, 0, 1, 1), (16, 1, 0, 1)])
def test_chunked_host_pipeline_equals_store_path(chunks, alt, stream, mma):
    """uz_estimate_edges_host cuts a batch into chunks that upload, match and solve on their own, consecutive chunks on two
    alternating compute streams (UZ_ALT_CHUNKS=0: one stream).  A later chunk reuses cameras an earlier chunk uploaded
    (every keyframe appears in many pairs, pairs shuffled), ragged sizes, both descriptor widths, pageable host memory."""
    kn, pn = _ragged_map(11, n_keyframes=60)
    kw, pw, _ = S.make_map(30, n_features=400, cluster=10, pool=400, n_shared=250, k_candidates=6, cross_cluster=2, seed=12,
                           desc_bytes=64)
    kfs = kn + kw
    pairs = np.concatenate([pn, pw + len(kn)])
    pairs = pairs[np.random.default_rng(3).permutation(len(pairs))]
    est = _estimator(UZ_STREAM_SOLVE=stream, UZ_STREAM_SOLVE_MIN_PAIRS=1, UZ_HOST_CHUNKS=chunks, UZ_ALT_CHUNKS=alt, UZ_MATCH_MMA=mma)
    try:
        h = est.add_keyframes(kfs)
        a = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
        for _ in range(2):
            b = est.estimateEdgesHost([([kfs[i]], [kfs[j]]) for i, j in pairs])
            assert a.tobytes() == b.tobytes()
        assert (a["consensus"] > 50).sum() > 100 and (a["ok"] == 0).any()
    finally:
        est.close()


def test_small_launch_forms_agree(oracle):
    """Launches that cannot fill the chip take two extra forms - train rows cut into segments (SEG match kernels + merge) and
    a 512-thread solve CTA: records must equal the plain forms (UZ_SEGMENT=0, UZ_SOLVE_WIDE=0) and the oracle."""
    kfs, pairs = _ragged_map(13, n_keyframes=36)
    kw, pw, _ = S.make_map(12, n_features=700, cluster=6, pool=700, n_shared=400, k_candidates=3, cross_cluster=1, seed=14,
                           desc_bytes=64)
    kfs = kfs + kw
    pairs = np.concatenate([pairs[:40], pw + 36])
    fast = _estimator(UZ_STREAM_SOLVE=0)
    plain = _estimator(UZ_STREAM_SOLVE=0, UZ_SEGMENT=0, UZ_SOLVE_WIDE=0)
    tensor = _estimator(UZ_MATCH_MMA=1)        # 256-bit rows on the tensor cores, 512-bit rows segmented beside them
    try:
        hf, hp = fast.add_keyframes(kfs), plain.add_keyframes(kfs)
        ht = tensor.add_keyframes(kfs)
        for n in (1, 7, len(pairs)):
            assert tensor.estimateEdges(ht[pairs[:n, 0]], ht[pairs[:n, 1]]).tobytes() == \
                plain.estimateEdges(hp[pairs[:n, 0]], hp[pairs[:n, 1]]).tobytes()
        for cross in (0, 1):
            fast.setConfig(cross_check=cross)
            plain.setConfig(cross_check=cross)
            for n in (1, 7, len(pairs)):
                a = fast.estimateEdges(hf[pairs[:n, 0]], hf[pairs[:n, 1]])
                b = plain.estimateEdges(hp[pairs[:n, 0]], hp[pairs[:n, 1]])
                assert a.tobytes() == b.tobytes(), (cross, n)
            for i in (0, 5, len(pairs) - 1):
                o = oracle.estimate_edge([kfs[pairs[i, 0]]], [kfs[pairs[i, 1]]], cross_check=bool(cross))
                assert a[i]["consensus"] == o["consensus"] and a[i]["n_matches"] == o["n_matches"]
    finally:
        fast.close()
        plain.close()
        tensor.close()
