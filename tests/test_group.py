"""uz_group: the batched path on several GPUs in ONE process behind the C-ABI (replicated store pulled over peer memory,
contiguous pair shards, records written by every device's solve kernel straight into one buffer).  SURVEY section 4: "same
batch on 1/2/4/8 devices -> identical per-pair outputs" - here byte for byte, on the hardware."""
import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def test_group_of_one_device_equals_plain_context(est):
    from uzliti_slam_b200 import GroupEstimator
    kfs, pairs, _ = S.make_map(80, n_features=500, cluster=8, pool=500, n_shared=300, k_candidates=6, cross_cluster=2, seed=31)
    est.clear()
    h = est.add_keyframes(kfs)
    want = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    g = GroupEstimator([0])
    try:
        hg = g.add_keyframes(kfs)
        got = g.estimateEdges(hg[pairs[:, 0]], hg[pairs[:, 1]])
        assert got.tobytes() == want.tobytes()
        g.set_gather(1)
        assert g.estimateEdges(hg[pairs[:, 0]], hg[pairs[:, 1]]).tobytes() == want.tobytes()
    finally:
        g.close()
    est.clear()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_group_records_equal_single_gpu_byte_for_byte(est, ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import torch
    from uzliti_slam_b200 import GroupEstimator
    from uzliti_slam_b200.binding import RESULT_DTYPE
    kfs, pairs, _ = S.make_map(400, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=10, cross_cluster=3, seed=33)
    pairs = pairs[:2500]
    est.clear()
    h = est.add_keyframes(kfs)
    want = est.estimateEdges(h[pairs[:, 0]], h[pairs[:, 1]])
    g = GroupEstimator(list(range(ndev)))
    try:
        hg = g.add_keyframes(kfs)
        assert np.array_equal(hg, h - h[0] + hg[0])
        assert g.store_size() == len(kfs)
        for mode in (0, 1):
            g.set_gather(mode)
            got = g.estimateEdges(hg[pairs[:, 0]], hg[pairs[:, 1]])
            assert got.tobytes() == want.tobytes(), f"gather mode {mode}"
        # records in the first device's memory, written over NVLink by the other devices' solve kernels
        buf = torch.zeros(len(pairs) * RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda:0")
        g.estimateEdgesDevice(hg[pairs[:, 0]], hg[pairs[:, 1]], buf.data_ptr())
        assert buf.cpu().numpy().tobytes() == want.tobytes()
        # every replica holds the same keyframes (pulled from device 0 over peer memory)
        for r in range(ndev):
            back = g.context(r).read_keyframe(int(hg[17]))
            assert np.array_equal(back["desc"], kfs[17]["desc"]) and np.array_equal(back["pos"], kfs[17]["pos"])
        # uneven shards, a batch smaller than the group, remove on every device
        few = g.estimateEdges(hg[pairs[:3, 0]], hg[pairs[:3, 1]])
        assert few.tobytes() == want[:3].tobytes()
        g.remove_keyframe(int(hg[399]))
        assert g.store_size() == len(kfs) - 1
        assert (g.last_timing() >= 0).all()
    finally:
        g.close()
    est.clear()


@pytest.mark.gpu
def test_group_begin_end_equals_the_synchronous_call(built):
    """uz_group_estimate_edges_begin / _end: the batch runs on the group's device workers while the caller does something else;
    same records as the one-call form, and the group refuses other work in between"""
    import torch
    from uzliti_slam_b200 import GroupEstimator
    from uzliti_slam_b200.binding import UzError
    n = min(torch.cuda.device_count(), 2)
    kfs, pairs, _ = S.make_map(60, n_features=500, cluster=6, pool=500, n_shared=300, k_candidates=6, cross_cluster=2, seed=35)
    g = GroupEstimator(list(range(n)))
    try:
        h = g.add_keyframes(kfs)
        f, t = h[pairs[:, 0]], h[pairs[:, 1]]
        want = g.estimateEdges(f, t)
        g.estimateEdgesBegin(f, t)
        with pytest.raises(UzError):
            g.estimateEdges(f[:4], t[:4])                # a batch is in flight
        with pytest.raises(UzError):
            g.add_keyframes(kfs[:1])
        busy = sum(i * i for i in range(20000))          # the caller's own work
        got = g.estimateEdgesEnd()
        assert busy > 0 and got.tobytes() == want.tobytes() and (got["ok"] == 1).any()
        g.estimateEdgesBegin(f[:0], t[:0])               # empty batch: nothing in flight
        assert len(g.estimateEdgesEnd()) == 0
        assert g.estimateEdges(f, t).tobytes() == want.tobytes()
    finally:
        g.close()
