"""CPU, world_size 2 over gloo: the N>1 path of the harness — pair-list sharding and the all-gather of the
fixed-size edge records (the only exchange step of the path)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from uzliti_slam_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 200000, 25001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import binding as O
    from uzliti_slam_b200 import synthetic as S
    from uzliti_slam_b200.binding import RESULT_DTYPE
    from uzliti_slam_b200.sharding import gather_records, shard_pairs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kfs, pairs, _ = S.make_map(12, n_features=120, cluster=4, pool=120, n_shared=70, k_candidates=3, cross_cluster=1, seed=2)
    # store maintenance: only rank 0 "ingests" the map, every rank ends up with identical keyframes
    from uzliti_slam_b200.sharding import broadcast_keyframes
    ragged = [dict(k, desc=k["desc"][:100 + i], pos=k["pos"][:100 + i], valid=k["valid"][:100 + i], sensor_frame=i % 3) for i, k in enumerate(kfs)]
    # BRISK / FREAK keyframes (64-byte rows) between the ORB ones: widths travel per keyframe
    wide, _, _ = S.make_map(3, n_features=50, cluster=3, pool=50, n_shared=30, k_candidates=1, seed=9, desc_bytes=64)
    ragged = ragged[:5] + wide + ragged[5:]
    got = broadcast_keyframes(ragged if rank == 0 else None, src=0)
    assert len(got) == len(ragged)
    for a, b in zip(got, ragged):
        assert np.array_equal(a["desc"], b["desc"]) and a["pos"].tobytes() == b["pos"].tobytes() and np.array_equal(a["valid"], b["valid"])
        assert a["sensor_frame"] == b["sensor_frame"] and a["feature_type"] == b["feature_type"]
        assert a["desc"].shape == b["desc"].shape
    mine, lo = shard_pairs(pairs, world, rank)
    rec = np.zeros(len(mine), RESULT_DTYPE)
    for i, (a, b) in enumerate(mine):       # the per-rank compute step, here through the oracle (CPU test)
        o = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        rec[i]["ok"] = o["ok"]; rec[i]["consensus"] = o["consensus"]; rec[i]["n_matches"] = o["n_matches"]
        rec[i]["T"] = o["T"].ravel(); rec[i]["mse"] = o["mse"]; rec[i]["best_iteration"] = lo + i
    full = gather_records(rec, len(pairs))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full.tobytes(), len(pairs)))


def test_two_rank_gather_equals_single_process():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1]                      # every rank ends up with the same full array
    sys.path.insert(0, ROOT)
    from oracle import binding as O
    from uzliti_slam_b200 import synthetic as S
    from uzliti_slam_b200.binding import RESULT_DTYPE
    full = np.frombuffer(got[0][1], RESULT_DTYPE)
    kfs, pairs, _ = S.make_map(12, n_features=120, cluster=4, pool=120, n_shared=70, k_candidates=3, cross_cluster=1, seed=2)
    assert len(full) == len(pairs) == got[0][2]
    assert full["best_iteration"].tolist() == list(range(len(pairs)))     # pair order preserved across shards
    for i, (a, b) in enumerate(pairs):
        o = O.estimate_edge([kfs[a]], [kfs[b]], want_debug=False)
        assert full[i]["consensus"] == o["consensus"] and np.array_equal(full[i]["T"].reshape(4, 4), o["T"])
