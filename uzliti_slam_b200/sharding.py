"""Pair-list sharding across the GPUs of one box and the gather of per-pair edge records.

Keyframe pairs are independent (the reference runs one estimateEdgeImpl per pair,
transformation_estimation/src/transformation_estimator.cpp:45-62), so the path shards with no data-path
collective: the keyframe store is replicated, rank r takes a contiguous chunk of the from-sorted pair list
(neighbouring pairs share their `from` keyframe, which keeps its descriptor tile in L2), and only the
fixed-size edge records are exchanged, with one all-gather over NCCL (gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_pairs, world_size, rank):
    """[lo, hi) of rank's contiguous chunk; chunks differ by at most one pair and cover [0, n_pairs)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_pairs), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs(pairs, world_size, rank):
    lo, hi = shard_bounds(len(pairs), world_size, rank)
    return pairs[lo:hi], lo


def gather_records(local_records, n_pairs, group=None):
    """All-gather fixed-size result records (numpy structured array) of every rank's shard into the full,
    pair-ordered array.  Works on any torch.distributed backend; tensors live where the backend wants them
    (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    itemsize = local_records.dtype.itemsize
    sizes = [shard_bounds(n_pairs, world, r) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros(max_n * itemsize, dtype=torch.uint8)
    raw = np.frombuffer(local_records.tobytes(), dtype=np.uint8)
    buf[:raw.size] = torch.from_numpy(raw.copy())
    buf = buf.to(dev)
    out = torch.empty(world * max_n * itemsize, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    flat = out.cpu().numpy()
    full = np.empty(n_pairs, dtype=local_records.dtype)
    for r, (lo, hi) in enumerate(sizes):
        chunk = flat[r * max_n * itemsize: r * max_n * itemsize + (hi - lo) * itemsize]
        full[lo:hi] = np.frombuffer(chunk.tobytes(), dtype=local_records.dtype)
    assert sizes[rank][1] - sizes[rank][0] == len(local_records)
    return full
