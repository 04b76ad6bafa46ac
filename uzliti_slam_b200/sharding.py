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


def broadcast_keyframes(keyframes, src=0, group=None):
    """Store maintenance for a replicated store: the rank that ingests keyframes (in the reference ONE
    process receives the sensor messages, graph_slam/src/graph_slam_node.cpp:207-301) broadcasts them packed — counts,
    descriptor rows, positions, valid bytes as four flat tensors — and every rank rebuilds the same keyframe list to feed
    its own `EdgeEstimator.add_keyframes`.  `keyframes` (a list of single-camera dicts) is only read on `src`."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    if rank == src:
        # per keyframe: rows, feature_type, sensor_frame, descriptor width (32: ORB / BRIEF, 64: BRISK / FREAK)
        meta = np.array([[len(k["desc"]), int(k.get("feature_type", 2)), int(k.get("sensor_frame", 0)),
                          int(np.asarray(k["desc"]).shape[1]) if np.asarray(k["desc"]).ndim == 2 else 32] for k in keyframes],
                        np.int64).reshape(-1, 4)
        head = torch.tensor([len(keyframes)], dtype=torch.int64)
    else:
        meta, head = None, torch.zeros(1, dtype=torch.int64)
    head = head.to(dev)
    dist.broadcast(head, src, group=group)
    n_kf = int(head.item())
    t_meta = (torch.from_numpy(meta) if rank == src else torch.zeros((n_kf, 4), dtype=torch.int64)).to(dev)
    dist.broadcast(t_meta, src, group=group)
    meta = t_meta.cpu().numpy()
    total = int(meta[:, 0].sum())
    desc_total = int((meta[:, 0] * meta[:, 3]).sum())
    if rank == src:
        # descriptors travel as ONE flat byte buffer (rows of different widths cannot share a 2-D tensor)
        desc = np.concatenate([np.ascontiguousarray(k["desc"], np.uint8).reshape(-1) for k in keyframes] + [np.zeros(0, np.uint8)])
        pos = np.concatenate([np.ascontiguousarray(k["pos"], np.float64).reshape(-1, 3) for k in keyframes] + [np.zeros((0, 3))])
        valid = np.concatenate([np.ascontiguousarray(k["valid"], np.uint8).reshape(-1) for k in keyframes] + [np.zeros(0, np.uint8)])
        assert desc.size == desc_total
        tens = [torch.from_numpy(desc), torch.from_numpy(pos), torch.from_numpy(valid)]
    else:
        tens = [torch.zeros(desc_total, dtype=torch.uint8), torch.zeros((total, 3), dtype=torch.float64),
                torch.zeros(total, dtype=torch.uint8)]
    out = []
    for t in tens:
        t = t.to(dev)
        dist.broadcast(t, src, group=group)
        out.append(t.cpu().numpy())
    desc, pos, valid = out
    kfs, o, od = [], 0, 0
    for n, ftype, frame, width in meta:
        n, width = int(n), int(width)
        kfs.append(dict(desc=desc[od:od + n * width].reshape(n, width), pos=pos[o:o + n], valid=valid[o:o + n],
                        feature_type=int(ftype), sensor_frame=int(frame)))
        o += n
        od += n * width
    return kfs
