"""The synthetic keyframe map of the benchmark, numpy mirror of include/uz_synth.h ("64-bit splitmix stream,
implemented once in C++ and mirrored in Python so both see identical bytes").

Same streams, same IEEE operations in the same order as the C header: `make_map` here and `uz_synth_map` there fill
byte-identical arrays (tests/test_synth_splitmix.py compiles the header and compares).  See the header for the scene.
A keyframe is the dict the rest of the package uses: desc uint8[N, desc_bytes], pos float64[N, 3], valid uint8[N],
feature_type, sensor_frame.
"""
import ctypes
import os
import subprocess

import numpy as np

U64 = np.uint64
GAMMA = U64(0x9E3779B97F4A7C15)
FX = FY = 525.0
CX, CY = 319.5, 239.5
COS_HALF_MAX = 0.9914448613738104
T_HALF = 0.43
ORB, BRISK = 2, 3
_TWO53 = 1.1102230246251565e-16


def _mix(z):
    z = np.asarray(z, U64)
    z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
    return z ^ (z >> U64(31))


def _draw(s, j):
    return _mix(U64(s) + (np.asarray(j, U64) + U64(1)) * GAMMA)


def _sub(s, k):
    return U64(_mix(U64(s) ^ _mix(U64(k) + U64(0x632BE59BD9B4E019))))


def _u01(x):
    return (np.asarray(x, U64) >> U64(11)).astype(np.float64) * _TWO53


def _gauss(s, j):
    j = np.asarray(j, U64)
    acc = _u01(_draw(s, U64(12) * j))
    for t in range(1, 12):
        acc = acc + _u01(_draw(s, U64(12) * j + U64(t)))
    return acc - 6.0


def _landmarks(s, j):
    j = np.asarray(j, U64)
    u = 640.0 * _u01(_draw(s, U64(3) * j))
    v = 480.0 * _u01(_draw(s, U64(3) * j + U64(1)))
    z = 0.5 + 6.5 * _u01(_draw(s, U64(3) * j + U64(2)))
    return ((u - CX) * z) / FX, ((v - CY) * z) / FY, z


def _order(s, n):
    return np.argsort(_draw(s, np.arange(n, dtype=U64)), kind="stable")


def _root(seed):
    return U64(_mix(U64(0xC4000000 + seed)))


def _kf_stream(root, i):
    return _sub(root, 0x10000000 + i)


def pose(seed, i):
    """cluster -> keyframe pose of keyframe i, float64[4, 4]"""
    with np.errstate(over="ignore"):
        sp = _sub(_kf_stream(_root(seed), i), 0)
        g = _gauss(sp, np.arange(3))
        g0, g1, g2 = float(g[0]), float(g[1]), float(g[2])
        n = np.sqrt((g0 * g0 + g1 * g1) + g2 * g2)
        a0, a1, a2 = (g0 / n, g1 / n, g2 / n) if n > 1e-12 else (1.0, 0.0, 0.0)
        w = 1.0 - float(_u01(_draw(sp, 100))) * (1.0 - COS_HALF_MAX)
        s2 = float(np.sqrt(1.0 - w * w))
        x, y, z = s2 * a0, s2 * a1, s2 * a2
        T = np.eye(4)
        T[0, :3] = (1.0 - 2.0 * ((y * y) + (z * z)), 2.0 * ((x * y) - (w * z)), 2.0 * ((x * z) + (w * y)))
        T[1, :3] = (2.0 * ((x * y) + (w * z)), 1.0 - 2.0 * ((x * x) + (z * z)), 2.0 * ((y * z) - (w * x)))
        T[2, :3] = (2.0 * ((x * z) - (w * y)), 2.0 * ((y * z) + (w * x)), 1.0 - 2.0 * ((x * x) + (y * y)))
        for k in range(3):
            T[k, 3] = T_HALF * (2.0 * float(_u01(_draw(sp, 101 + k))) - 1.0)
    return T


def _pool(root, cluster_index, pool, words):
    sc = _sub(root, cluster_index)
    jj = np.arange(pool, dtype=U64)
    px, py, pz = _landmarks(_sub(sc, 0), jj)
    pd = _draw(_sub(sc, 1), np.arange(pool * words, dtype=U64)).reshape(pool, words)
    return (px, py, pz), pd


def keyframe(i, n_features=1000, cluster=25, pool=1000, n_shared=600, seed=4, invalid_frac=0.15, desc_bytes=32, _pool_cache=None):
    """One keyframe of the map (== uz_synth_keyframe)."""
    N, P, S, W = n_features, pool, n_shared, desc_bytes // 8
    if N <= 0 or P <= 0 or not (0 <= S <= min(P, N)) or desc_bytes not in (32, 64) or cluster <= 0:
        raise ValueError("bad synthetic configuration")
    with np.errstate(over="ignore"):
        root = _root(seed)
        ci = i // cluster
        if _pool_cache is not None and _pool_cache.get("ci") == ci:
            (px, py, pz), pd = _pool_cache["v"]
        else:
            (px, py, pz), pd = _pool(root, ci, P, W)
            if _pool_cache is not None:
                _pool_cache["ci"], _pool_cache["v"] = ci, ((px, py, pz), pd)
        sk = _kf_stream(root, i)
        T = pose(seed, i)
        sel = _order(_sub(sk, 1), P)[:S]
        Lx = np.empty(N); Ly = np.empty(N); Lz = np.empty(N)
        D = np.empty((N, W), U64)
        x, y, z = px[sel], py[sel], pz[sel]
        Lx[:S] = ((T[0, 0] * x + T[0, 1] * y) + T[0, 2] * z) + T[0, 3]
        Ly[:S] = ((T[1, 0] * x + T[1, 1] * y) + T[1, 2] * z) + T[1, 3]
        Lz[:S] = ((T[2, 0] * x + T[2, 1] * y) + T[2, 2] * z) + T[2, 3]
        at = np.arange(S * W, dtype=U64) * U64(4)
        s_flip = _sub(sk, 3)
        m = _draw(s_flip, at) & _draw(s_flip, at + U64(1)) & _draw(s_flip, at + U64(2)) & _draw(s_flip, at + U64(3))
        D[:S] = pd[sel] ^ m.reshape(S, W)
        fx, fy, fz = _landmarks(_sub(sk, 2), np.arange(N - S, dtype=U64))
        Lx[S:], Ly[S:], Lz[S:] = fx, fy, fz
        D[S:] = _draw(_sub(sk, 4), np.arange((N - S) * W, dtype=U64)).reshape(N - S, W)
        s_noise = _sub(sk, 5)
        r3 = np.arange(N, dtype=U64) * U64(3)
        az = np.abs(Lz)
        zs = np.where(az > 0.3, az, 0.3)
        sz = (0.0012 * zs) * zs
        zn = Lz + _gauss(s_noise, r3) * sz
        scale = zn / np.where(az > 1e-9, Lz, 1.0)
        X0 = Lx * scale + ((_gauss(s_noise, r3 + U64(1)) * 0.5) * zs) / FX
        X1 = Ly * scale + ((_gauss(s_noise, r3 + U64(2)) * 0.5) * zs) / FY
        bad = _u01(_draw(_sub(sk, 6), np.arange(N, dtype=U64))) < invalid_frac
        perm = _order(_sub(sk, 7), N)
    pos = np.stack([X0, X1, zn], 1)
    pos[bad] = (0.0, 0.0, -1.0)
    valid = np.where(bad, 0, 1).astype(np.uint8)
    desc = np.ascontiguousarray(D[perm]).view(np.uint8).reshape(N, desc_bytes)       # little-endian words
    return dict(desc=desc, pos=np.ascontiguousarray(pos[perm]), valid=np.ascontiguousarray(valid[perm]),
                feature_type=ORB if desc_bytes == 32 else BRISK, sensor_frame=0)


def candidates(i, n_keyframes, cluster=25, k_candidates=20, cross_cluster=4, seed=4):
    """candidate partners of keyframe i (== uz_synth_candidates): list of (from, to)"""
    with np.errstate(over="ignore"):
        sk = _kf_stream(_root(seed), i)
        cl = i // cluster
        c0 = cl * cluster
        nk = min(cluster, n_keyframes - c0)
        order = _order(_sub(sk, 8), nk)
        want_own = min(k_candidates - cross_cluster, nk - 1)
        out = []
        for j in order:
            if len(out) >= want_own:
                break
            if c0 + int(j) != i:
                out.append((i, c0 + int(j)))
        if n_keyframes > cluster:
            sx = _sub(sk, 9)
            t = 0
            while len(out) < k_candidates:
                js = _draw(sx, np.arange(t, t + 64, dtype=U64)) % U64(n_keyframes)
                for j in js:
                    if len(out) >= k_candidates:
                        break
                    if int(j) // cluster != cl:
                        out.append((i, int(j)))
                t += 64
    return out


_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libuz_synth.so")


def build_native():
    """Compile include/uz_synth.h into libuz_synth.so (gcc, no FMA contraction): the fast path of make_map."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "uzliti_slam_b200", "csrc", "uz_synth_lib.c")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-I", os.path.join(root, "include"), "-o", _LIB, src, "-lm"])


class _Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_keyframes", "n_features", "cluster", "pool", "n_shared", "k_candidates",
                                              "cross_cluster", "desc_bytes")] + [("invalid_frac", ctypes.c_double), ("seed", ctypes.c_uint64)]


def _native_map(cfg, out, threads):
    lib = ctypes.CDLL(_LIB)
    lib.uz_synth_map_mt.restype = ctypes.c_longlong
    lib.uz_synth_map_mt.argtypes = [ctypes.POINTER(_Cfg)] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
    n, N = cfg.n_keyframes, cfg.n_features
    desc, pos, valid = out
    assert desc.dtype == np.uint8 and pos.dtype == np.float64 and valid.dtype == np.uint8
    assert desc.flags.c_contiguous and pos.flags.c_contiguous and valid.flags.c_contiguous
    assert desc.shape == (n, N, cfg.desc_bytes) and pos.shape == (n, N, 3) and valid.shape == (n, N)
    pairs = np.empty((n * max(cfg.k_candidates, 1), 2), np.int32)
    k = lib.uz_synth_map_mt(ctypes.byref(cfg), desc.ctypes.data, pos.ctypes.data, valid.ctypes.data, pairs.ctypes.data, threads)
    if k < 0:
        raise ValueError("bad synthetic configuration")
    return pairs[:k].copy()


def make_map(n_keyframes, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=20, cross_cluster=4, seed=4,
             invalid_frac=0.15, out=None, desc_bytes=32, native=None, threads=8):
    """Same signature and return value as synthetic.make_map: (keyframes, pairs int32[n, 2], poses float64[n, 4, 4]).
    `out` = (desc, pos, valid) big arrays that receive the keyframes as slices (e.g. pinned host memory).
    native = None: use libuz_synth.so (the C header, same bytes) when it has been built, else numpy; True / False force one."""
    if native is None:
        native = os.path.exists(_LIB)
    if native:
        cfg = _Cfg(n_keyframes, n_features, cluster, pool, n_shared, k_candidates, cross_cluster, desc_bytes, invalid_frac, seed)
        if out is None:
            out = (np.empty((n_keyframes, n_features, desc_bytes), np.uint8), np.empty((n_keyframes, n_features, 3), np.float64),
                   np.empty((n_keyframes, n_features), np.uint8))
        pairs = _native_map(cfg, out, threads)
        ftype = ORB if desc_bytes == 32 else BRISK
        kfs = [dict(desc=out[0][i], pos=out[1][i], valid=out[2][i], feature_type=ftype, sensor_frame=0) for i in range(n_keyframes)]
        lib = ctypes.CDLL(_LIB)
        lib.uz_synth_pose_c.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_int32, ctypes.c_void_p]
        poses = np.zeros((n_keyframes, 4, 4))
        poses[:, 3, 3] = 1.0
        row = np.empty(12)
        for i in range(n_keyframes):
            lib.uz_synth_pose_c(ctypes.byref(cfg), i, row.ctypes.data)
            poses[i, :3, :] = row.reshape(3, 4)
        return kfs, pairs, poses
    kfs, pairs, poses = [], [], []
    cache = {}
    for i in range(n_keyframes):
        kf = keyframe(i, n_features, cluster, pool, n_shared, seed, invalid_frac, desc_bytes, _pool_cache=cache)
        if out is not None:
            out[0][i] = kf["desc"]; out[1][i] = kf["pos"]; out[2][i] = kf["valid"]
            kf = dict(kf, desc=out[0][i], pos=out[1][i], valid=out[2][i])
        kfs.append(kf)
        poses.append(pose(seed, i))
        pairs += candidates(i, n_keyframes, cluster, k_candidates, cross_cluster, seed)
    return kfs, np.array(pairs, np.int32).reshape(-1, 2), np.array(poses)


def checksum(a):
    """== uz_synth_checksum over the bytes of `a`"""
    b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    pad = (-len(b)) % 8
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view(U64)
    with np.errstate(over="ignore"):
        return int(np.sum(_mix(w + (np.arange(len(w), dtype=U64) + U64(1)) * GAMMA), dtype=U64))
