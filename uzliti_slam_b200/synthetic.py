"""Synthetic Kinect-shaped keyframes with planted correspondences.

Camera 640x480, fx=fy=525, cx=319.5, cy=239.5 (the constants of the reference's legacy estimator,
graph_slam_common/src/transformation/feature_transformation_estimator.cpp:37); depth uniform in
[0.5, 7] m (feature_max_depth, iti_slam_launch/yaml/slam.yaml:7); missing depth is z=-1, x=y=0,
valid=False (feature_extraction/src/feature_extraction_core.cpp:286-289); descriptors are 32-byte
ORB-256 rows (feature_extraction/external/aorb/aorb.h:54) or, with desc_bytes=64, BRISK/FREAK-512 rows
(feature_extraction_core.cpp:69-77).

A keyframe is a dict: desc uint8[N,desc_bytes], pos float64[N,3] (memory layout == Eigen 3xN column-major),
valid uint8[N], feature_type (2 = ORB), sensor_frame (interned tag).
"""
import numpy as np

FX = FY = 525.0
CX, CY = 319.5, 239.5
W, H = 640, 480
ORB = 2
BRISK = 3


def _rand_rotation(rng, max_angle_rad):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(0.0, max_angle_rad)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _rand_pose(rng, max_angle_deg, max_trans):
    T = np.eye(4)
    T[:3, :3] = _rand_rotation(rng, np.deg2rad(max_angle_deg))
    v = rng.normal(size=3)
    v *= max_trans * rng.uniform() ** (1 / 3) / np.linalg.norm(v)
    T[:3, 3] = v
    return T


def _landmarks(rng, n):
    u = rng.uniform(0, W, n)
    v = rng.uniform(0, H, n)
    z = rng.uniform(0.5, 7.0, n)
    return np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1)


def _observe(rng, X, noise=True):
    """Kinect-like noise: sigma_z = 0.0012 z^2, lateral noise through the pinhole + 0.5 px."""
    if not noise:
        return X.copy()
    z = X[:, 2]
    zs = np.maximum(np.abs(z), 0.3)
    sz = 0.0012 * zs * zs
    zn = z + rng.normal(size=len(z)) * sz
    scale = zn / np.where(np.abs(z) > 1e-9, z, 1.0)
    x = X[:, 0] * scale + rng.normal(size=len(z)) * 0.5 * zs / FX
    y = X[:, 1] * scale + rng.normal(size=len(z)) * 0.5 * zs / FY
    return np.stack([x, y, zn], 1)


def _flip(rng, desc, k=4):
    """flip each bit with p = 2^-k (k=4: 0.0625 ~ SURVEY's 0.06)."""
    m = rng.integers(0, 256, desc.shape, dtype=np.uint8)
    for _ in range(k - 1):
        m &= rng.integers(0, 256, desc.shape, dtype=np.uint8)
    return desc ^ m


def _finish(rng, desc, pos, invalid_frac, tie_stress, sensor_frame, shuffle=True, feature_type=ORB):
    n = len(desc)
    valid = np.ones(n, np.uint8)
    if invalid_frac > 0 and n > 0:
        bad = rng.random(n) < invalid_frac
        valid[bad] = 0
        pos = pos.copy()
        pos[bad] = (0.0, 0.0, -1.0)
    if tie_stress:
        desc = desc.copy()
        desc[:, 4:] = 0
    perm = rng.permutation(n) if shuffle else np.arange(n)
    return dict(desc=np.ascontiguousarray(desc[perm]), pos=np.ascontiguousarray(pos[perm]),
                valid=np.ascontiguousarray(valid[perm]), feature_type=feature_type, sensor_frame=sensor_frame), perm


def make_pair(n_from, n_to=None, seed=0, rho=0.5, invalid_frac=0.15, gross_outlier_frac=0.10,
              tie_stress=False, noise=True, max_angle_deg=30.0, max_trans=1.5, sensor_frame=0, desc_bytes=32):
    """One (from, to) keyframe pair.  Returns (kf_from, kf_to, T_gt) with T_gt * p_to = x_from,
    i.e. the transform the reference stores in edge.transform_ (estimateSVD(Pd /*to*/, Xd /*from*/))."""
    n_to = n_from if n_to is None else n_to
    rng = np.random.default_rng(0x5EED0000 + seed)
    ns = int(round(rho * min(n_from, n_to)))
    T_gt = _rand_pose(rng, max_angle_deg, max_trans)
    Tinv = np.linalg.inv(T_gt)
    Xs = _landmarks(rng, ns)                                   # shared, in from-frame
    ftype = ORB if desc_bytes == 32 else BRISK
    Ds = rng.integers(0, 256, (ns, desc_bytes), dtype=np.uint8)
    Xf = np.concatenate([Xs, _landmarks(rng, n_from - ns)])
    Df = np.concatenate([_flip(rng, Ds), rng.integers(0, 256, (n_from - ns, desc_bytes), dtype=np.uint8)])
    Ps = Xs @ Tinv[:3, :3].T + Tinv[:3, 3]
    if gross_outlier_frac > 0 and ns > 0:
        g = rng.random(ns) < gross_outlier_frac
        Ps[g] = _landmarks(rng, int(g.sum()))
    Pt = np.concatenate([Ps, _landmarks(rng, n_to - ns)])
    Dt = np.concatenate([_flip(rng, Ds), rng.integers(0, 256, (n_to - ns, desc_bytes), dtype=np.uint8)])
    kf_from, _ = _finish(rng, Df, _observe(rng, Xf, noise), invalid_frac, tie_stress, sensor_frame, feature_type=ftype)
    kf_to, _ = _finish(rng, Dt, _observe(rng, Pt, noise), invalid_frac, tie_stress, sensor_frame, feature_type=ftype)
    return kf_from, kf_to, T_gt


def make_map(n_keyframes, n_features=1000, cluster=25, pool=1000, n_shared=600, k_candidates=20,
             cross_cluster=4, seed=0, invalid_frac=0.15, out=None, desc_bytes=32):
    """A keyframe map with loop-closure candidates (configs C3/C4 of BASELINE.json).

    Keyframes come in clusters that share a pool of landmarks: each keyframe observes n_shared pool
    landmarks (own pose within 15 deg / 0.75 m of the cluster frame, so relative motion stays within
    the reference's 30 deg / 1.5 m gates) plus fresh ones.  Every keyframe gets k_candidates candidate
    partners: k_candidates - cross_cluster from its own cluster (true loop closures) and
    cross_cluster from random other clusters (place-recognition false positives).

    `out` = (desc uint8[n,N,desc_bytes], pos float64[n,N,3], valid uint8[n,N]) optionally receives the keyframes
    as slices of three big arrays (e.g. pinned host memory), so the whole map is contiguous per field.

    Returns (keyframes list, pairs int32[n_pairs,2] (from,to), poses float64[n,4,4] (cluster->keyframe)).
    """
    rng = np.random.default_rng(0xC4000000 + seed)
    kfs, poses = [], []
    cluster_of = []
    for c0 in range(0, n_keyframes, cluster):
        nk = min(cluster, n_keyframes - c0)
        Xp = _landmarks(rng, pool)
        Dp = rng.integers(0, 256, (pool, desc_bytes), dtype=np.uint8)
        for _ in range(nk):
            G = _rand_pose(rng, 15.0, 0.75)
            sel = rng.permutation(pool)[:n_shared]
            Xs = Xp[sel] @ G[:3, :3].T + G[:3, 3]
            X = np.concatenate([Xs, _landmarks(rng, n_features - n_shared)])
            D = np.concatenate([_flip(rng, Dp[sel]),
                                rng.integers(0, 256, (n_features - n_shared, desc_bytes), dtype=np.uint8)])
            kf, _ = _finish(rng, D, _observe(rng, X), invalid_frac, False, 0, feature_type=ORB if desc_bytes == 32 else BRISK)
            if out is not None:
                k = len(kfs)
                out[0][k] = kf["desc"]; out[1][k] = kf["pos"]; out[2][k] = kf["valid"]
                kf = dict(kf, desc=out[0][k], pos=out[1][k], valid=out[2][k])
            kfs.append(kf)
            poses.append(G)
            cluster_of.append(c0 // cluster)
    cluster_of = np.array(cluster_of)
    pairs = []
    for i in range(n_keyframes):
        c = cluster_of[i]
        own = np.flatnonzero(cluster_of == c)
        own = own[own != i]
        k_own = min(len(own), k_candidates - cross_cluster)
        cand = list(rng.choice(own, k_own, replace=False)) if k_own > 0 else []
        while len(cand) < k_candidates and n_keyframes > cluster:
            j = int(rng.integers(0, n_keyframes))
            if cluster_of[j] != c:
                cand.append(j)
        pairs += [(i, int(j)) for j in cand]
    return kfs, np.array(pairs, np.int32).reshape(-1, 2), np.array(poses)


def gt_transform(poses, i_from, i_to):
    """T with T * p_to = x_from for two keyframes of the same cluster."""
    return poses[i_from] @ np.linalg.inv(poses[i_to])


def rot_angle(Ra, Rb):
    """Well-conditioned rotation distance: atan2(|skew part|, (trace-1)/2)."""
    R = Ra.T @ Rb
    s = 0.5 * np.sqrt((R[2, 1] - R[1, 2]) ** 2 + (R[0, 2] - R[2, 0]) ** 2 + (R[1, 0] - R[0, 1]) ** 2)
    c = 0.5 * (np.trace(R) - 1.0)
    return float(np.arctan2(s, c))
