#!/usr/bin/env python
"""How much do the outcomes of the path depend on the exact arithmetic of the pose solve?  

Stage 3 of the oracle restates PCL's TransformationFromCorrespondences + Eigen's JacobiSVD<Matrix3f> from the published
algorithms; the binaries the reference links cannot be run here ("parity unpinned").  This script measures the EXPOSURE: every
hypothesis of every pair is solved a second time with an independent solver - float64 Kabsch (numpy SVD), result rounded to
float32 like the reference's `.cast<double>()` of a float solve - and the sequential RANSAC semantics are replayed on its
counts.  Reported: how often a hypothesis' inlier count, the winning hypothesis, the final inlier set, the final consensus
change, and how far the final transforms are apart.  CPU only (oracle + numpy).
usage: python scripts/stage3_sensitivity.py [pairs per configuration]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def kabsch_f32(P, Q):
    """least-squares rigid T with T p ~ q, solved in float64, rounded to float32 (as a float32 solver's result would be)"""
    cp, cq = P.mean(0), Q.mean(0)
    H = (P - cp).T @ (Q - cq)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T)) or 1.0
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = cq - R @ cp
    return T.astype(np.float32).astype(np.float64)


def inliers(P, Q, T, thr):
    return np.linalg.norm(P @ T[:3, :3].T + T[:3, 3] - Q, axis=1) < thr


def replay(P, Q, samples, thr, bp):
    """prosac() of the reference (feature_transformation_estimator.cpp:214-297) with the alternative solver"""
    M = len(P)
    best, best_T, best_i, counts = 0, np.eye(4), -1, []
    for i, s in enumerate(samples):
        T = kabsch_f32(P[s], Q[s])
        c = int(inliers(P, Q, T, thr).sum())
        counts.append(c)
        if c > best:
            best, best_T, best_i = c, T, i
            if best >= 3 and best > bp * M:
                break
    if best < 3:
        return dict(counts=counts, winner=best_i, mask=np.zeros(M, bool), consensus=0, T=np.eye(4))
    m = inliers(P, Q, best_T, thr)
    T = kabsch_f32(P[m], Q[m])
    m2 = inliers(P, Q, T, thr)
    return dict(counts=counts, winner=best_i, mask=m2, consensus=int(m2.sum()), T=T)


def rot_angle(Ra, Rb):
    R = Ra.T @ Rb
    s = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.arctan2(s, (np.trace(R) - 1) / 2))


def measure(n_pairs=40, iterations=100, thr=0.1, bp=0.6):
    from oracle import binding as O
    from uzliti_slam_b200 import synthetic as S
    configs = {"C1 (500 features)": dict(n_from=500), "C2-C4 (1000 features)": dict(n_from=1000),
               "C5 camera (2000 features)": dict(n_from=2000), "tie-heavy descriptors (300)": dict(n_from=300, tie_stress=True)}
    out = {}
    for name, kw in configs.items():
        st = dict(pairs=0, hyp=0, hyp_count_differs=0, winner_differs=0, mask_differs=0, consensus_differs=0,
                  mask_bits_differing=0, dt_max=0.0, dr_max=0.0, dt_when_same_set=0.0, dr_when_same_set=0.0)
        for p in range(n_pairs):
            f, t, _ = S.make_pair(seed=7000 + p, **kw)
            o = O.estimate_edge([f], [t], thr=thr, iterations=iterations, bp=bp)
            if not o["ok"] or o["n_matches"] < 3:
                continue
            m = o["matches"]
            P, Q = t["pos"][m[:, 0]], f["pos"][m[:, 1]]
            samples = O.sample_list(len(P), iterations, True)
            a = replay(P, Q, samples, thr, bp)
            n = min(len(a["counts"]), int(o["iterations_run"]))
            oc = o["counts"][:n]
            st["pairs"] += 1
            st["hyp"] += n
            st["hyp_count_differs"] += int((np.asarray(a["counts"][:n]) != oc).sum())
            st["winner_differs"] += int(a["winner"] != o["best_iteration"])
            diff = int((a["mask"] != o["inlier_mask"]).sum())
            st["mask_differs"] += int(diff > 0)
            st["mask_bits_differing"] += diff
            st["consensus_differs"] += int(a["consensus"] != o["consensus"])
            dt = float(np.linalg.norm(a["T"][:3, 3] - o["T"][:3, 3]))
            dr = rot_angle(a["T"][:3, :3], o["T"][:3, :3])
            st["dt_max"] = max(st["dt_max"], dt); st["dr_max"] = max(st["dr_max"], dr)
            if diff == 0:
                st["dt_when_same_set"] = max(st["dt_when_same_set"], dt); st["dr_when_same_set"] = max(st["dr_when_same_set"], dr)
        out[name] = st
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    res = measure(n)
    print(json.dumps(res, indent=1))
    tot = {k: sum(v[k] for v in res.values()) for k in ("pairs", "hyp", "hyp_count_differs", "winner_differs", "mask_differs", "consensus_differs")}
    print("ALL:", json.dumps(tot))
