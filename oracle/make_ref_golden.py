#!/usr/bin/env python
"""Golden vectors for stage 3 from the REFERENCE's own code (oracle/_ref/ref_stage3, built by oracle/build_ref.sh where Eigen,
PCL and Boost headers exist) -> tests/golden/ref_stage3.npz.  The cases are generated here (seeded), so the file can be
regenerated; tests/test_golden.py compares the oracle with it bit for bit."""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def make_cases():
    """(kind, P[M,3], Q[M,3], max_error, iterations, break_percentage, do_prosac, T_in[4,4])"""
    from uzliti_slam_b200 import synthetic as S
    from oracle import binding as O
    rng = np.random.default_rng(20260)
    cases = []
    eye = np.eye(4)
    # (a) the matched point sets of real synthetic pairs (what estimateEdgeDirect hands to estimateSVD, :118-130)
    for seed, n in ((1, 500), (2, 1000), (3, 300), (4, 64), (5, 2000)):
        f, t, _ = S.make_pair(n, seed=seed)
        o = O.estimate_edge([f], [t])
        m = o["matches"]
        P, Q = t["pos"][m[:, 0]], f["pos"][m[:, 1]]
        cases.append((0, P, Q, 0.1, 100, 0.6, 1, eye))
        cases.append((0, P, Q, 0.3, 200, 1.0, 0, eye))                # the TransformationFilter call (transformation_filter.cpp:272)
        cases.append((0, P, Q, 0.02, 1000, 0.6, 1, eye))
    # (b) 3-point and small solves incl. degenerate samples (collinear, duplicate points, far from the origin, tiny)
    for k in range(40):
        m = int(rng.integers(3, 12))
        P = rng.normal(size=(m, 3)) * rng.choice([1e-3, 1.0, 30.0])
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(R) < 0:
            R[:, 0] = -R[:, 0]
        Q = P @ R.T + rng.normal(size=3) + rng.normal(size=(m, 3)) * 1e-3
        if k % 5 == 0:
            P[1] = P[0]
        if k % 7 == 0:
            P = np.outer(np.linspace(0, 1, m), rng.normal(size=3))          # collinear
        cases.append((1, P, Q, 0.0, 1, 0.0, 0, eye))
    # (c) consensus3D on thresholds that sit ON residuals (strict '<' at the boundary)
    for k in range(10):
        m = 200
        P = rng.uniform(-3, 3, size=(m, 3))
        T = np.eye(4)
        T[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        T[:3, 3] = rng.normal(size=3)
        Q = P @ T[:3, :3].T + T[:3, 3] + rng.normal(size=(m, 3)) * 0.05
        r = np.linalg.norm(P @ T[:3, :3].T + T[:3, 3] - Q, axis=1)
        cases.append((2, P, Q, float(np.sort(r)[m // 2]), 1, 0.0, 0, T))
    # (d) fewer than three points, and nothing but outliers
    cases.append((0, np.zeros((2, 3)), np.ones((2, 3)), 0.1, 100, 0.6, 1, eye))
    cases.append((0, rng.normal(size=(50, 3)), rng.normal(size=(50, 3)) * 10, 0.01, 100, 0.6, 1, eye))
    return cases


def main():
    exe = os.path.join(HERE, "_ref", "ref_stage3")
    if not os.path.exists(exe):
        raise SystemExit("oracle/_ref/ref_stage3 is missing: run oracle/build_ref.sh first (needs Eigen 3, PCL and Boost headers)")
    cases = make_cases()
    blob = struct.pack("<i", len(cases))
    for kind, P, Q, thr, it, bp, prosac, T in cases:
        P = np.ascontiguousarray(P, np.float64); Q = np.ascontiguousarray(Q, np.float64)
        blob += struct.pack("<iidid", kind, len(P), thr, it, bp) + struct.pack("<i", prosac)
        blob += np.ascontiguousarray(T, np.float64).tobytes() + P.tobytes() + Q.tobytes()
    tmp = os.path.join(HERE, "_ref")
    open(os.path.join(tmp, "cases.bin"), "wb").write(blob)
    subprocess.check_call([exe, os.path.join(tmp, "cases.bin"), os.path.join(tmp, "out.bin")])
    raw = open(os.path.join(tmp, "out.bin"), "rb").read()
    out, at = {}, 0
    for i, (kind, P, Q, thr, it, bp, prosac, T) in enumerate(cases):
        m = len(P)
        Tout = np.frombuffer(raw, np.float64, 16, at).reshape(4, 4); at += 128
        cons = struct.unpack_from("<i", raw, at)[0]; at += 4
        mse = struct.unpack_from("<d", raw, at)[0]; at += 8
        mask = np.frombuffer(raw, np.uint8, m, at); at += m
        out.update({f"c{i}_kind": kind, f"c{i}_P": P, f"c{i}_Q": Q, f"c{i}_par": np.array([thr, it, bp, prosac], np.float64),
                    f"c{i}_Tin": T, f"c{i}_T": Tout.copy(), f"c{i}_consensus": cons, f"c{i}_mse": mse, f"c{i}_mask": mask.copy()})
    assert at == len(raw)
    out["n_cases"] = len(cases)
    dst = os.path.join(ROOT, "tests", "golden", "ref_stage3.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {len(cases)} cases from the reference's own code")


if __name__ == "__main__":
    main()
