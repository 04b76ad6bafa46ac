"""Generates tests/golden/*.npz — run once in the build container (python tests/golden/make_golden.py).

There are no golden vectors in the reference (it has no tests at all), so the fixtures pin
  * the matching stage to the live OpenCV matcher the reference calls (cv2.BFMatcher(NORM_HAMMING).knnMatch)
  * the sample-index list to the real libstdc++ std::random_shuffle over glibc rand() (via the oracle, which
    calls the real functions)
  * the full edge estimate to the oracle's outputs at generation time (regression anchor for oracle AND GPU).
Inputs are stored alongside the outputs so the GPU box needs neither cv2 nor /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from oracle import binding as O  # noqa: E402
from uzliti_slam_b200 import synthetic as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cv2_knn(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx = np.full((len(q), 2), -1, np.int32)
    dist = np.full((len(q), 2), -1, np.int32)
    for i, row in enumerate(m):
        for j, d in enumerate(row):
            idx[i, j] = d.trainIdx
            dist[i, j] = int(d.distance)
    return idx, dist


def main():
    out = {}
    cases = [("random300", dict(n_from=300, seed=501)), ("ties256", dict(n_from=256, seed=502, tie_stress=True)),
             ("ragged", dict(n_from=190, n_to=333, seed=503)), ("highinlier", dict(n_from=200, seed=504, rho=0.9)),
             ("nodepth", dict(n_from=150, seed=505, invalid_frac=0.95)),
             # BRISK / FREAK rows (64 bytes, feature_extraction_core.cpp:69-77)
             ("wide300", dict(n_from=300, seed=506, desc_bytes=64)),
             ("wide_ties", dict(n_from=200, n_to=257, seed=507, desc_bytes=64, tie_stress=True))]
    for name, kw in cases:
        f, t, Tgt = S.make_pair(**kw)
        ci, cd = cv2_knn(t["desc"], f["desc"])
        o = O.estimate_edge([f], [t])
        out[name] = dict(f_desc=f["desc"], f_pos=f["pos"], f_valid=f["valid"], t_desc=t["desc"], t_pos=t["pos"],
                         t_valid=t["valid"], cv2_idx=ci, cv2_dist=cd, ok=np.int32(o["ok"]),
                         n_ratio=np.int32(o["n_ratio_matches"]), n_matches=np.int32(o["n_matches"]),
                         consensus=np.int32(o["consensus"]), best_iteration=np.int32(o["best_iteration"]),
                         iterations_run=np.int32(o["iterations_run"]), mse=np.float64(o["mse"]),
                         info_scale=np.float64(o["info_scale"]), T=o["T"], matches=o["matches"],
                         inlier_mask=o["inlier_mask"].astype(np.uint8), counts=o["counts"], T_gt=Tgt)
    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, f"edge_{name}.npz"), **d)
    samples = {f"M{M}_I{I}_p{p}": O.sample_list(M, I, bool(p))
               for (M, I, p) in [(3, 100, 1), (57, 100, 1), (302, 100, 1), (557, 100, 1), (100, 200, 0), (300, 1000, 1)]}
    samples["rand_seed1_first64"] = O.glibc_rand(64)
    np.savez_compressed(os.path.join(HERE, "sample_lists.npz"), **samples)
    print("wrote", sorted(os.listdir(HERE)), "cv2", cv2.__version__)


if __name__ == "__main__":
    main()
