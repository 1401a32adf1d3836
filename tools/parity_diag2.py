"""GPU-box diagnostic, step 2: for the pairs dumped by parity_diag.py, replay the minimal solver and the exact scorer on the
device for the first 640 iterations' Philox samples and save everything for comparison with the host build of the same code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spherical_sfm_b200 as S  # noqa: E402
from conftest import THR2, E_of  # noqa: E402

g = np.load(os.path.join(ROOT, "tools", "parity_pairs.npz"))
eng = S.Engine(0)
out = {}
for k in range(int(g["num"])):
    rays = g["rays_%d" % k]
    p, flsq = int(g["pid_%d" % k]), int(g["flsq_%d" % k])
    n = len(rays)
    samples = np.array([S.sample(0, p, i, 3, n) for i in range(640)], np.int32)
    models, nm = eng.minimal_solve(rays, samples, 0)
    E9 = np.array([E_of(m).ravel() for m in models.reshape(-1, 6)])
    E9 = np.nan_to_num(E9, nan=0.0)
    se, ce = eng.score_exact(E9, rays, THR2)
    opt = S.default_options(squared_inlier_threshold=THR2, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=flsq, first_pair_id=p)
    res, flags = eng.estimate_pairs(rays, np.array([0, n], np.int64), opt)
    out["models_%d" % k] = models
    out["score_%d" % k] = se
    out["count_%d" % k] = ce
    out["res_%d" % k] = res
    out["flags_%d" % k] = flags
    print(k, p, res["num_iterations"], res["best_num_inliers"], res["number_lo_iterations"], res["best_model_score"])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "parity_diag2.npz"), **out)
