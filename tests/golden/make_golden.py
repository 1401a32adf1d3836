"""Generates tests/golden/lomsac_golden.npz with oracle/_ref: the reference's own RansacLib driver
loops (include/RansacLib/ransac.h, evaluation/vanilla_ransac.h, compiled from /root/reference) around
the restated spherical estimator.  Run here (needs /root/reference); the .npz travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import spherical_sfm_b200 as S  # noqa: E402

THR2 = (2.0 / 600.0) ** 2
ref = O.load_ref()
assert ref is not None and ref.is_reference
cases = [
    (dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1000, 500, False, 11),  # config C1
    (dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1500, 1050, False, 12),  # config C3 pair
    (dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1, inward=1), 400, 120, True, 13),
    (dict(), 500, 250, False, 14),  # RansacLib default LO options
    (dict(driver=1), 500, 200, False, 15),  # VanillaMSAC (evaluation/test_ransac.cpp)
    (dict(solver_kind=1, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 300, 90, False, 16),
    (dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 40, 36, False, 17),  # hopeless: runs to max
]
out = {"num_cases": len(cases)}
for k, (kw, n, nout, inward, pid) in enumerate(cases):
    pr = S.problems.make_problem(S.problems.make_rng(2024, pid), n, inward, None, 1 / 600, nout, 20.0)
    opt = O.default_options(squared_inlier_threshold=THR2, **kw)
    res, inl = ref.estimate_pair(pr.rays, opt, pid)
    names = ["squared_inlier_threshold"] + list(kw.keys())
    out["opt_names_%d" % k] = np.array(names)
    out["opt_vals_%d" % k] = np.array([getattr(opt, nme) for nme in names], np.float64)
    out["rays_%d" % k] = pr.rays
    out["pair_id_%d" % k] = pid
    out["E_%d" % k] = np.array(res.E)
    out["r_%d" % k] = np.array(res.r)
    out["num_iterations_%d" % k] = res.num_iterations
    out["best_num_inliers_%d" % k] = res.best_num_inliers
    out["number_lo_iterations_%d" % k] = res.number_lo_iterations
    out["best_model_score_%d" % k] = res.best_model_score
    out["inliers_%d" % k] = inl
    out["R_gt_%d" % k] = pr.R
    print(k, kw, res.num_iterations, res.best_num_inliers, res.number_lo_iterations, res.status)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lomsac_golden.npz"), **out)
