"""Generates tests/golden/refgen_golden.npz: problems drawn by the reference's OWN generator
(evaluation/problem_generator/problem_generator.cpp, compiled into oracle/_ref/libssfm_refgen.so: std::default_random_engine,
default seed) and what the reference's own estimator sources (oracle/_ref/libssfm_reffull.so) make of them:
  * evaluation/test_random_problems.cpp: MinimalSolver on the sample {0,1,2}, best-of-solutions Frobenius error
  * evaluation/test_ransac.cpp: VanillaMSAC with thr^2 = (2/focal)^2 on 100 correspondences
  * the pipeline's LO-MSAC options on the same problems
Run here (needs /root/reference); the .npz travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

THR2 = (2.0 / 600.0) ** 2
gen, rf = O.load_ref_gen(), O.load_ref_full()
assert gen is not None and rf is not None
cases = []
for k in range(24):
    inward, noise = bool(k % 2), (0.0 if k < 12 else 1.0 / 600)
    rays, E, R, t = gen.make_random_problem(100, inward, -1.0, noise)
    cases.append((rays, E, R, t, inward, noise))
out = {"num_cases": len(cases)}
for k, (rays, E, R, t, inward, noise) in enumerate(cases):
    out["rays_%d" % k], out["E_%d" % k], out["R_%d" % k], out["t_%d" % k] = rays, E, R, t
    out["cfg_%d" % k] = np.array([int(inward), noise])
    nm, models = rf.solve(rays, np.array([0, 1, 2], np.int32), 0)
    out["min_models_%d" % k] = models[:nm]
    for name, kw in (("van", dict(driver=1, max_num_iterations=2 ** 31 - 1)),
                     ("lo", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1))):
        opt = O.default_options(squared_inlier_threshold=THR2, inward=int(inward), **kw)
        res, inl = rf.estimate_pair(rays, opt, 100 + k)
        out["%s_E_%d" % (name, k)] = np.array(res.E)
        out["%s_r_%d" % (name, k)] = np.array(res.r)
        out["%s_stats_%d" % (name, k)] = np.array([res.num_iterations, res.best_num_inliers, res.number_lo_iterations])
        out["%s_inliers_%d" % (name, k)] = inl
    print(k, inward, noise, nm, out["van_stats_%d" % k], out["lo_stats_%d" % k])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_golden.npz"), **out)
