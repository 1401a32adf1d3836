"""Generates tests/golden/refsrc_golden.npz from the reference's OWN sources compiled into oracle/_ref:
  * libssfm_reflegacy.so : include/sphericalsfm/msac.h, preemptive_ransac.h + src/spherical_fast_estimator.cpp (config C2)
  * libssfm_reftri.so    : src/triangulation_estimator.cpp + RansacLib (SfM::Retriangulate)
  * libssfm_reffull.so   : RansacLib + src/spherical_estimator.cpp + src/spherical_solvers.cpp (config C1 / C3 pairs)
Run here (needs /root/reference); the .npz travels to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import spherical_sfm_b200 as S  # noqa: E402

THR2 = (2.0 / 600.0) ** 2
out = {}
# --- legacy drivers around the reference's SphericalFastEstimator
rl = O.load_ref_legacy()
assert rl is not None
legacy = [(2, 2000, 0.3, 512, 10, 41), (2, 300, 0.5, 200, 10, 42), (3, 2000, 0.3, 512, 10, 43), (3, 400, 0.4, 256, 7, 44)]
out["num_legacy"] = len(legacy)
for k, (drv, n, outl, M, B, pid) in enumerate(legacy):
    pr = S.problems.make_problem(S.problems.make_rng(2025, pid), n, False, 1.0, 1 / 600, int(outl * n), 20.0)
    opt = O.default_options(squared_inlier_threshold=THR2, driver=drv, solver_kind=2, legacy_budget=M, preemptive_block=B, random_seed=9)
    res, inl = rl.estimate_pair(pr.rays, opt, pid)
    out["lg_rays_%d" % k] = pr.rays
    out["lg_cfg_%d" % k] = np.array([drv, M, B, pid, 9], np.int64)
    out["lg_E_%d" % k] = np.array(res.E)
    out["lg_r_%d" % k] = np.array(res.r)
    out["lg_iters_%d" % k] = res.num_iterations
    out["lg_ninl_%d" % k] = res.best_num_inliers
    out["lg_inliers_%d" % k] = inl
    print("legacy", k, drv, res.status, res.num_iterations, res.best_num_inliers)
# --- Retriangulate
rt = O.load_ref_tri()
assert rt is not None
cam, offs, oc, oxy, f, X = S.problems.make_tracks(77, 50, 24, obs_range=(3, 20), noise_px=0.5, outlier_frac=0.15)
opt = O.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
pts, ninl, status, iters = [], [], [], []
for p in range(24):
    a, b = offs[p], offs[p + 1]
    res, inl = rt.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
    pts.append(res.E[:3]); ninl.append(res.best_num_inliers); status.append(res.status); iters.append(res.num_iterations)
out.update(tri_cam=cam, tri_offs=offs, tri_oc=oc, tri_oxy=oxy, tri_focal=f, tri_points=np.array(pts), tri_ninl=np.array(ninl),
           tri_status=np.array(status), tri_iters=np.array(iters))
print("triangulation", status, ninl)
# --- the whole 3-point path from the reference's estimator + solver sources (pipeline options)
rf = O.load_ref_full()
assert rf is not None
full = [(1000, 500, 51), (1500, 1050, 52)]
out["num_full"] = len(full)
for k, (n, nout, pid) in enumerate(full):
    pr = S.problems.make_problem(S.problems.make_rng(2026, pid), n, False, None, 1 / 600, nout, 20.0)
    opt = O.pipeline_options(THR2)
    res, inl = rf.estimate_pair(pr.rays, opt, pid)
    out["fu_rays_%d" % k] = pr.rays
    out["fu_pid_%d" % k] = pid
    out["fu_E_%d" % k] = np.array(res.E)
    out["fu_r_%d" % k] = np.array(res.r)
    out["fu_iters_%d" % k] = res.num_iterations
    out["fu_ninl_%d" % k] = res.best_num_inliers
    out["fu_nlo_%d" % k] = res.number_lo_iterations
    out["fu_inliers_%d" % k] = inl
    print("full", k, res.status, res.num_iterations, res.best_num_inliers, res.number_lo_iterations)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refsrc_golden.npz"), **out)
