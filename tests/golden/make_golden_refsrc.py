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
# --- the whole 3-point path from the reference's estimator + solver sources:
#       C1 shape (1000 corr, 50 % outliers), C3 shape (1500 corr, 70 % outliers) with estimate_pairwise's options
#       (examples/spherical_sfm_tools.cpp:314-318), C3 shape with make_loop_closures' options (final_least_squares_ = false,
#       :609-615), inward pairs, and the polynomial solver.
#     Each case is run twice: "up" = upstream as written (models of complex eigenvalues = Re of Eigen's eigenvector, Eigen 3.4
#     restated in oracle/eigen_shim) and "skip" = the same sources with those models masked in the Eigen stand-in (upstream's
#     own commented-out filter, src/spherical_solvers.cpp:294).  "skip" is the exact pin; "up" differs from any second
#     implementation wherever such a model wins an iteration (its phase is rounding noise), and
#     fu_same_canonical records on which cases the restated oracle (canonical representative) nevertheless follows it.
#     The rays are not stored: they come from problems.make_problem's portable stream (sha256 recorded).
import hashlib  # noqa: E402

rf = O.load_ref_full()
orc = O.load()
assert rf is not None
full = []
for i in range(64):
    full.append((1000, 500, 1000 + i, 0, 1, 0))      # C1
for i in range(64):
    full.append((1500, 1050, 2000 + i, 0, 1, 0))     # C3, estimate_pairwise options
for i in range(16):
    full.append((1500, 1050, 3000 + i, 0, 0, 0))     # C3, make_loop_closures options
for i in range(8):
    full.append((800, 400, 4000 + i, 1, 1, 0))       # inward
for i in range(16):
    full.append((1000, 500, 5000 + i, 0, 1, 1))      # polynomial solver (deterministic upstream: no mask needed)
cfg = np.array(full, np.int64)
K = len(full)
stats = {m: np.zeros((K, 4), np.int64) for m in ("up", "skip")}
rr = {m: np.zeros((K, 3)) for m in ("up", "skip")}
EE = {m: np.zeros((K, 9)) for m in ("up", "skip")}
score = {m: np.zeros(K) for m in ("up", "skip")}
inl_bits = {m: [] for m in ("up", "skip")}
sha = []
same_canon = np.zeros(K, np.uint8)
same_skip = np.zeros(K, np.uint8)


def same(a, ia, b, ib):
    return (a.status == b.status and a.num_iterations == b.num_iterations and a.best_num_inliers == b.best_num_inliers and
            a.number_lo_iterations == b.number_lo_iterations and len(ia) == len(ib) and bool((ia == ib).all()))


for k, (n, nout, pid, inward, flsq, kind) in enumerate(full):
    pr = S.problems.make_problem(S.problems.make_rng(2026, pid), n, bool(inward), None, 1 / 600, nout, 20.0)
    sha.append(hashlib.sha256(np.ascontiguousarray(pr.rays).tobytes()).hexdigest())
    res = {}
    for mode, cm in (("up", O.COMPLEX_EIGEN), ("skip", O.COMPLEX_SKIP)):
        opt = O.default_options(squared_inlier_threshold=THR2, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=flsq,
                                inward=inward, solver_kind=kind, complex_mode=cm)
        r, inl = rf.estimate_pair(pr.rays, opt, pid)
        res[mode] = (r, inl)
        stats[mode][k] = (r.status, r.num_iterations, r.best_num_inliers, r.number_lo_iterations)
        rr[mode][k] = np.array(r.r)
        EE[mode][k] = np.array(r.E)
        score[mode][k] = r.best_model_score
        f = np.zeros(n, np.uint8)
        f[inl] = 1
        inl_bits[mode].append(np.packbits(f))
    oc = O.default_options(squared_inlier_threshold=THR2, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=flsq,
                           inward=inward, solver_kind=kind, complex_mode=O.COMPLEX_CANONICAL)
    a, ia = orc.estimate_pair(pr.rays, oc, pid)
    same_canon[k] = same(a, ia, *res["up"])
    oc.complex_mode = O.COMPLEX_SKIP
    a, ia = orc.estimate_pair(pr.rays, oc, pid)
    same_skip[k] = same(a, ia, *res["skip"])
    print("full", k, full[k], stats["up"][k], stats["skip"][k], same_canon[k], same_skip[k], flush=True)
out["fu_cfg"] = cfg
out["fu_sha"] = np.array(sha)
for m in ("up", "skip"):
    out["fu_stats_" + m] = stats[m]
    out["fu_r_" + m] = rr[m]
    out["fu_E_" + m] = EE[m]
    out["fu_score_" + m] = score[m]
    out["fu_inl_" + m] = np.concatenate(inl_bits[m])
    out["fu_inl_off_" + m] = np.concatenate([[0], np.cumsum([len(x) for x in inl_bits[m]])]).astype(np.int64)
out["fu_same_canonical"] = same_canon
print("oracle (canonical) follows upstream-as-written on %d / %d cases; oracle (skip) follows masked upstream on %d / %d"
      % (same_canon.sum(), K, same_skip.sum(), K))
for lo, hi, name in ((0, 64, "C1"), (64, 128, "C3"), (128, 144, "C3 loop-closure options"), (144, 152, "inward"), (152, 168, "polynomial")):
    print("  %-24s canonical %d / %d   skip %d / %d" % (name, same_canon[lo:hi].sum(), hi - lo, same_skip[lo:hi].sum(), hi - lo))
assert same_skip.all()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refsrc_golden.npz"), **out)
