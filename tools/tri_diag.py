import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, spherical_sfm_b200 as S, oracle as O
orc = O.load(); eng = S.Engine(0)
cam, offs, oc, oxy, f, X = S.problems.make_tracks(5, 80, 400, obs_range=(1, 30), noise_px=0.5, outlier_frac=0.2)
rng = np.random.default_rng(0)
for p in (7, 19, 33):
    oxy[offs[p]:offs[p + 1]] = rng.uniform(-300, 300, (offs[p + 1] - offs[p], 2))
opt = S.default_options(squared_inlier_threshold=4.0, final_least_squares=1, first_pair_id=2)
pts, ninl, status, iters = eng.retriangulate(cam, offs, oc, oxy, f, opt)
oopt = O.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
bad = 0
for p in range(400):
    a, b = offs[p], offs[p + 1]
    res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, oopt, 2 + p)
    d = np.abs(pts[p] - np.array(res.E[:3])).max()
    if int(status[p]) != res.status or int(iters[p]) != res.num_iterations or int(ninl[p]) != res.best_num_inliers or d > 1e-7 * max(1, np.abs(pts[p]).max()):
        bad += 1
        print(p, "n", b - a, "status", status[p], res.status, "iters", iters[p], res.num_iterations, "ninl", ninl[p], res.best_num_inliers, "lo", res.number_lo_iterations, "d", d, pts[p], res.E[:3])
print("mismatches", bad)
