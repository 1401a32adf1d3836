"""Retriangulate timing probe: P points, ragged tracks."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, spherical_sfm_b200 as S, oracle as O
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
base = S.problems.make_tracks(7, 200, 2000, obs_range=(3, 30), noise_px=0.5, outlier_frac=0.2)
cam, offs0, oc0, oxy0, f, X0 = base
reps = (P + 1999) // 2000
oc = np.tile(oc0, reps); oxy = np.tile(oxy0, (reps, 1))
lens = np.tile(np.diff(offs0), reps)[:P]
offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
oc = oc[:offs[-1]]; oxy = oxy[:offs[-1]]
eng = S.Engine(0)
opt = S.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
for rep in range(3):
    t0 = time.time(); pts, ninl, status, iters = eng.retriangulate(cam, offs, oc, oxy, f, opt); dt = time.time() - t0
    print(f"rep {rep}: {P} points, {offs[-1]} observations: {dt*1e3:.1f} ms -> {P/dt:.3e} points/s; ok {np.mean(status==0):.3f} mean iters {iters.mean():.1f}")
orc = O.load(); oopt = O.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
t0 = time.time(); n = 300
for p in range(n):
    a, b = offs[p], offs[p + 1]
    orc.triangulate(cam[oc[a:b]], oxy[a:b], f, oopt, p)
print(f"oracle (1 thread): {n/(time.time()-t0):.1f} points/s")
