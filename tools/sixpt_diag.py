"""GPU diagnostic: compare device six-point solutions with the numpy oracle sample by sample."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, sixpt_oracle as X, oracle as O, spherical_sfm_b200 as S
orc = O.load(); eng = S.Engine(0)
rng = np.random.default_rng(6)
rays, R, t, f = X.make_problem(rng, 200, rng.uniform(400, 1200), outlier_frac=0.3, noise_px=0.5)
samples = np.array([orc.philox_sample(7, 11, it, 6, 200) for it in range(10000)])
models, nm = eng.sixpt_solve(rays, samples)
bad = 0
for it in range(len(samples)):
    b = X.minimal_solver(rays[samples[it]])
    if len(b) != nm[it]:
        bad += 1
        u = rays[samples[it]][:, :3]; v = rays[samples[it]][:, 3:]
        s = np.sqrt((np.sum(u[:, :2] ** 2) + np.sum(v[:, :2] ** 2)) / 12.0); D = np.diag([1 / s, 1 / s, 1.0])
        F0, F1, F2 = X.nullspace_basis(u @ D, v @ D); M0, M1, M2 = X.constraint_matrices(F0, F1, F2)
        print(it, "device", nm[it], models[it, :nm[it], 6], "oracle", len(b), [m[2] for m in b], "xyw", X.solve_xyw(M0, M1, M2))
print("mismatching samples:", bad, "of", len(samples))
