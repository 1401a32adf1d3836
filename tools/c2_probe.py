"""Config C2 probe: 1999 pairs x 2000 correspondences, 30 % outliers, 1 deg rotation, Sturm-variant solver, MSAC_FIXED M = 512."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, spherical_sfm_b200 as S
thr2 = (2.0 / 600.0) ** 2
P2, N2 = 1999, 2000
rays2, offs2, _ = S.problems.make_batch(2, P2, N2, noise=1 / 600, outlier_frac=0.3, rotation_deg=1.0)
opt2 = S.default_options(squared_inlier_threshold=thr2, driver=S.DRIVER_MSAC_FIXED, solver=S.SOLVER_FAST_STURM, fixed_budget=512)
eng = S.Engine(0)
eng.upload(rays2, offs2)
eng.run(opt2)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); eng.run(opt2); ts.append((time.perf_counter() - t0) * 1e3)
st = eng.stats()
r, _ = eng.download(want_flags=False)
print("C2: %.3f ms (min of 5)  solve %.3f score %.3f chain %.3f ms, %d rounds, %d launches; mean iterations %.1f; %.3e pairs/s" % (
    min(ts), st.solve_ms, st.score_ms, st.chain_ms, st.rounds, st.kernel_launches, r["num_iterations"].mean(), P2 / (min(ts) * 1e-3)))
