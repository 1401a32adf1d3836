"""Config C1 latency probe: one pair x 1000 correspondences, 50 % outliers, pipeline options; and small batches.
python tools/c1_probe.py   (honours the engine's env knobs, e.g. SSFM_NO_DEFER=1)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, spherical_sfm_b200 as S

thr2 = (2.0 / 600.0) ** 2
eng = S.Engine(0)
sizes = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64, 256, 1024]
for P in sizes:
    rays, offsets, probs = S.problems.make_batch(99, P, 1000, noise=1.0 / 600, outlier_frac=0.5, max_angle_deg=20.0)
    opt = S.pipeline_options(thr2)
    for _ in range(5):
        res, flags = eng.estimate_pairs(rays, offsets, opt)
    ts = []
    for _ in range(30 if P <= 1024 else 6):
        t0 = time.perf_counter(); res, flags = eng.estimate_pairs(rays, offsets, opt); ts.append(time.perf_counter() - t0)
    st = eng.stats()
    print("P=%d: median %.3f ms min %.3f ms  (solve %.3f score %.3f chain %.3f ms, %d launches, %d refit waves) iterations %s lo %s" % (
        P, 1e3 * np.median(ts), 1e3 * min(ts), st.solve_ms, st.score_ms, st.chain_ms, st.kernel_launches,
        getattr(st, "refit_waves", -1), res["num_iterations"][:4].tolist(), res["number_lo_iterations"][:4].tolist()))
