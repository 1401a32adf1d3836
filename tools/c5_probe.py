"""Config C5 probe: 8 pairs x N correspondences x 4096 hypotheses, 90 % outliers, k_score_models only (one launch per N)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, spherical_sfm_b200 as S
thr2 = (2.0 / 600.0) ** 2
eng = S.Engine(0)
peak = eng.measure_fp32_peak() if hasattr(eng, "measure_fp32_peak") else 72.5
for n5 in (10000, 20000, 50000, 100000, 200000):
    rays5, offs5, _ = S.problems.make_batch(500, 8, n5, noise=1 / 600, outlier_frac=0.9)
    m5 = np.zeros((8, 4096, 6))
    for p in range(8):
        samples = np.array([S.sample(3, p, i, 3, n5) for i in range(1024)], np.int32)
        mm, _ = eng.minimal_solve(rays5[offs5[p]:offs5[p + 1]], samples, 0)
        m5[p] = mm.reshape(-1, 6)
    best = min(eng.score_pairs(m5, rays5, offs5, thr2)[2] for _ in range(3))
    ev = 8 * 4096.0 * n5
    print("N=%d: %.3f ms  %.3e evals/s  %.3f of %.1f TFLOP/s" % (n5, best, ev / (best * 1e-3), ev * 42 / (best * 1e-3) / 1e12 / peak, peak))
