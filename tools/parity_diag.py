"""GPU-box diagnostic: run the at-scale parity batches through the engine (default and wide FP32 pre-filter margins) and the
oracle, and dump every pair on which they differ (rays + both results) to gpurun_out/ for analysis on the CPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle as O  # noqa: E402
import spherical_sfm_b200 as S  # noqa: E402
from conftest import THR2, to_oracle_options  # noqa: E402

orc = O.load()
eng = S.Engine(0)
out = {}
k = 0
for name, P, N, outl, kw in (("C3", 16384, 1500, 0.7, dict(final_least_squares=1)), ("C3lc", 4096, 1500, 0.7, dict(final_least_squares=0)),
                            ("C1", 2048, 1000, 0.5, dict(final_least_squares=1))):
    rays_t, offsets, _ = bench.make_batch_torch(P, N, outl, seed=1000 + P, device="cuda")
    rays = rays_t.cpu().numpy()
    opt = S.default_options(squared_inlier_threshold=THR2, num_lo_steps=0, num_lsq_iterations=0, **kw)
    tabs = {}
    for margin in ("2e-4", "1e-2", "1e9"):
        os.environ["SSFM_CAND_MARGIN"] = margin
        res, flags = eng.estimate_pairs(rays, offsets, opt)
        tabs[margin] = (res.copy(), flags.copy())
    del os.environ["SSFM_CAND_MARGIN"]
    ores, oflags, secs = orc.estimate_batch_flags(rays, offsets, to_oracle_options(O, opt), 0)
    o = np.array([(r.status, r.num_iterations, r.best_num_inliers, r.number_lo_iterations) for r in ores], np.int64)
    for margin, (res, flags) in tabs.items():
        mine = np.stack([res["status"], res["num_iterations"], res["best_num_inliers"], res["number_lo_iterations"]], 1).astype(np.int64)
        bad = np.nonzero((o != mine).any(axis=1))[0]
        print(name, "margin", margin, "mismatching pairs:", bad.tolist(), "flags equal:", bool((flags == oflags).all()),
              "identical to default table:", res.tobytes() == tabs["2e-4"][0].tobytes(), flush=True)
        for p in bad:
            print("   pair", p, "oracle", o[p].tolist(), "gpu", mine[p].tolist())
            if margin == "2e-4" or True:
                out["rays_%d" % k] = rays[offsets[p]:offsets[p + 1]]
                out["meta_%d" % k] = np.array([P, p, kw["final_least_squares"], float(margin)])
                out["oracle_%d" % k] = o[p]
                out["gpu_%d" % k] = mine[p]
                k += 1
out["num"] = k
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "parity_diag.npz"), **out)
