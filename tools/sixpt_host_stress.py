"""CPU: the product's six-point solver (host build, tests/hostshim) against oracle/sixpt_oracle.py on RANSAC-like samples
(random 6-subsets of noisy correspondences with outliers).  python tools/sixpt_host_stress.py [samples] [seed]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sixpt_oracle as X  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "tests", "hostshim", "libhostshim.so"))
dp = C.POINTER(C.c_double)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = tot = nsol = 0
t_h = t_o = 0.0
while tot < ns:
    rays, R, t, f = X.make_problem(rng, 200, rng.uniform(400, 1200), outlier_frac=0.5, noise_px=0.5)
    for _ in range(50):
        idx = rng.choice(200, 6, replace=False)
        r6 = np.ascontiguousarray(rays[idx])
        m = np.zeros((15, 7))
        G = np.zeros((15, 9))
        t0 = time.perf_counter()
        n = lib.hs_sixpt_solve(r6.ctypes.data_as(dp), m.ctypes.data_as(dp), G.ctypes.data_as(dp), 1)
        t_h += time.perf_counter() - t0
        t0 = time.perf_counter()
        b = X.minimal_solver(r6)
        t_o += time.perf_counter() - t0
        tot += 1
        nsol += len(b)
        ok = n == len(b)
        if ok:
            for i, (tt, r, ff) in enumerate(b):
                d = max(np.abs(m[i, :3] - tt).max(), np.abs(m[i, 3:6] - r).max(), abs(m[i, 6] - ff) / ff)
                ok = ok and d < 1e-6
        if not ok:
            bad += 1
            if bad <= 5:
                print("mismatch at sample", tot, "host", n, "oracle", len(b), [x[2] for x in b], m[:n, 6])
print("samples %d, oracle solutions %d, mismatching samples %d; host solver %.1f us/sample, oracle %.1f ms/sample" % (
    tot, nsol, bad, t_h / tot * 1e6, t_o / tot * 1e3))
