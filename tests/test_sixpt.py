"""Six-point shared-focal estimator: numpy oracle (oracle/sixpt_oracle.py,
LAPACK eigen-solver) vs the product's own solver (Hessenberg-QR; host build in tests/hostshim without a GPU, the
device build through the C ABI with one).  PoseLib's arithmetic is absent from the reference tree, so the pins are
the ground truth of synthetic problems and the agreement of two independent implementations."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sixpt_oracle as X  # noqa: E402


def _host_solve(shim_lib, rays, focal_scoring=0):
    rays = np.ascontiguousarray(rays, np.float64)
    m = np.zeros((15, 7))
    G = np.zeros((15, 9))
    dp = C.POINTER(C.c_double)
    n = shim_lib.hs_sixpt_solve(rays.ctypes.data_as(dp), m.ctypes.data_as(dp), G.ctypes.data_as(dp), focal_scoring)
    return m[:n], G[:n]


def _model_diff(a, b):
    """a: 7-vector (t, r, f); b: oracle tuple"""
    t, r, f = b
    return max(np.abs(a[:3] - t).max(), np.abs(a[3:6] - r).max(), abs(a[6] - f) / f)


def test_oracle_solver_recovers_ground_truth():
    rng = np.random.default_rng(1)
    for k in range(60):
        rays, R, t, f = X.make_problem(rng, 6, rng.uniform(400, 1200))
        sols = X.minimal_solver(rays)
        assert 1 <= len(sols) <= 15
        best = min(max(np.linalg.norm(X.so3exp(r) - R), np.linalg.norm(tt - t), abs(ff - f) / f) for tt, r, ff in sols)
        assert best < 1e-6
        # every solution satisfies the six epipolar constraints with its own focal
        for m in sols:
            F = X.scoring_matrix(m, focal_scoring=True)
            assert np.abs(X.sampson(F, rays)).max() < 1e-12 * f * f


def test_product_solver_matches_oracle_on_host(shim):
    """csrc/ssfm_sixpt.cuh compiled for the host (tests only): same number of solutions, same models to 1e-9
    (north_star bar: 1e-5 after root matching), noise-free and noisy samples, and the scoring matrices."""
    rng = np.random.default_rng(2)
    nsol = 0
    for k in range(200):
        rays, R, t, f = X.make_problem(rng, 6, rng.uniform(400, 1200), noise_px=0.0 if k < 100 else 1.0)
        a, G = _host_solve(shim.lib, rays, focal_scoring=k % 2)
        b = X.minimal_solver(rays)
        assert len(a) == len(b), k
        nsol += len(a)
        for i, mb in enumerate(b):
            assert _model_diff(a[i], mb) < 1e-9, (k, i)
            Go = X.scoring_matrix(mb, focal_scoring=bool(k % 2))
            assert np.abs(G[i].reshape(3, 3) - Go).max() < 1e-8 * np.abs(Go).max()
    assert nsol > 200


def _host_ls(shim_lib, rays, sample, model):
    rays = np.ascontiguousarray(rays, np.float64)
    sample = np.ascontiguousarray(sample, np.int32)
    m = np.concatenate([model[0], model[1], [model[2]]]).astype(np.float64)
    costs = np.zeros(2)
    dp = C.POINTER(C.c_double)
    it = shim_lib.hs_sixpt_least_squares(rays.ctypes.data_as(dp), sample.ctypes.data_as(C.POINTER(C.c_int)), len(sample),
                                         m.ctypes.data_as(dp), costs.ctypes.data_as(dp))
    return m, it, costs


def _host_lo_msac(shim, rays, kw, focal_scoring, thr2, seed, pair_id=0):
    from conftest import HsParams
    rays = np.ascontiguousarray(rays, np.float64)
    n = len(rays)
    hp = HsParams(min_iters=kw.get("min_num_iterations", 100), max_iters=kw.get("max_num_iterations", 10000),
                  success_probability=0.9999, thr2=thr2, seed=seed, num_lo_steps=kw.get("num_lo_steps", 10), thr_mult=2 ** 0.5,
                  num_lsq_iters=kw.get("num_lsq_iterations", 4), min_sample_mult=7, non_min_mult=3,
                  lo_start=kw.get("lo_starting_iterations", 50), final_lsq=kw.get("final_least_squares", 0))
    out7, flags = np.zeros(7), np.zeros(max(n, 1), np.uint8)
    sc, it, nlo = C.c_double(), C.c_uint32(), C.c_int()
    dp = C.POINTER(C.c_double)
    shim.lib.hs_sixpt_lo_msac.restype = C.c_int
    ninl = shim.lib.hs_sixpt_lo_msac(rays.ctypes.data_as(dp), n, C.byref(hp), C.c_uint32(pair_id), focal_scoring,
                                     out7.ctypes.data_as(dp), C.byref(sc), C.byref(it), C.byref(nlo),
                                     flags.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return dict(model=out7, score=sc.value, num_iterations=it.value, number_lo_iterations=nlo.value, best_num_inliers=ninl,
                flags=flags[:n])


def _oracle_lo_msac(sampler, rays, kw, focal_scoring, thr2, seed):
    return X.lo_msac(rays, sampler, thr2, seed=seed, min_iters=kw.get("min_num_iterations", 100),
                     max_iters=kw.get("max_num_iterations", 10000), num_lo_steps=kw.get("num_lo_steps", 10),
                     num_lsq_iterations=kw.get("num_lsq_iterations", 4), lo_starting_iterations=kw.get("lo_starting_iterations", 50),
                     final_least_squares=bool(kw.get("final_least_squares", 0)), focal_scoring=bool(focal_scoring))


LO_CASES = [  # (seed, N, outliers, options)
    (1, 300, 0.4, dict(num_lo_steps=2, num_lsq_iterations=2, lo_starting_iterations=20, final_least_squares=1,
                       min_num_iterations=50, max_num_iterations=400)),
    (2, 300, 0.3, dict()),  # RansacLib defaults: 10 LO steps x 4 LSQ iterations, LO from iteration 50
    (3, 150, 0.2, dict(num_lo_steps=1, num_lsq_iterations=2, lo_starting_iterations=200, min_num_iterations=60,
                       max_num_iterations=60, final_least_squares=1)),  # the loop ends before lo_start: LO after the loop (:246-257)
]


def test_lo_msac_product_on_host_matches_oracle(shim, S):
    """LocallyOptimizedMSAC around SixPointEstimator (NonMinimalSolver / LeastSquares, six_point_estimator.cpp:121-192):
    csrc/ssfm_sixpt_lo.cuh compiled for the host -- the functions k_sixpt_lo runs per parked pair -- against the numpy
    restatement of ransac.h: same iteration and LO counts, same inlier set, same model."""
    for seed, N, outl, kw in LO_CASES:
        rays, offs, f, R, t = S.problems.make_sixpt_batch(seed, 1, N, outlier_frac=outl)
        got = _host_lo_msac(shim, rays, kw, 1, 4.0, 7)

        def sampler(i):
            idx = np.zeros(6, np.int32)
            shim.lib.hs_sample(C.c_uint32(7), C.c_uint32(0), C.c_uint32(i), 6, N, idx.ctypes.data_as(C.POINTER(C.c_int)))
            return idx
        st = _oracle_lo_msac(sampler, rays, kw, 1, 4.0, 7)
        assert got["num_iterations"] == st["num_iterations"], seed
        assert got["number_lo_iterations"] == st["number_lo_iterations"] and st["number_lo_iterations"] >= 1, seed
        assert got["best_num_inliers"] == st["best_num_inliers"], seed
        assert np.nonzero(got["flags"])[0].tolist() == st["inliers"].tolist(), seed
        assert _model_diff(got["model"], st["model"]) < 1e-7, seed
        assert abs(got["score"] - st["best_model_score"]) <= 1e-8 * st["best_model_score"], seed


class _ToyLine:
    """The toy estimator of oracle/ref_toy.cpp (a 2-D line: minimal = through two points (+ a second, shifted hypothesis),
    non-minimal and LeastSquares = closed-form total least squares with sequential sums), restated for lo_msac_generic."""
    min_sample_size, non_minimal_sample_size = 2, 3

    def __init__(self, xy):
        self.xy, self.n = xy, len(xy)

    def errors(self, m):
        d = m[0] * self.xy[:, 0] + m[1] * self.xy[:, 1] + m[2]
        return d * d

    def minimal_solver(self, sample):
        import math
        (x0, y0), (x1, y1) = self.xy[sample[0]], self.xy[sample[1]]
        dx, dy = x1 - x0, y1 - y0
        nrm = math.sqrt(dx * dx + dy * dy)
        if not nrm > 0:
            return []
        a, b = dy / nrm, -dx / nrm
        c = -(a * x0 + b * y0)
        return [(a, b, c), (a, b, c + 0.25)]

    def _fit(self, sample):
        import math
        m = len(sample)
        if m < 2:
            return None
        mx = my = 0.0
        for i in sample:
            mx += self.xy[i, 0]
            my += self.xy[i, 1]
        mx /= m
        my /= m
        sxx = sxy = syy = 0.0
        for i in sample:
            dx, dy = self.xy[i, 0] - mx, self.xy[i, 1] - my
            sxx += dx * dx
            sxy += dx * dy
            syy += dy * dy
        if sxx + syy <= 0:
            return None
        th = 0.5 * math.atan2(2 * sxy, sxx - syy)
        a, b = -math.sin(th), math.cos(th)
        return (a, b, -(a * mx + b * my))

    def non_minimal_solver(self, sample):
        return self._fit([int(i) for i in sample])

    def least_squares(self, sample, m):
        r = self._fit([int(i) for i in sample])
        return m if r is None else r


def test_restated_lo_msac_driver_equals_reference_header():
    """The Python restatement of LocallyOptimizedMSAC that drives the six-point oracle (sixpt_oracle.lo_msac_generic) against the
    reference's own include/RansacLib/ransac.h, compiled unmodified around the same toy estimator (oracle/_ref/libssfm_reftoy.so):
    identical iteration counts, LO counts, inlier lists, scores and models over option sets that reach every branch -- the LO at
    lo_starting_iterations_ (:166-177), the LO after a loop that ends before it (:246-257), final_least_squares_ (:259-275),
    num_lo_steps 0, samples that resize beyond the inlier list, and the shared mt19937 / uniform_int stream throughout."""
    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libssfm_reftoy.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libssfm_reftoy.so not built (needs the reference tree)")
    lib = C.CDLL(lib_path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    cases = [  # (n, outlier fraction, noise, options)
        (200, 0.4, 0.01, dict()),
        (200, 0.6, 0.02, dict(num_lo_steps=3, num_lsq_iterations=2, lo_starting_iterations=10, final_least_squares=True)),
        (60, 0.3, 0.01, dict(num_lo_steps=2, num_lsq_iterations=3, lo_starting_iterations=500, min_iters=40, max_iters=40,
                             final_least_squares=True)),
        (120, 0.5, 0.02, dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=True)),
        (12, 0.2, 0.005, dict(num_lo_steps=4, lo_starting_iterations=5, min_iters=30, max_iters=200)),
    ]
    for ci, (n, outl, noise, kw) in enumerate(cases):
        for seed in range(4):
            rng = np.random.default_rng(1000 * ci + seed)
            t = rng.uniform(-1, 1, n)
            ang = rng.uniform(0, np.pi)
            xy = np.stack([t * np.cos(ang) + 0.3, t * np.sin(ang) - 0.2], 1) + noise * rng.standard_normal((n, 2))
            out = rng.random(n) < outl
            xy[out] = rng.uniform(-1.5, 1.5, (int(out.sum()), 2))
            xy = np.ascontiguousarray(xy)
            thr2 = (3 * noise) ** 2
            o = dict(num_lo_steps=10, num_lsq_iterations=4, min_sample_multiplicator=7, non_min_sample_multiplier=3,
                     lo_starting_iterations=50, final_least_squares=False, min_iters=100, max_iters=10000)
            o.update(kw)
            model, score, stats, inl = np.zeros(3), C.c_double(), np.zeros(3, np.int32), np.zeros(n, np.int32)
            lib.ref_toy_lomsac.restype = C.c_int
            ninl = lib.ref_toy_lomsac(xy.ctypes.data_as(dp), n, C.c_double(thr2), C.c_uint(7 + seed), o["num_lo_steps"],
                                      o["num_lsq_iterations"], o["min_sample_multiplicator"], o["non_min_sample_multiplier"],
                                      C.c_uint(o["lo_starting_iterations"]), int(o["final_least_squares"]), C.c_uint(o["min_iters"]),
                                      C.c_uint(o["max_iters"]), model.ctypes.data_as(dp), C.byref(score), stats.ctypes.data_as(ip),
                                      inl.ctypes.data_as(ip))

            def sampler(it):
                first = (7 * it) % n
                return [first, (first + 1 + (13 * it) % (n - 1)) % n]
            st = X.lo_msac_generic(_ToyLine(xy), sampler, thr2, seed=7 + seed, **o)
            tag = (ci, seed)
            assert st["num_iterations"] == stats[0], tag
            assert st["number_lo_iterations"] == stats[1], tag
            assert st["best_num_inliers"] == stats[2] == ninl, tag
            assert st["inliers"].tolist() == inl[:ninl].tolist(), tag
            assert abs(st["best_model_score"] - score.value) <= 1e-12 * max(score.value, 1e-300), tag
            assert np.abs(np.array(st["model"]) - model).max() < 1e-12, tag


def test_restated_vanilla_msac_driver_equals_reference_header():
    """The driver of config C4: evaluation/vanilla_ransac.h compiled unmodified around the toy estimator against its Python
    restatement (sixpt_oracle.vanilla_msac_generic, what the six-point oracle and the device tests run): identical iteration
    counts, inlier lists, scores and models, including runs that stop at min / max iterations and a 12-point problem."""
    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libssfm_reftoy.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libssfm_reftoy.so not built (needs the reference tree)")
    lib = C.CDLL(lib_path)
    lib.ref_toy_vanilla.restype = C.c_int
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    for ci, (n, outl, noise, min_it, max_it) in enumerate([(200, 0.4, 0.01, 100, 10000), (300, 0.7, 0.02, 100, 10000),
                                                           (80, 0.5, 0.01, 30, 60), (12, 0.2, 0.005, 20, 500)]):
        for seed in range(4):
            rng = np.random.default_rng(2000 * ci + seed)
            t = rng.uniform(-1, 1, n)
            ang = rng.uniform(0, np.pi)
            xy = np.stack([t * np.cos(ang) + 0.3, t * np.sin(ang) - 0.2], 1) + noise * rng.standard_normal((n, 2))
            out = rng.random(n) < outl
            xy[out] = rng.uniform(-1.5, 1.5, (int(out.sum()), 2))
            xy = np.ascontiguousarray(xy)
            thr2 = (3 * noise) ** 2
            model, score, stats, inl = np.zeros(3), C.c_double(), np.zeros(3, np.int32), np.zeros(n, np.int32)
            ninl = lib.ref_toy_vanilla(xy.ctypes.data_as(dp), n, C.c_double(thr2), C.c_uint(seed), C.c_uint(min_it), C.c_uint(max_it),
                                       model.ctypes.data_as(dp), C.byref(score), stats.ctypes.data_as(ip), inl.ctypes.data_as(ip))

            def sampler(it):
                first = (7 * it) % n
                return [first, (first + 1 + (13 * it) % (n - 1)) % n]
            st = X.vanilla_msac_generic(_ToyLine(xy), sampler, thr2, min_iters=min_it, max_iters=max_it)
            tag = (ci, seed)
            assert st["num_iterations"] == stats[0] and st["best_num_inliers"] == stats[2] == ninl, tag
            assert st["inliers"].tolist() == inl[:ninl].tolist(), tag
            assert abs(st["best_model_score"] - score.value) <= 1e-12 * max(score.value, 1e-300), tag
            assert np.abs(np.array(st["model"]) - model).max() < 1e-12, tag


def _refit_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n_cases:
        rays, R, t, f = X.make_problem(rng, 60, rng.uniform(400, 1200), noise_px=0.5)
        sols = X.minimal_solver(rays[:6])
        if sols:
            out.append((rays, R, t, f, min(sols, key=lambda m: np.linalg.norm(X.so3exp(m[1]) - R))))
    return out


def test_least_squares_product_matches_oracle_and_improves_the_model(shim):
    """SixPointEstimator::LeastSquares: the product's LM (analytic chain rule through E-jets, Householder sphere
    manifold) against the numpy oracle (dual numbers): same iteration counts, same refined model; and the refit moves
    a minimal-sample model towards the truth while keeping |t| = 1."""
    better = 0
    cases = _refit_cases(12, 5)
    for rays, R, t, f, m0 in cases:
        sample = np.arange(6, 48)
        mo, ito, c0, c1 = X.least_squares(rays, sample, m0)
        mh, ith, costs = _host_ls(shim.lib, rays, sample, m0)
        assert ito == ith and abs(costs[1] - c1) <= 1e-9 * max(c1, 1e-30)
        assert max(np.abs(mo[0] - mh[:3]).max(), np.abs(mo[1] - mh[3:6]).max(), abs(mo[2] - mh[6]) / mo[2]) < 1e-6
        assert abs(np.linalg.norm(mh[:3]) - 1.0) < 1e-12 and c1 <= c0
        e0 = np.linalg.norm(X.so3ln(X.so3exp(m0[1]).T @ R))
        e1 = np.linalg.norm(X.so3ln(X.so3exp(mo[1]).T @ R))
        better += e1 < e0
    assert better >= 10


def test_oracle_vanilla_msac_recovers_focal_and_rotation():
    """Config C4 shape at reduced size: pixel-unit rays, unknown shared focal, 30 % outliers."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    orc = O.load()
    rng = np.random.default_rng(3)
    for p in range(2):
        f = rng.uniform(400, 1200)
        rays, R, t, f = X.make_problem(rng, 300, f, outlier_frac=0.3, noise_px=0.5)
        st = X.vanilla_msac(rays, lambda it: orc.philox_sample(7, p, it, 6, len(rays)), 4.0, focal_scoring=True)
        assert st["status"] == 0 and st["best_num_inliers"] >= 0.6 * 300
        tt, r, ff = st["model"]
        # a minimal-sample model without refit: the focal is only weakly constrained by the Sampson error
        assert 0.5 < ff / f < 2.0
        assert np.rad2deg(np.linalg.norm(X.so3ln(X.so3exp(r).T @ R))) < 10.0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_solver_matches_oracle(S, engine):
    rng = np.random.default_rng(4)
    rays_all, samples = [], []
    for k in range(64):
        rays, R, t, f = X.make_problem(rng, 6, rng.uniform(400, 1200), noise_px=0.0 if k < 32 else 1.0)
        samples.append(np.arange(6) + 6 * k)
        rays_all.append(rays)
    rays_all = np.concatenate(rays_all)
    models, nm = engine.sixpt_solve(rays_all, np.array(samples))
    for k in range(64):
        b = X.minimal_solver(rays_all[6 * k:6 * k + 6])
        assert nm[k] == len(b), k
        for i, mb in enumerate(b):
            assert _model_diff(models[k, i], mb) < 1e-9


@pytest.mark.gpu
def test_device_least_squares_matches_oracle(S, engine):
    cases = _refit_cases(8, 6)
    for rays, R, t, f, m0 in cases:
        sample = np.arange(6, 40)
        mo, ito, c0, c1 = X.least_squares(rays, sample, m0)
        md = engine.sixpt_least_squares(rays, [sample], np.concatenate([m0[0], m0[1], [m0[2]]])[None])[0]
        assert max(np.abs(mo[0] - md[:3]).max(), np.abs(mo[1] - md[3:6]).max(), abs(mo[2] - md[6]) / mo[2]) < 1e-6


def _run_c4(S, engine, orc, P, N, outl, focal_scoring, seed):
    rng = np.random.default_rng(seed)
    parts, truth = [], []
    for p in range(P):
        rays, R, t, f = X.make_problem(rng, N, rng.uniform(400, 1200), outlier_frac=outl, noise_px=0.5)
        parts.append(rays)
        truth.append((R, t, f))
    rays = np.concatenate(parts)
    offsets = (np.arange(P + 1) * N).astype(np.int64)
    thr2 = 4.0 if focal_scoring else 1e-3
    opt = S.default_options(squared_inlier_threshold=thr2, driver=S.DRIVER_VANILLA_MSAC, solver=S.SOLVER_SIXPT_FOCAL,
                            sixpt_focal_scoring=focal_scoring, random_seed=7, first_pair_id=11)
    res, flags = engine.estimate_pairs(rays, offsets, opt)
    for p in range(P):
        pr = rays[offsets[p]:offsets[p + 1]]
        st = X.vanilla_msac(pr, lambda it: orc.philox_sample(7, 11 + p, it, 6, N), thr2, focal_scoring=bool(focal_scoring))
        assert int(res["status"][p]) == st["status"]
        assert int(res["num_iterations"][p]) == st["num_iterations"], p
        # two independent eigen-solvers may disagree on a near-double root once in ~10^4 samples
        assert abs(int(res["evals"][p]) - st["evals"]) <= 2 * N, p
        if st["status"] != 0:
            continue
        tt, r, ff = st["model"]
        assert max(np.abs(res["t"][p] - tt).max(), np.abs(res["r"][p] - r).max(), abs(res["focal"][p] - ff) / ff) < 1e-7
        assert abs(res["best_model_score"][p] - st["best_model_score"]) <= 1e-9 * st["best_model_score"]
        # inlier masks: identical except for points within 1e-6 (relative) of the threshold (north_star)
        fl = np.zeros(N, np.uint8)
        fl[st["inliers"]] = 1
        diff = np.nonzero(flags[offsets[p]:offsets[p + 1]] != fl)[0]
        assert (np.abs(st["errors"][diff] - thr2) <= 1e-6 * thr2).all()
        assert abs(int(res["best_num_inliers"][p]) - st["best_num_inliers"]) <= len(diff)
    return res, truth


@pytest.mark.gpu
def test_config_c4_six_point_vanilla_msac(S, engine, orc):
    """Config C4 (1000 correspondences, 50 % outliers, shared focal U[400,1200], VanillaMSAC) on the device vs the
    numpy oracle loop: same iteration counts, models, costs, inlier masks; and the focal / rotation are right."""
    res, truth = _run_c4(S, engine, orc, 3, 1000, 0.5, 1, 5)
    for p, (R, t, f) in enumerate(truth):
        assert 0.5 < res["focal"][p] / f < 2.0  # minimal-sample model, no refit: focal weakly constrained
        assert np.rad2deg(np.linalg.norm(X.so3ln(X.so3exp(res["r"][p]).T @ R))) < 10.0
        assert res["best_num_inliers"][p] >= 0.45 * 1000


@pytest.mark.gpu
def test_six_point_reference_literal_scoring_and_small_pairs(S, engine, orc):
    """focal_scoring = 0 is SixPointEstimator::EvaluateModelOnPoint as written upstream (E on the raw rays)."""
    _run_c4(S, engine, orc, 2, 200, 0.3, 0, 6)
    _run_c4(S, engine, orc, 4, 120, 0.2, 1, 8)


@pytest.mark.gpu
def test_six_point_edge_cases_and_driver_check(S, engine):
    rng = np.random.default_rng(9)
    sizes = [0, 5, 6, 40]
    parts = [X.make_problem(rng, max(n, 1), 700.0, noise_px=0.2)[0][:n] for n in sizes]
    rays = np.concatenate(parts)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    opt = S.default_options(squared_inlier_threshold=4.0, driver=S.DRIVER_VANILLA_MSAC, solver=S.SOLVER_SIXPT_FOCAL,
                            sixpt_focal_scoring=1)
    res, flags = engine.estimate_pairs(rays, offsets, opt)
    assert list(res["status"][:2]) == [1, 1]
    assert res["status"][3] == 0 and res["best_num_inliers"][3] >= 30 and 350 < res["focal"][3] < 1400
    assert res["num_iterations"][2] == 10000 or res["status"][2] == 0  # six points: every sample is the same set
    opt.driver = S.DRIVER_LO_MSAC  # the same edge cases under LO-MSAC
    res2, flags2 = engine.estimate_pairs(rays, offsets, opt)
    assert list(res2["status"][:2]) == [1, 1]
    assert res2["status"][3] == 0 and res2["best_num_inliers"][3] >= 30 and 350 < res2["focal"][3] < 1400
    opt.driver = S.DRIVER_MSAC_FIXED
    with pytest.raises(S.SsfmError):
        engine.estimate_pairs(rays, offsets, opt)


@pytest.mark.gpu
def test_six_point_lo_msac_batched_matches_oracle(S, engine, orc):
    """LO-MSAC around the six-point estimator, batched on the device (k_sixpt_chain_lo walks, k_sixpt_lo runs the parked
    LocalOptimizations): per pair the same iteration count, LO count, inlier mask and model as the numpy restatement of
    ransac.h:127-276 with SixPointEstimator::NonMinimalSolver / LeastSquares."""
    for seed, N, outl, kw in LO_CASES:
        P = 3
        rays, offsets, f, R, t = S.problems.make_sixpt_batch(seed, P, N, outlier_frac=outl)
        opt = S.default_options(squared_inlier_threshold=4.0, driver=S.DRIVER_LO_MSAC, solver=S.SOLVER_SIXPT_FOCAL,
                                sixpt_focal_scoring=1, random_seed=7, first_pair_id=5, **kw)
        res, flags = engine.estimate_pairs(rays, offsets, opt)
        for p in range(P if not kw else 2):
            pr = rays[offsets[p]:offsets[p + 1]]
            st = _oracle_lo_msac(lambda it: orc.philox_sample(7, 5 + p, it, 6, N), pr, kw, 1, 4.0, 7)
            assert int(res["status"][p]) == st["status"] == 0
            assert int(res["num_iterations"][p]) == st["num_iterations"], (seed, p)
            assert int(res["number_lo_iterations"][p]) == st["number_lo_iterations"], (seed, p)
            assert int(res["best_num_inliers"][p]) == st["best_num_inliers"], (seed, p)
            assert np.nonzero(flags[offsets[p]:offsets[p + 1]])[0].tolist() == st["inliers"].tolist(), (seed, p)
            tt, r, ff = st["model"]
            assert max(np.abs(res["t"][p] - tt).max(), np.abs(res["r"][p] - r).max(), abs(res["focal"][p] - ff) / ff) < 1e-7
            assert abs(res["best_model_score"][p] - st["best_model_score"]) <= 1e-8 * st["best_model_score"]


@pytest.mark.gpu
def test_six_point_lo_msac_improves_on_vanilla_at_scale(S, engine):
    """2 500 pairs (more than one sub-pass of 2 048): every pair finishes, LO-MSAC never ends with a worse MSAC cost than
    the VanillaMSAC run on the same samples would at the same iteration budget, and its refits pull the focal in."""
    P, N = 2500, 400
    rays, offsets, f, R, t = S.problems.make_sixpt_batch(4, P, N, outlier_frac=0.4)
    kw = dict(squared_inlier_threshold=4.0, solver=S.SOLVER_SIXPT_FOCAL, sixpt_focal_scoring=1, random_seed=3,
              min_num_iterations=200, max_num_iterations=200)
    lo, _ = engine.estimate_pairs(rays, offsets, S.default_options(driver=S.DRIVER_LO_MSAC, num_lo_steps=2, num_lsq_iterations=2, **kw))
    va, _ = engine.estimate_pairs(rays, offsets, S.default_options(driver=S.DRIVER_VANILLA_MSAC, **kw))
    assert (lo["status"] == 0).all() and (va["status"] == 0).all()
    assert (lo["num_iterations"] == 200).all() and (lo["number_lo_iterations"] >= 1).all()
    assert (lo["best_model_score"] <= va["best_model_score"] * (1 + 1e-12)).all()
    err_lo = np.abs(lo["focal"] / f - 1)
    err_va = np.abs(va["focal"] / f - 1)
    assert np.median(err_lo) < np.median(err_va)
