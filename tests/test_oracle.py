"""The oracle against what pins it: the generator's ground truth, an independent LAPACK (numpy)
solve of the same polynomial system, the reference's own RansacLib (oracle/_ref) and the committed
golden vectors (tests/golden, produced by oracle/_ref)."""
import os

import numpy as np
import pytest

from conftest import THR2, E_of, check_full_path_goldens, match_models, model_dist


def numpy_solutions(rays3):
    """Independent restatement by linear algebra only: null space by SVD, the nine cubic constraints
    det(E)=0 and 2EE^TE - tr(EE^T)E = 0 sampled on E(x,y,1) and fitted by least squares, then solved by
    the hidden-variable resultant in y via numpy.roots on a dense polynomial fit."""
    u, v = rays3[:, :3], rays3[:, 3:]
    A = np.stack([u[:, 0] * v[:, 0] - u[:, 1] * v[:, 1], u[:, 0] * v[:, 1] + u[:, 1] * v[:, 0], u[:, 2] * v[:, 0],
                  u[:, 2] * v[:, 1], u[:, 0] * v[:, 2], u[:, 1] * v[:, 2]], axis=1)
    _, _, Vt = np.linalg.svd(A)
    B = Vt[3:].T  # 6x3

    def cons(b):
        E = E_of(B @ b)
        T = 2 * E @ E.T @ E - np.trace(E @ E.T) * E
        return np.concatenate([T.ravel(), [np.linalg.det(E)]])

    # Newton from many starts on the 10 constraints in (x, y) with z = 1, plus (x, 1, 0)-type charts skipped
    sols = []
    rng = np.random.default_rng(0)
    for _ in range(200):
        b = np.array([*rng.standard_normal(2) * 3, 1.0])
        for _ in range(60):
            f = cons(b)
            J = np.zeros((10, 2))
            for k in range(2):
                d = np.zeros(3)
                d[k] = 1e-6
                J[:, k] = (cons(b + d) - cons(b - d)) / 2e-6
            step = np.linalg.lstsq(J, -f, rcond=None)[0]
            b[:2] += step
            if np.linalg.norm(step) < 1e-13:
                break
        if np.linalg.norm(cons(b)) < 1e-9 * max(1, np.linalg.norm(b) ** 3):
            p = B @ b
            E = E_of(p)
            p = p / np.linalg.norm(E)
            if not any(model_dist(p, q) < 1e-6 for q in sols):
                sols.append(p)
    return sols


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_solver_recovers_ground_truth(S, orc, kind):
    """evaluation/test_random_problems.cpp:94-132 + run_stability_experiment.py: noise-free best-of-4
    Frobenius error is at machine precision."""
    rng = S.problems.make_rng(1, kind)
    errs = []
    for _ in range(300):
        pr = S.problems.make_problem(rng, 6, bool(_ % 2), None, 0.0, 0, 180.0)
        nm, models = orc.solve(pr.rays, [0, 1, 2], kind)
        errs.append(min([S.problems.frob_error(pr.E, E_of(m)) for m in models[:nm]] + [9.0]))
    errs = np.array(errs)
    assert np.median(errs) < 1e-13
    if kind == 2:  # the Sturm variant only brackets y in [-10, 10] (spherical_fast_estimator.cpp:223)
        assert (errs > 1e-8).mean() < 0.12
    else:
        assert errs.max() < 1e-8


def test_solver_matches_independent_numpy_solve(S, orc):
    rng = S.problems.make_rng(2, 0)
    for _ in range(6):
        pr = S.problems.make_problem(rng, 3, False, None, 0.0, 0, 60.0)
        real = numpy_solutions(pr.rays)
        nm, models = orc.solve(pr.rays, [0, 1, 2], 0)
        assert len(real) >= 1
        # every real solution found by brute-force Newton is among the oracle's models
        for p in real:
            assert min(model_dist(p, m) for m in models) < 1e-7
        # the polynomial and action-matrix variants agree on the models of real roots (those of complex roots differ by
        # construction upstream: Re(y) of the Ferrari root vs the real part of a complex eigenvector)
        _, mp = orc.solve(pr.rays, [0, 1, 2], 1)
        _, mreal = orc.solve(pr.rays, [0, 1, 2], 0, 2)  # COMPLEX_SKIP: real eigenvalues only
        assert match_models(mreal, mp) < 1e-7


def test_models_satisfy_constraints_and_sample(S, orc):
    rng = S.problems.make_rng(3, 0)
    pr = S.problems.make_problem(rng, 50, False, None, 1 / 600, 10, 30.0)
    for it in range(20):
        s = orc.philox_sample(0, 0, it, 3, 50)
        assert len(set(s.tolist())) == 3 and s.min() >= 0 and s.max() < 50
        nm, models = orc.solve(pr.rays, s, 0)
        assert nm == 4
        for m in models:
            E = E_of(m)
            assert abs(np.linalg.norm(E) - 1) < 1e-12
        # at least the real roots vanish on the sample
        best = min(orc.sampson(E_of(m), pr.rays[s]).max() for m in models)
        assert best < 1e-20


def test_sampson_matches_numpy(S, orc):
    rng = S.problems.make_rng(4, 0)
    pr = S.problems.make_problem(rng, 200, False, None, 1 / 600, 50, 30.0)
    E = pr.E / np.linalg.norm(pr.E)
    u, v = pr.rays[:, :3], pr.rays[:, 3:]
    Eu = u @ E.T
    Etv = v @ E
    d = (v * Eu).sum(1)
    want = d * d / (Eu[:, 0] ** 2 + Eu[:, 1] ** 2 + Etv[:, 0] ** 2 + Etv[:, 1] ** 2)
    got = orc.sampson(E, pr.rays)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-22)  # d = v.Eu cancels for inliers
    s, n = orc.score(E, pr.rays, THR2)
    assert n == (want < THR2).sum() and np.isclose(s, np.minimum(want, THR2).sum(), rtol=1e-10)


@pytest.mark.parametrize("inward", [False, True])
def test_decompose_round_trip(S, orc, inward):
    rng = np.random.default_rng(5)
    for _ in range(50):
        r = rng.standard_normal(3)
        r *= rng.uniform(0.01, 3.0) / np.linalg.norm(r)
        E = orc.make_E(r, inward)
        r2, t2 = orc.decompose(E, inward)
        R, R2 = S.problems.so3exp(r), S.problems.so3exp(r2)
        assert S.problems.rot_error(R, R2) < 1e-7
        Ew, tw = S.problems.make_spherical_E(R, inward)
        assert np.allclose(Ew, E, atol=1e-12) and np.allclose(t2, tw, atol=1e-7)


def test_lm_refit_improves_and_converges(S, orc):
    rng = S.problems.make_rng(6, 0)
    for tr in range(10):
        pr = S.problems.make_problem(rng, 100, False, None, 1 / 600, 0, 20.0)
        r, t = orc.decompose(pr.E)
        E0 = orc.make_E(r + 0.01 * rng.standard_normal(3))
        E1, iters, term, costs = orc.lm_refit(pr.rays, np.arange(100), E0)
        assert costs[1] < costs[0] and 1 <= iters <= 200 and term in (1, 2, 3)
        assert S.problems.frob_error(pr.E, E1) < S.problems.frob_error(pr.E, E0)
        assert S.problems.frob_error(pr.E, E1) < 5e-3


CASES = [
    ("pipeline50", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1000, 0.5),
    ("pipeline70", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 600, 0.7),
    ("defaultLO", dict(), 400, 0.5),
    ("vanilla", dict(driver=1), 400, 0.5),
    ("poly", dict(solver_kind=1, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 300, 0.4),
]


@pytest.mark.parametrize("name,kw,n,outl", CASES)
def test_restated_drivers_equal_reference_ransaclib(S, O, orc, ref, name, kw, n, outl):
    """The restated LO-MSAC / VanillaMSAC loops (oracle/lomsac.hpp) against the reference's own
    include/RansacLib/ransac.h and evaluation/vanilla_ransac.h compiled in oracle/_ref: bit-identical."""
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    opt = O.default_options(squared_inlier_threshold=THR2, **kw)
    for p in range(8):
        pr = S.problems.make_problem(S.problems.make_rng(7, p), n, False, None, 1 / 600, int(outl * n), 20.0)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        b, ib = ref.estimate_pair(pr.rays, opt, p)
        assert list(a.E) == list(b.E)
        assert (a.num_iterations, a.best_num_inliers, a.number_lo_iterations, a.best_model_score, a.inlier_ratio) == (
            b.num_iterations, b.best_num_inliers, b.number_lo_iterations, b.best_model_score, b.inlier_ratio)
        assert (ia == ib).all() and a.evals == b.evals


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_restated_preemptive_ransac_equals_reference_header(S, O, orc, ref, kind):
    """oracle/lomsac.hpp::preemptive_ransac against the reference's own include/sphericalsfm/preemptive_ransac.h
    compiled in oracle/_ref (its rand() call redirected to the same Philox-backed stream): bit-identical winner,
    inlier mask and count, over ragged sizes, hypothesis budgets and block sizes."""
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for p in range(18):
        n = [4, 5, 30, 200, 1000, 1500][p % 6]
        pr = S.problems.make_problem(S.problems.make_rng(11, p), n, p % 2 == 1, None, 1 / 600, int(0.4 * n), 20.0)
        opt = O.default_options(squared_inlier_threshold=THR2, driver=3, solver_kind=kind, legacy_budget=[64, 500, 512][p % 3],
                                preemptive_block=[10, 7, 3][p % 3], random_seed=5, inward=p % 2)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        b, ib = ref.estimate_pair(pr.rays, opt, p)
        assert a.status == b.status and list(a.E) == list(b.E)
        assert (a.num_iterations, a.best_num_inliers, a.best_model_score, a.inlier_ratio, a.evals) == (
            b.num_iterations, b.best_num_inliers, b.best_model_score, b.inlier_ratio, b.evals)
        assert (ia == ib).all()


def test_preemptive_ransac_recovers_pose(S, O, orc):
    """The pre-emptive driver with the Sturm solver (its upstream pairing) finds the rotation on C2-like pairs."""
    opt = O.default_options(squared_inlier_threshold=THR2, driver=3, solver_kind=2, legacy_budget=512, preemptive_block=10)
    ok = 0
    for p in range(10):
        pr = S.problems.make_problem(S.problems.make_rng(12, p), 2000, False, None, 1 / 600, 600, 1.0)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        assert a.status == 0
        ok += np.rad2deg(S.problems.rot_error(pr.R, S.problems.so3exp(np.array(a.r)))) < 0.5
    assert ok >= 9


def test_selection_sample_properties(S, orc):
    """random_sample (Knuth 3.4.2S): k distinct indices in increasing order, uniform marginals; the product's host
    hook reproduces the oracle's stream."""
    hits = np.zeros(50)
    for h in range(4000):
        idx = orc.knuth_sample(3, 9, h, 50, 4)
        assert (np.diff(idx) > 0).all() and idx[0] >= 0 and idx[-1] < 50
        hits[idx] += 1
    assert abs(hits / 4000 - 4 / 50).max() < 0.02
    for h in range(200):
        assert (S.selection_sample(3, 9, h, 977, 4) == orc.knuth_sample(3, 9, h, 977, 4)).all()
    assert (orc.knuth_sample(1, 1, 0, 4, 4) == np.arange(4)).all()


def test_restated_sturm_solver_matches_reference_fast_estimator_source(S, O, orc):
    """The FAST_STURM solver of the oracle (config C2's solver) against the reference's own orphan
    src/spherical_fast_estimator.cpp, compiled unmodified with stand-ins for the old Estimator base and
    Polynomial<4>::realRootsSturm (oracle/_ref/libssfm_reffast.so): same number of solutions, same matrices (up to
    sign) to 1e-8; score() and decomposeE() identical to the oracle's Sampson error and decomposition."""
    rf = O.load_ref_fast()
    if rf is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    total = 0
    for k in range(200):
        pr = S.problems.make_problem(S.problems.make_rng(21, k), 30, k % 2 == 1, None, (1 / 600 if k % 3 else 0.0), 0,
                                     20.0 if k % 4 else 180.0)
        smp = np.array([1, 7, 13], np.int32)
        nm, models = orc.solve(pr.rays, smp, 2)
        Er = rf.compute(pr.rays[smp])
        assert len(Er) == nm, k
        total += nm
        for i in range(nm):
            Eo = E_of(models[i])
            Eo = Eo / np.linalg.norm(Eo)
            assert min(min(np.abs(Eo - e).max(), np.abs(Eo + e).max()) for e in Er) < 1e-8
    assert total > 300
    pr = S.problems.make_problem(S.problems.make_rng(22, 0), 50, False, None, 1 / 600, 10, 20.0)
    E = pr.E / np.linalg.norm(pr.E)
    assert np.abs(rf.score(E, pr.rays) - orc.sampson(E, pr.rays)).max() == 0.0
    for inward in (False, True):
        r1, t1 = rf.decompose(E, inward)
        r2, t2 = orc.decompose(E, inward)
        assert np.abs(r1 - r2).max() < 1e-12 and np.abs(t1 - t2).max() < 1e-12


@pytest.mark.parametrize("driver", [2, 3])
def test_restated_legacy_path_matches_reference_sources_end_to_end(S, O, orc, driver):
    """Config C2 as upstream wrote it -- include/sphericalsfm/msac.h (driver 2) / preemptive_ransac.h (driver 3) around
    src/spherical_fast_estimator.cpp, all compiled unmodified (oracle/_ref/libssfm_reflegacy.so; rand() redirected to
    the Philox-backed stream) -- against the restated drivers + restated Sturm solver: same status, iteration counts
    and inlier sets; the winning matrix to 1e-7 (up to sign)."""
    rl = O.load_ref_legacy()
    if rl is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for p in range(18):
        n = [2000, 300, 60][p % 3]
        pr = S.problems.make_problem(S.problems.make_rng(31, p), n, False, 1.0, 1 / 600, int(0.3 * n), 20.0)
        opt = O.default_options(squared_inlier_threshold=THR2, driver=driver, solver_kind=2, legacy_budget=[512, 200][p % 2],
                                preemptive_block=[10, 5][p % 2], random_seed=3)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        b, ib = rl.estimate_pair(pr.rays, opt, p)
        assert (a.status, a.num_iterations, a.best_num_inliers) == (b.status, b.num_iterations, b.best_num_inliers), p
        assert (ia == ib).all()
        if a.status == 0:
            Ea, Eb = np.array(a.E) / np.linalg.norm(a.E), np.array(b.E) / np.linalg.norm(b.E)
            assert min(np.abs(Ea - Eb).max(), np.abs(Ea + Eb).max()) < 1e-7
            assert abs(a.best_model_score - b.best_model_score) <= 1e-6 * b.best_model_score  # models differ by ~1e-9


def test_golden_vectors_from_reference_sources(S, O, orc):
    """tests/golden/refsrc_golden.npz (made by make_golden_refsrc.py from the reference's own sources compiled into
    oracle/_ref) against the restated oracle: legacy drivers + Sturm solver (config C2), Retriangulate, and the
    3-point pipeline path (estimator + solvers + RansacLib)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "refsrc_golden.npz"))
    for k in range(int(g["num_legacy"])):
        drv, M, B, pid, seed = [int(x) for x in g["lg_cfg_%d" % k]]
        opt = O.default_options(squared_inlier_threshold=THR2, driver=drv, solver_kind=2, legacy_budget=M, preemptive_block=B,
                                random_seed=seed)
        res, inl = orc.estimate_pair(g["lg_rays_%d" % k], opt, pid)
        assert res.num_iterations == int(g["lg_iters_%d" % k]) and res.best_num_inliers == int(g["lg_ninl_%d" % k])
        assert (inl == g["lg_inliers_%d" % k]).all()
        Ea, Eb = np.array(res.E) / np.linalg.norm(res.E), g["lg_E_%d" % k] / np.linalg.norm(g["lg_E_%d" % k])
        assert min(np.abs(Ea - Eb).max(), np.abs(Ea + Eb).max()) < 1e-7
    opt = O.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
    cam, offs, oc, oxy, f = g["tri_cam"], g["tri_offs"], g["tri_oc"], g["tri_oxy"], float(g["tri_focal"])
    for p in range(len(offs) - 1):
        a, b = offs[p], offs[p + 1]
        res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
        assert (res.status, res.num_iterations, res.best_num_inliers) == (int(g["tri_status"][p]), int(g["tri_iters"][p]), int(g["tri_ninl"][p]))
        assert np.abs(np.array(res.E[:3]) - g["tri_points"][p]).max() <= 1e-6 * max(1.0, np.abs(g["tri_points"][p]).max())
    check_full_path_goldens(S, g, lambda rays, cfg, skip: _oracle_full_case(O, orc, rays, cfg, skip))


def _oracle_full_case(O, orc, rays, cfg, skip):
    n, nout, pid, inward, flsq, kind = cfg
    opt = O.default_options(squared_inlier_threshold=THR2, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=flsq,
                            inward=inward, solver_kind=kind, complex_mode=O.COMPLEX_SKIP if skip else O.COMPLEX_CANONICAL)
    res, inl = orc.estimate_pair(rays, opt, pid)
    f = np.zeros(n, np.uint8)
    f[inl] = 1
    return (res.status, res.num_iterations, res.best_num_inliers, res.number_lo_iterations), np.array(res.r), np.array(res.E), f


def test_reference_generator_problems_and_metrics(S, O, orc):
    """Problems drawn by the reference's own ProblemGenerator (tests/golden/refgen_golden.npz) through the restated
    oracle: evaluation/test_random_problems.cpp's stability check (minimal solve on {0,1,2}, best-of-solutions Frobenius
    error at machine precision without noise), evaluation/test_ransac.cpp's VanillaMSAC run and the pipeline's LO-MSAC --
    against what the reference's own estimator sources produced.  Where oracle/_ref exists, the Python error metrics
    are also checked against RelativePoseSolution::calc_*_error."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "refgen_golden.npz"))
    gen = O.load_ref_gen()
    for k in range(int(g["num_cases"])):
        rays, E, R, t = g["rays_%d" % k], g["E_%d" % k], g["R_%d" % k], g["t_%d" % k]
        inward, noise = bool(g["cfg_%d" % k][0]), float(g["cfg_%d" % k][1])
        nm, models = orc.solve(rays, np.array([0, 1, 2], np.int32), 0)
        gm = g["min_models_%d" % k]
        assert nm == len(gm)
        best = min(S.problems.frob_error(E, E_of(m)) for m in models[:nm] if np.isfinite(m).all())
        if noise == 0.0:
            assert best < 1e-9  # evaluation/scripts/run_stability_experiment.py: ~machine precision
            real = [m for m in gm if np.isfinite(m).all() and S.problems.frob_error(E, E_of(m)) < 1e-6]
            assert real and min(model_dist(E_of(real[0]) / np.linalg.norm(E_of(real[0])), E_of(m) / np.linalg.norm(E_of(m)))
                                for m in models[:nm]) < 1e-8
        for name, kw in (("van", dict(driver=1, max_num_iterations=2 ** 31 - 1)),
                         ("lo", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1))):
            opt = O.default_options(squared_inlier_threshold=THR2, inward=int(inward), **kw)
            res, inl = orc.estimate_pair(rays, opt, 100 + k)
            st = g["%s_stats_%d" % (name, k)]
            # on (nearly) noise-free data most minimal models tie at a cost of ~0, so `local_best < best_min_score`
            # (and with it the number of LO runs) is decided by rounding noise between implementations
            assert (res.num_iterations, res.best_num_inliers) == (int(st[0]), int(st[1])), (name, k)
            assert noise == 0.0 or abs(res.number_lo_iterations - int(st[2])) <= 1
            assert (inl == g["%s_inliers_%d" % (name, k)]).all()
            Eg = g["%s_E_%d" % (name, k)]
            assert model_dist(np.array(res.E).reshape(3, 3) / np.linalg.norm(res.E), Eg.reshape(3, 3) / np.linalg.norm(Eg)) < 1e-6
        if gen is not None:
            rng = np.random.default_rng(k)
            Es, Rs, ts = E + rng.normal(size=(3, 3)) * 0.01, S.problems.so3exp(rng.normal(size=3) * 0.1) @ R, t + rng.normal(size=3) * 0.05
            ref_err = gen.errors(E, R, t, Es, Rs, ts)
            assert abs(ref_err[0] - S.problems.frob_error(E, Es)) < 1e-12 and abs(ref_err[1] - S.problems.rot_error(R, Rs)) < 1e-9


def test_oracle_recovers_pose_with_outliers(S, O, orc):
    """Config C1: 1000 correspondences, 50 % outliers, calibrated solver, pipeline options."""
    opt = O.pipeline_options(THR2)
    for p in range(5):
        pr = S.problems.make_problem(S.problems.make_rng(8, p), 1000, False, None, 1 / 600, 500, 20.0)
        res, inl = orc.estimate_pair(pr.rays, opt, p)
        assert res.status == 0 and res.best_num_inliers > 400
        assert np.rad2deg(S.problems.rot_error(pr.R, S.problems.so3exp(np.array(res.r)))) < 0.2
        assert pr.inlier_mask[inl].mean() > 0.97


def test_edge_cases(S, O, orc):
    opt = O.pipeline_options(THR2)
    pr = S.problems.make_problem(S.problems.make_rng(9, 0), 10, False, None, 0.0, 0, 20.0)
    res, inl = orc.estimate_pair(pr.rays[:2], opt, 0)  # fewer points than the minimal sample
    assert res.status == 1 and res.best_num_inliers == 0 and res.num_iterations == 0
    res, inl = orc.estimate_pair(pr.rays[:3], opt, 0)  # exactly minimal
    assert res.status == 0 and res.best_num_inliers == 3
    res, inl = orc.estimate_pair(pr.rays, opt, 0)
    assert res.best_num_inliers == 10 and res.num_iterations == 100  # clamped to min_num_iterations


def test_golden_vectors(S, O, orc):
    """tests/golden/lomsac_golden.npz was produced by oracle/_ref (reference RansacLib driver) with
    tests/golden/make_golden.py; the source-only restatement must reproduce it wherever it runs."""
    path = os.path.join(os.path.dirname(__file__), "golden", "lomsac_golden.npz")
    g = np.load(path)
    for k in range(int(g["num_cases"])):
        opt = O.default_options()
        for name, val in zip(g["opt_names_%d" % k], g["opt_vals_%d" % k]):
            cur = getattr(opt, str(name))
            setattr(opt, str(name), type(cur)(val))
        rays = g["rays_%d" % k]
        res, inl = orc.estimate_pair(rays, opt, int(g["pair_id_%d" % k]))
        assert np.array_equal(np.array(res.E), g["E_%d" % k])
        assert res.num_iterations == g["num_iterations_%d" % k]
        assert res.best_num_inliers == g["best_num_inliers_%d" % k]
        assert res.number_lo_iterations == g["number_lo_iterations_%d" % k]
        assert res.best_model_score == g["best_model_score_%d" % k]
        assert np.array_equal(inl, g["inliers_%d" % k])


# ---------------------------------------------------------------------------------------------------
# The restatement against the reference's OWN sources (oracle/_ref/libssfm_reffull.so: RansacLib,
# src/spherical_estimator.cpp, src/spherical_solvers.cpp, src/so3.cpp, src/spherical_utils.cpp compiled
# unmodified against stand-ins for Eigen and Ceres).
# ---------------------------------------------------------------------------------------------------
def _action_matrix(S, rays3):
    """The 4x4 action matrix M of a minimal sample (src/spherical_solvers.cpp:281-285), built independently in numpy:
    null space by SVD, the six cubic constraints by fitting, G by a linear solve."""
    u, v = rays3[:, :3], rays3[:, 3:]
    A = np.stack([u[:, 0] * v[:, 0] - u[:, 1] * v[:, 1], u[:, 0] * v[:, 1] + u[:, 1] * v[:, 0], u[:, 2] * v[:, 0],
                  u[:, 2] * v[:, 1], u[:, 0] * v[:, 2], u[:, 1] * v[:, 2]], 1)
    B = np.linalg.svd(A)[2][3:].T

    def T(b):
        p = B @ b
        E = E_of(p)
        M3 = 2 * E @ E.T @ E - np.trace(E @ E.T) * E
        return np.array([M3[1, 0], M3[2, 0], M3[0, 0], M3[2, 1], M3[1, 2], M3[2, 2]])
    mon = lambda b: np.array([b[0] ** 3, b[0] ** 2 * b[1], b[0] * b[1] ** 2, b[1] ** 3, b[0] ** 2 * b[2], b[0] * b[1] * b[2],
                              b[1] ** 2 * b[2], b[0] * b[2] ** 2, b[1] * b[2] ** 2, b[2] ** 3])
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((40, 3))
    Cm = np.linalg.lstsq(np.array([mon(b) for b in pts]), np.array([T(b) for b in pts]), rcond=None)[0].T
    G = np.linalg.solve(Cm[:, :6], Cm[:, 6:])
    M = np.zeros((4, 4))
    M[0], M[1], M[2] = -G[2], -G[4], -G[5]
    M[3, 1] = 1
    return M


def test_eigen_restatement_is_an_eigendecomposition(S, orc):
    """oracle's Eigen::EigenSolver<Matrix4d> restatement against LAPACK: eigenvalues, M v = lambda v, unit columns,
    conjugate pairs stored as (re + i im, re - i im) from one real column pair."""
    rng = np.random.default_rng(5)
    for tr in range(300):
        M = rng.standard_normal((4, 4)) * rng.choice([1e-3, 1.0, 1e3])
        if tr % 3 == 0:
            M[3] = [0, 1, 0, 0]
        ev, V, ok = orc.eigen34(M)
        assert ok
        want = np.linalg.eigvals(M)
        for lam in ev:
            assert np.min(np.abs(want - lam)) < 1e-9 * np.abs(want).max()
        for k in range(4):
            assert abs(np.linalg.norm(V[:, k]) - 1) < 1e-12
            assert np.abs(M @ V[:, k] - ev[k] * V[:, k]).max() < 1e-9 * np.abs(M).max()
            if ev[k].imag > 0:
                assert ev[k + 1] == np.conj(ev[k]) and (V[:, k + 1] == np.conj(V[:, k])).all()


def test_complex_root_models_are_ill_conditioned_upstream(S, orc):
    """WHY models from complex eigenvalues of the action matrix cannot be pinned.  The reference returns
    Re(eigenvector) (src/spherical_solvers.cpp:294-297); the phase of Eigen's complex eigenvector is fixed by the last
    Francis sweeps, which act on a converged (rounding-noise sized) sub-diagonal.  Measured on Eigen 3.4's algorithm
    restated: perturbing the action matrix by ONE ulp leaves real eigenvectors where they were (< 1e-9) and moves the real
    part of complex ones by more than 1e-6 in over a quarter of the cases, by more than 1e-2 in some.  So two builds of
    the reference itself (different compiler, FMA contraction, Eigen vectorisation) disagree on these models."""
    moved_real, moved_cplx = [], []
    for tr in range(400):
        pr = S.problems.make_problem(S.problems.make_rng(33, tr), 6, False, None, 1 / 600, 0, 180.0)
        M = _action_matrix(S, pr.rays[:3])
        ev, V, ok = orc.eigen34(M)
        rng = np.random.default_rng(tr)
        M2 = M * (1 + rng.integers(-1, 2, (4, 4)) * 2.0 ** -52)
        ev2, V2, ok2 = orc.eigen34(M2)
        assert ok and ok2
        for k in range(4):
            a, b = V[1:, k].real, V2[1:, k].real
            d = min(np.abs(a - b).max(), np.abs(a + b).max())
            (moved_cplx if ev[k].imag != 0 else moved_real).append(d)
    moved_real, moved_cplx = np.array(moved_real), np.array(moved_cplx)
    print("1-ulp perturbation: real eigenvectors move by max %.1e (n=%d); Re(complex eigenvector) by median %.1e, "
          "%.0f %% > 1e-6, %.0f %% > 1e-2 (n=%d)" % (moved_real.max(), len(moved_real), np.median(moved_cplx),
                                                     100 * (moved_cplx > 1e-6).mean(), 100 * (moved_cplx > 1e-2).mean(), len(moved_cplx)))
    assert len(moved_cplx) > 200 and len(moved_real) > 200
    assert np.percentile(moved_real, 99) < 1e-9
    assert (moved_cplx > 1e-6).mean() > 0.25 and (moved_cplx > 1e-2).mean() > 0.02


@pytest.mark.parametrize("kind", [0, 1])
def test_restated_solvers_match_reference_sources(S, O, orc, reffull, kind):
    """Every model of every sample against src/spherical_solvers.cpp compiled in oracle/_ref.  Polynomial variant: all four
    models, the Re(y) ones of complex Ferrari roots included (deterministic upstream, :73-83, :631-657).  Action matrix: all
    models of real eigenvalues; with the complex ones masked on both sides the model sets are identical, NaN pattern included."""
    if reffull is None:
        pytest.skip("oracle/_ref not built")
    rng = S.problems.make_rng(21, kind)
    worst, ncomplex = [], 0
    for tr in range(300):
        pr = S.problems.make_problem(rng, 8, bool(tr % 3 == 0), None, 0.0 if tr % 2 == 0 else 1 / 600, 0, 180.0)
        mode = O.COMPLEX_SKIP if kind == 0 else O.COMPLEX_CANONICAL
        nmo, mo = orc.solve(pr.rays, [0, 1, 2], kind, mode)
        nmr, mr = reffull.solve(pr.rays, [0, 1, 2], kind, mode)
        assert nmo == nmr == 4
        no, nr = np.isnan(mo).any(axis=1), np.isnan(mr).any(axis=1)
        assert no.sum() == nr.sum()
        ncomplex += int(no.sum())
        if (~no).any():
            worst.append(max(match_models(mo, mr), match_models(mr, mo)))
    worst = np.array(worst)
    assert kind == 1 or ncomplex > 100
    assert np.median(worst) < 1e-13
    assert worst.max() < (1e-8 if kind == 0 else 1e-5)  # the reference's Ferrari quartic is the less accurate one


def test_restated_scoring_refit_geometry_match_reference_sources(S, orc, reffull):
    if reffull is None:
        pytest.skip("oracle/_ref not built")
    rng = S.problems.make_rng(22, 0)
    for tr in range(10):
        inward = bool(tr % 2)
        pr = S.problems.make_problem(rng, 120, inward, None, 1 / 600, 30, 20.0)
        E = pr.E / np.linalg.norm(pr.E)
        assert np.array_equal(orc.sampson(E, pr.rays), reffull.sampson(E, pr.rays))  # EvaluateModelOnPoint: bit-exact
        assert orc.score(E, pr.rays, THR2) == reffull.score(E, pr.rays, THR2)
        ro, to = orc.decompose(E, inward)
        rr, tr_ = reffull.decompose(E, inward)
        assert np.abs(ro - rr).max() < 1e-12 and np.abs(to - tr_).max() < 1e-12
        assert np.abs(orc.make_E(ro, inward) - reffull.make_E(ro, inward)).max() < 1e-15
        # SphericalEstimator::LeastSquares: the reference's autodiff'd SampsonError through the Ceres stand-in
        inl = np.nonzero(pr.inlier_mask)[0].astype(np.int32)
        E0 = orc.make_E(ro + 0.01 * rng.standard_normal(3), inward)
        Ea, it, term, costs = orc.lm_refit(pr.rays, inl[:21], E0, inward)
        Eb, _, _, _ = reffull.lm_refit(pr.rays, inl[:21], E0, inward)
        assert np.abs(Ea - Eb).max() < 1e-9


REF_CASES = [
    ("pipeline50", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1000, 0.5, 24),
    ("pipeline70", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1500, 0.7, 16),
    ("loopclosure70", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=0), 1500, 0.7, 8),
    ("defaultLO", dict(), 400, 0.5, 4),
    ("vanilla", dict(driver=1), 400, 0.4, 6),
    ("poly", dict(solver_kind=1, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 400, 0.5, 8),
]


def _same_trajectory(a, ia, b, ib):
    return (a.status == b.status and a.num_iterations == b.num_iterations and a.best_num_inliers == b.best_num_inliers and
            a.number_lo_iterations == b.number_lo_iterations and len(ia) == len(ib) and bool((ia == ib).all()))


@pytest.mark.parametrize("name,kw,n,outl,trials", REF_CASES)
def test_restatement_follows_reference_sources_end_to_end(S, O, orc, reffull, name, kw, n, outl, trials):
    """Whole-path agreement with the reference's sources, no tolerance on the trajectory: with the models of complex
    action-matrix eigenvalues masked on both sides (upstream's own commented-out filter; the only ill-posed piece, see
    test_complex_root_models_are_ill_conditioned_upstream) EVERY pair gives the same iteration count, LO count and inlier
    set, E to ~1e-15 and the pose within 0.01 deg.  Unmasked, a pair can only differ where such a model wins an iteration;
    the fraction that does is printed (and pinned case by case in tests/golden/refsrc_golden.npz)."""
    if reffull is None:
        pytest.skip("oracle/_ref not built")
    followed = 0
    for p in range(trials):
        pr = S.problems.make_problem(S.problems.make_rng(7, p), n, False, None, 1 / 600, int(outl * n), 20.0)
        opt = O.default_options(squared_inlier_threshold=THR2, complex_mode=O.COMPLEX_SKIP, **kw)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        b, ib = reffull.estimate_pair(pr.rays, opt, p)
        assert a.status == b.status == 0
        assert _same_trajectory(a, ia, b, ib), (name, p)
        # LM stops on Ceres' 1e-6 function tolerance, so long refit chains (default LO: 51 LMs per LO)
        # amplify rounding differences between the two minimiser implementations to ~1e-7
        assert model_dist(np.array(a.E) / np.linalg.norm(a.E), np.array(b.E) / np.linalg.norm(b.E)) < 1e-5
        assert abs(a.best_model_score - b.best_model_score) <= 1e-6 * a.best_model_score
        d = S.problems.rot_error(S.problems.so3exp(np.array(a.r)), S.problems.so3exp(np.array(b.r)))
        assert np.rad2deg(d) < 0.01
        if kw.get("solver_kind", 0) == 0:
            opt.complex_mode = O.COMPLEX_CANONICAL
            a, ia = orc.estimate_pair(pr.rays, opt, p)
            opt.complex_mode = O.COMPLEX_EIGEN
            b, ib = reffull.estimate_pair(pr.rays, opt, p)
            followed += _same_trajectory(a, ia, b, ib)
        else:
            followed += 1
    print("%s: canonical-representative oracle follows upstream-as-written on %d / %d pairs" % (name, followed, trials))


def test_lm_refit_termination_is_sensitive_upstream(S, O, orc):
    """tests/golden/lm_sensitive_pair.npz: a C1-shaped pair found by the at-scale GPU parity run (pair 1132 of 2048) on which the
    engine, the host build of the engine's code and the oracle ended with 480 / 481 / 482 inliers -- same iterations, same LO
    count.  The oracle alone reproduces all three outcomes when its INPUT rays are changed by one ulp: the final least-squares
    refit (Ceres trust region, function tolerance 1e-6, on a cost whose residual is the squared Sampson value, so the minimum
    is degenerate) stops one step earlier or later.  The poses stay within 0.01 deg of each other, which is the bar that
    applies to refitted models; trajectories of such pairs are not reproducible by any second implementation."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lm_sensitive_pair.npz"))
    rays, pid = g["rays"], int(g["pair_id"])
    opt = O.pipeline_options(THR2)
    base, _ = orc.estimate_pair(rays, opt, pid)
    outcomes, worst = set(), 0.0
    for j in range(1, 9):
        rng = np.random.default_rng(j)
        m = np.ones_like(rays)
        m[:, [0, 1, 3, 4]] = 1 + rng.integers(-1, 2, (len(rays), 4)) * 2.0 ** -52
        a, _ = orc.estimate_pair(rays * m, opt, pid)
        assert (a.num_iterations, a.number_lo_iterations) == (base.num_iterations, base.number_lo_iterations)
        outcomes.add(a.best_num_inliers)
        worst = max(worst, np.rad2deg(S.problems.rot_error(S.problems.so3exp(np.array(base.r)), S.problems.so3exp(np.array(a.r)))))
    print("1-ulp input perturbations: inlier counts", sorted(outcomes), "worst pose difference %.4f deg" % worst)
    assert len(outcomes) >= 2 and worst < 0.01
