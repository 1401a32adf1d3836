"""SfM::Retriangulate (src/sfm.cpp:156-192): per-point LO-MSAC with TriangulationEstimator.
Oracle = oracle/tri_oracle.hpp under the generic lo_msac (pinned bit-exact against the reference's RansacLib headers
below); product = csrc/ssfm_triangulate.cuh (host build without a GPU, k_retriangulate with one)."""
import numpy as np
import pytest


def _opts(O, **kw):
    # Retriangulate's options (sfm.cpp:176-178): RansacLib defaults + threshold 4 px^2 + final least squares
    return O.default_options(squared_inlier_threshold=4.0, final_least_squares=1, **kw)


def test_oracle_triangulation_recovers_points_and_matches_reference_ransaclib(S, O, orc, ref):
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(1, 60, 120, noise_px=0.5, outlier_frac=0.15)
    opt = _opts(O)
    errs = []
    for p in range(120):
        a, b = offs[p], offs[p + 1]
        res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
        if ref is not None:  # same loop through the reference's own include/RansacLib/ransac.h: bit-identical
            r2, inl2 = ref.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
            assert list(res.E) == list(r2.E) and res.num_iterations == r2.num_iterations
            assert res.number_lo_iterations == r2.number_lo_iterations and (inl == inl2).all()
        if res.status == 0:
            errs.append(np.linalg.norm(np.array(res.E[:3]) - X[p]) / np.linalg.norm(X[p]))
    assert len(errs) >= 110 and np.median(errs) < 5e-3
    # noise-free, outlier-free: the point is recovered to rounding
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(2, 40, 20, noise_px=0.0, outlier_frac=0.0)
    for p in range(20):
        a, b = offs[p], offs[p + 1]
        res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
        assert res.status == 0 and res.best_num_inliers == b - a
        assert np.abs(np.array(res.E[:3]) - X[p]).max() < 1e-7 * np.abs(X[p]).max()


def test_restated_estimator_follows_reference_sources(S, O, orc):
    """oracle/tri_oracle.hpp against the reference's own src/triangulation_estimator.cpp + sfm_types.cpp + so3.cpp +
    RansacLib, compiled unmodified against the Eigen/Ceres stand-ins (oracle/_ref/libssfm_reftri.so): same status,
    iteration counts, inlier sets; points to 1e-6 relative.  (The number of LO runs may differ by one: with two
    observations per sample, repeated samples tie up to rounding in `local_best < best_min_score`.)"""
    rt = O.load_ref_tri()
    if rt is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(1, 60, 150, noise_px=0.5, outlier_frac=0.15)
    opt = _opts(O)
    for p in range(150):
        a, b = offs[p], offs[p + 1]
        r1, i1 = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
        r2, i2 = rt.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
        assert (r1.status, r1.num_iterations, r1.best_num_inliers) == (r2.status, r2.num_iterations, r2.best_num_inliers), p
        assert (i1 == i2).all() and abs(r1.number_lo_iterations - r2.number_lo_iterations) <= 1
        assert np.abs(np.array(r1.E[:3]) - np.array(r2.E[:3])).max() <= 1e-6 * max(1.0, np.abs(np.array(r1.E[:3])).max())


def test_product_triangulation_matches_oracle_on_host(S, O, orc, shim):
    """The product's LO-MSAC + estimator (analytic Jacobian, own Jacobi eigen-solver) against the oracle (jets): same
    iteration counts, LO counts, inlier counts; points to 1e-7 relative -- default LO schedule and the no-LO one."""
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(3, 60, 150, noise_px=0.5, outlier_frac=0.2)
    for kw in (dict(), dict(num_lo_steps=0, num_lsq_iterations=0), dict(random_seed=9, lo_starting_iterations=20)):
        opt = _opts(O, **kw)
        for p in range(0, 150, 1 if not kw else 3):
            a, b = offs[p], offs[p + 1]
            res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
            Xh, n, it, nlo = shim.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p)
            if p % 5 == 0:  # the resumable form (the kernel runs the loop in phases) gives the same answer
                Xc, nc, itc, nloc = shim.triangulate(cam[oc[a:b]], oxy[a:b], f, opt, p, chunk=37)
                assert (itc, nc, nloc) == (it, n, nlo) and (Xc == Xh).all()
            assert (it, n, nlo) == (res.num_iterations, res.best_num_inliers, res.number_lo_iterations), (kw, p)
            if res.best_num_inliers >= 3:
                assert np.abs(Xh - np.array(res.E[:3])).max() <= 1e-7 * max(1.0, np.abs(Xh).max())


@pytest.mark.gpu
def test_retriangulate_on_device(S, O, engine, orc):
    """ssfm_retriangulate: all points in one call vs the oracle point by point; ragged tracks including fewer than
    three observations (skipped, point stays at zero) and hopeless tracks (fewer than three inliers)."""
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(5, 80, 400, obs_range=(1, 30), noise_px=0.5, outlier_frac=0.2)
    # make a few tracks hopeless: all observations random
    rng = np.random.default_rng(0)
    for p in (7, 19, 33):
        oxy[offs[p]:offs[p + 1]] = rng.uniform(-300, 300, (offs[p + 1] - offs[p], 2))
    opt = S.default_options(squared_inlier_threshold=4.0, final_least_squares=1, first_pair_id=2)
    pts, ninl, status, iters = engine.retriangulate(cam, offs, oc, oxy, f, opt)
    oopt = _opts(O)
    nok = 0
    for p in range(400):
        a, b = offs[p], offs[p + 1]
        res, inl = orc.triangulate(cam[oc[a:b]], oxy[a:b], f, oopt, 2 + p)
        assert int(status[p]) == res.status, p
        if b - a < 3:
            assert (pts[p] == 0).all()
            continue
        if res.status == 0:
            nok += 1
            assert int(iters[p]) == res.num_iterations and int(ninl[p]) == res.best_num_inliers, p
            assert np.abs(pts[p] - np.array(res.E[:3])).max() <= 1e-7 * max(1.0, np.abs(pts[p]).max())
        else:
            # hopeless tracks (fewer than 3 inliers): every model scores n * thr, the loop is decided by rounding
            # noise; only the outcome is compared
            assert (pts[p] == 0).all() and int(ninl[p]) < 3
    assert nok > 300
    good = status == 0
    rel = np.linalg.norm(pts[good] - X[good], axis=1) / np.linalg.norm(X[good], axis=1)
    assert np.median(rel) < 5e-3


@pytest.mark.gpu
def test_retriangulate_rejects_bad_input(S, engine):
    cam, offs, oc, oxy, f, X = S.problems.make_tracks(6, 10, 5)
    opt = S.default_options(squared_inlier_threshold=4.0, final_least_squares=1)
    oc2 = oc.copy()
    oc2[0] = 99
    with pytest.raises(S.SsfmError):
        engine.retriangulate(cam, offs, oc2, oxy, f, opt)
    pts, ninl, status, iters = engine.retriangulate(cam, offs[:1], oc[:0], oxy[:0], f, opt)
    assert len(pts) == 0
