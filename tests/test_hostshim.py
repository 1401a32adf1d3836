"""The product's __host__ __device__ arithmetic (csrc/ssfm_math.cuh, ssfm_chain.cuh), compiled for the
host by the TEST-ONLY shim in tests/hostshim, against the oracle.  The two implementations use
different numerics (oracle: Hessenberg-QR eigenvalues + complex null vectors; product: real quadratic
factorisation of the characteristic quartic + Bairstow polish + 2x2 eigenvector solves)."""
import ctypes as C

import numpy as np
import pytest

from conftest import THR2, E_of, match_models, model_dist


def test_philox_sampler_identical(S, orc, shim):
    for seed, pair, it, n in [(0, 0, 0, 1000), (1234, 77, 5, 10), (5, 9, 1000, 5), (0, 3, 2, 4), (9, 9, 9, 3),
                              (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 200000)]:
        a = orc.philox_sample(seed, pair, it, 3, n)
        b = np.zeros(3, np.int32)
        shim.lib.hs_sample(C.c_uint32(seed), C.c_uint32(pair), C.c_uint32(it), 3, n, shim.ip(b))
        c = S.sample(seed, pair, it, 3, n)  # the product library's host entry
        assert (a == b).all() and (a == c).all()


def test_philox_known_answer(orc):
    """Philox4x32-10 known-answer vector (Random123 kat_vectors): ctr=0,key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8.
    Checked through the sampler: word j -> (w*n)>>32 with n = 2^31."""
    n = 1 << 31
    want = [0x6627e8d5 >> 1, 0xe169c58d >> 1, 0xbc57ac4c >> 1]
    assert orc.philox_sample(0, 0, 0, 3, n).tolist() == want


def test_lo_generator_is_libstdcxx_mt19937(orc, shim):
    """The restated mt19937 + uniform_int_distribution (Lemire) consumes exactly the draws of the real
    std::mt19937 / std::uniform_int_distribution used by RansacLib's LO shuffles (utils.h:34-52)."""
    rng = np.random.default_rng(0)
    sizes = np.concatenate([[450, 21, 3, 1000, 2, 1, 7, 0], rng.integers(1, 3000, 40)]).astype(np.int32)
    targets = np.minimum(sizes, 21).astype(np.int32)
    for seed in (0, 42, 0xDEADBEEF):
        a = orc.lo_shuffle(seed, sizes, targets)
        b = np.zeros(int(targets.sum()), np.int32)
        shim.lib.hs_lo_shuffle(C.c_uint32(seed), len(sizes), shim.ip(sizes), shim.ip(targets), shim.ip(b))
        assert (a == b).all()


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_device_solver_matches_oracle(S, orc, shim, kind):
    """north_star: solved models agree within 1e-5 relative after root matching."""
    rng = S.problems.make_rng(3, kind)
    worst = []
    for tr in range(300):
        pr = S.problems.make_problem(rng, 6, bool(tr % 3 == 0), None, 0.0 if tr % 2 == 0 else 1 / 600, 0, 180.0)
        nm, models = shim.solve(pr.rays, [0, 1, 2], kind)
        nmo, mo = orc.solve(pr.rays, [0, 1, 2], kind)
        assert nm == nmo
        worst.append(max(match_models(models[:nm], mo[:nmo]), match_models(mo[:nmo], models[:nm])))
    worst = np.array(worst)
    assert np.median(worst) < 1e-12 and worst.max() < 1e-5


def test_device_quartic(shim):
    rng = np.random.default_rng(1)
    for _ in range(500):
        roots = rng.standard_normal(4) * rng.choice([0.1, 1, 10])
        if rng.random() < 0.5:  # one conjugate pair
            a, b = rng.standard_normal(2)
            c = np.poly(np.array([roots[0], roots[1], a + 1j * b, a - 1j * b])).real
            want = np.array([roots[0], roots[1], a + 1j * b, a - 1j * b])
        else:
            c = np.poly(roots)
            want = roots.astype(complex)
        c = c * rng.uniform(0.5, 2)
        re, im = np.zeros(4), np.zeros(4)
        shim.lib.hs_quartic(shim.dp(np.ascontiguousarray(c)), shim.dp(re), shim.dp(im))
        got = re + 1j * im
        for w in want:
            assert np.min(np.abs(got - w)) < 1e-6 * max(1, abs(w))


def test_device_decompose_and_refit(S, orc, shim):
    worst_d = worst_l = 0.0
    for tr in range(30):
        rng = S.problems.make_rng(11, tr)
        inward = bool(tr % 2)
        pr = S.problems.make_problem(rng, 200, inward, None, 1 / 600, 60, 20.0)
        nm, mo = orc.solve(pr.rays, [0, 1, 2], 0)
        for m in mo:
            E = np.ascontiguousarray(E_of(m).reshape(9))
            r1, t1 = orc.decompose(E, inward)
            r2, t2 = np.zeros(3), np.zeros(3)
            shim.lib.hs_decompose(shim.dp(E), int(inward), shim.dp(r2), shim.dp(t2))
            worst_d = max(worst_d, np.abs(r1 - r2).max(), np.abs(t1 - t2).max())
        inl = np.nonzero(pr.inlier_mask)[0].astype(np.int32)
        r, t = orc.decompose(pr.E / np.linalg.norm(pr.E), inward)
        Ep = orc.make_E(r + 0.01 * rng.standard_normal(3), inward)
        Ea, it, term, costs = orc.lm_refit(pr.rays, inl, Ep, inward)
        Eb = np.ascontiguousarray(Ep.reshape(9)).copy()
        shim.lib.hs_least_squares(shim.dp(pr.rays), shim.ip(inl), len(inl), int(inward), shim.dp(Eb))
        worst_l = max(worst_l, np.abs(Ea.reshape(9) - Eb).max())
    assert worst_d < 1e-9 and worst_l < 1e-8


def test_staged_refit_is_bit_identical(S, shim):
    """The refit kernels gather a refit's correspondences once into a contiguous copy (shared memory / an L1-resident slot)
    and minimise over that copy with an identity sample list: same values, same order, so the refined model must be
    bit-identical to the gathered-on-every-pass version -- also through the small-batch inline path's entry
    (least_squares_as_deferred), whatever the hand-over iteration."""
    for tr in range(12):
        rng = S.problems.make_rng(17, tr)
        inward = bool(tr % 2)
        pr = S.problems.make_problem(rng, 300, inward, None, 1 / 600, 90, 20.0)
        inl = np.nonzero(pr.inlier_mask)[0].astype(np.int32)
        if tr % 3 == 0:
            inl = inl[:21]  # the size of an LO refit
        E0 = np.ascontiguousarray((pr.E / np.linalg.norm(pr.E)).reshape(9))
        E0 = E0 + 1e-3 * rng.standard_normal(9)
        want = E0.copy()
        shim.lib.hs_least_squares(shim.dp(pr.rays), shim.ip(inl), len(inl), int(inward), shim.dp(want))
        for mode, handover in ((1, 0), (2, 0), (2, 3), (2, 24)):
            got = E0.copy()
            shim.lib.hs_least_squares_staged(shim.dp(pr.rays), shim.ip(inl), len(inl), int(inward), shim.dp(got), mode, handover)
            assert got.tobytes() == want.tobytes(), (tr, mode, handover)


def test_required_iterations(shim):
    import math
    for w in (0.0, 1.0, 0.05, 0.3, 0.31234, 0.9, 0.999999, 1e-9):
        got = shim.lib.hs_required_iterations(C.c_double(w), C.c_double(1e-4), 3, 100, 10000)
        if w <= 0:
            want = 10000
        elif w >= 1:
            want = 100
        else:
            miss = 1 - w ** 3
            want = 10000 if miss >= 0.99999999999999 else max(100, min(10000, int(math.ceil(math.log(1e-4) / math.log(miss) + 0.5))))
        assert got == want


CHAIN_CASES = [
    ("pipeline50", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1000, 0.5, False, 12),
    ("pipeline70", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 1500, 0.7, False, 8),
    ("inward", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1, inward=1), 500, 0.4, True, 6),
    ("defaultLO", dict(), 600, 0.5, False, 8),
    ("vanilla", dict(driver=1), 600, 0.5, False, 8),
    ("legacy_fast", dict(driver=2, solver_kind=2, legacy_budget=512), 800, 0.3, False, 8),
    ("poly", dict(solver_kind=1, num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 600, 0.6, False, 6),
    ("tiny8", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 8, 0.25, False, 6),
    ("tiny3", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 3, 0.0, False, 2),
    ("tiny2", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 2, 0.0, False, 2),
    ("hopeless", dict(num_lo_steps=0, num_lsq_iterations=0, final_least_squares=1), 300, 0.92, False, 3),
]


@pytest.mark.parametrize("name,kw,n,outl,inward,trials", CHAIN_CASES)
def test_chain_logic_reproduces_reference_loop(S, O, orc, shim, name, kw, n, outl, inward, trials):
    """The engine's look-ahead rounds + FP32 pre-filter + FP64 certification (ssfm_chain.cuh, run
    serially by the shim) must follow the reference loop exactly: same iteration count, LO count,
    inlier set; same model up to sign."""
    opt = O.default_options(squared_inlier_threshold=THR2, **kw)
    for p in range(trials):
        pr = S.problems.make_problem(S.problems.make_rng(7, p), n, inward, None, 1 / 600, int(outl * n), 20.0)
        a, ia = orc.estimate_pair(pr.rays, opt, p)
        res, fl = shim.estimate_pair(pr.rays, opt, p, defer=p % 2)  # both the inline and the deferred-refit protocol
        fa = np.zeros(n, np.uint8)
        fa[ia] = 1
        assert a.status == res.status
        assert a.num_iterations == res.num_iterations
        assert a.best_num_inliers == res.best_num_inliers
        assert a.number_lo_iterations == res.num_lo
        assert (fa == fl).all()
        if a.status == 0 and n > 3:  # n == 3: every real root fits exactly, the winner is rounding noise
            assert model_dist(np.array(a.E) / np.linalg.norm(a.E), np.array(res.E) / np.linalg.norm(res.E)) < 1e-7
            assert abs(a.best_model_score - res.best_model_score) <= 1e-9 * abs(a.best_model_score)
