"""CPU ORACLE (test infrastructure only) for the six-point shared-focal estimator.

Follows examples/six_point_estimator.{h,cpp} of the reference:
  MinimalSolver          six_point_estimator.cpp:93-119  (poselib::relpose_6pt_shared_focal, :104)
  EvaluateModelOnPoint   :78-91   (E = skew3(t) * so3exp(r); the focal is NOT used, as upstream)
  SampsonError functor   :25-76   (F = Kinv E Kinv, Kinv = diag(1,1,focal)) -> `focal_scoring=True`
and evaluation/vanilla_ransac.h:23-99 for the driver (config C4 runs VanillaMSAC);
  NonMinimalSolver       :121-144, LeastSquares :146-192 and include/RansacLib/ransac.h:127-276, 341-420 for the LO-MSAC
                         driver around the same estimator (`lo_msac` = `lo_msac_generic` + `SixPointEstimator`).  The DRIVER is pinned:
                         the reference's header compiled around a toy estimator (oracle/ref_toy.cpp) gives the same
                         trajectories as lo_msac_generic; the estimator's arithmetic stays unpinned as below.

PARITY UNPINNED: the arithmetic of the minimal solver lives in PoseLib (vlarsson/PoseLib, cloned at an
unpinned HEAD by docker/Dockerfile:58-62, absent from /root/reference and from this image); the reference has
no test, golden vector or caller for SixPointEstimator.  What is restated here is the *published* problem
(Stewenius et al. 2005 / Kukelova, Bujnak, Pajdla BMVC 2008, "Polynomial eigenvalue solutions to the 5-pt and
6-pt relative pose problems"): with F = x F0 + y F1 + F2 spanning the null space of the six epipolar
constraints, w = 1/f^2 and Q = diag(1,1,w), all real solutions (x, y, w>0) of

      det F = 0,      2 F Q F^T Q F - trace(F Q F^T Q) F = 0          (10 cubics in x,y; quadratic in w)

each decomposed into (R, t) with PoseLib's cheirality rule (every sample point in front of both cameras,
|t| = 1).  Solutions are returned sorted by focal length (PoseLib's own order is an artefact of its
elimination template and is not reproducible without its source).  This oracle solves the quadratic
eigenvalue problem with LAPACK (scipy.linalg.eig on the companion pencil); the product uses its own
Hessenberg-QR iteration, so the two are independent implementations.  Pins available: ground truth of
synthetic problems (focal, rotation, translation direction) and the residual of the polynomial system.
"""
import numpy as np
import scipy.linalg

# monomial order of the 10 cubic monomials in (x, y)
MONO = [(3, 0), (2, 1), (1, 2), (0, 3), (2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0)]
MONO_IDX = {m: i for i, m in enumerate(MONO)}


def _pmul(a, b):
    """product of polynomials in (x, y, w) stored as dense arrays P[i, j, k] = coeff of x^i y^j w^k"""
    out = np.zeros((a.shape[0] + b.shape[0] - 1, a.shape[1] + b.shape[1] - 1, a.shape[2] + b.shape[2] - 1))
    for i, j, k in zip(*np.nonzero(a)):
        out[i:i + b.shape[0], j:j + b.shape[1], k:k + b.shape[2]] += a[i, j, k] * b
    return out


def _padd(a, b):
    s = tuple(max(p, q) for p, q in zip(a.shape, b.shape))
    out = np.zeros(s)
    out[:a.shape[0], :a.shape[1], :a.shape[2]] += a
    out[:b.shape[0], :b.shape[1], :b.shape[2]] += b
    return out


def nullspace_basis(x1, x2):
    """three 3x3 matrices spanning {F : x2_i^T F x1_i = 0, i < 6}"""
    A = np.zeros((6, 9))
    for i in range(6):
        A[i] = np.outer(x2[i], x1[i]).reshape(9)
    _, _, vt = np.linalg.svd(A)
    return vt[6].reshape(3, 3), vt[7].reshape(3, 3), vt[8].reshape(3, 3)


def constraint_matrices(F0, F1, F2):
    """M0, M1, M2 (10x10): row e of (M0 + w M1 + w^2 M2) . monomials(x,y) is equation e"""
    def lin(i, j):
        p = np.zeros((2, 2, 1))
        p[1, 0, 0] = F0[i, j]
        p[0, 1, 0] = F1[i, j]
        p[0, 0, 0] = F2[i, j]
        return p
    F = [[lin(i, j) for j in range(3)] for i in range(3)]
    wpoly = np.zeros((1, 1, 2))
    wpoly[0, 0, 1] = 1.0
    q = [np.ones((1, 1, 1)), np.ones((1, 1, 1)), wpoly]  # Q = diag(1, 1, w)
    # G = F Q F^T
    G = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            acc = np.zeros((1, 1, 1))
            for k in range(3):
                acc = _padd(acc, _pmul(_pmul(F[i][k], q[k]), F[j][k]))
            G[i][j] = acc
    H = [[_pmul(G[i][j], q[j]) for j in range(3)] for i in range(3)]  # H = G Q
    tr = _padd(_padd(H[0][0], H[1][1]), H[2][2])
    eqs = []
    det = _padd(_padd(_pmul(F[0][0], _padd(_pmul(F[1][1], F[2][2]), -_pmul(F[1][2], F[2][1]))),
                      -_pmul(F[0][1], _padd(_pmul(F[1][0], F[2][2]), -_pmul(F[1][2], F[2][0])))),
                _pmul(F[0][2], _padd(_pmul(F[1][0], F[2][1]), -_pmul(F[1][1], F[2][0]))))
    eqs.append(det)
    for i in range(3):
        for j in range(3):
            acc = -_pmul(tr, F[i][j])
            for k in range(3):
                acc = _padd(acc, 2.0 * _pmul(H[i][k], F[k][j]))
            eqs.append(acc)
    M = np.zeros((3, 10, 10))
    for e, p in enumerate(eqs):
        for i, j, k in zip(*np.nonzero(p)):
            M[k, e, MONO_IDX[(i, j)]] += p[i, j, k]
    return M[0], M[1], M[2]


def _monomials(x, y):
    return np.array([x ** a * y ** b for a, b in MONO])


def _dmonomials(x, y):
    dx = np.array([a * x ** max(a - 1, 0) * y ** b if a > 0 else 0.0 for a, b in MONO])
    dy = np.array([b * x ** a * y ** max(b - 1, 0) if b > 0 else 0.0 for a, b in MONO])
    return dx, dy


def solve_xyw(M0, M1, M2, polish=12):
    """all real (x, y, w > 0): quadratic eigenvalue problem in w, then Gauss-Newton on the 10 equations"""
    Z, I = np.zeros((10, 10)), np.eye(10)
    A = np.block([[Z, I], [-M0, -M1]])
    B = np.block([[I, Z], [Z, M2]])
    with np.errstate(all="ignore"):
        vals, vecs = scipy.linalg.eig(A, B)
    sols = []
    for k in range(20):
        w = vals[k]
        # nearly real eigenvalues are kept: a close pair of real roots shows up as a conjugate pair with a
        # small imaginary part; the polish below starts on either side and the residual test decides
        if not np.isfinite(w) or abs(w.imag) > 1e-4 * max(abs(w.real), 1e-300) or w.real <= 0:
            continue
        m = vecs[:10, k]
        m = (m / m[9]).real if abs(m[9]) > 1e-14 * np.abs(m).max() else None
        if m is None:
            continue
        x, y, w = m[7], m[8], (w.real + w.imag if w.real + w.imag > 0 else w.real)
        ok = True
        converged = False
        for _ in range(polish):
            mo = _monomials(x, y)
            dx, dy = _dmonomials(x, y)
            Mw = M0 + w * M1 + w * w * M2
            res = Mw @ mo
            J = np.stack([Mw @ dx, Mw @ dy, (M1 + 2 * w * M2) @ mo], axis=1)
            step, *_ = np.linalg.lstsq(J, -res, rcond=None)
            x, y, w = x + step[0], y + step[1], w + step[2]
            if not np.isfinite([x, y, w]).all():
                ok = False
                break
            # accepted only once the iteration has settled (a start that wanders is a spurious eigenvalue)
            if (abs(step[0]) <= 1e-12 * (1 + abs(x)) and abs(step[1]) <= 1e-12 * (1 + abs(y))
                    and abs(step[2]) <= 1e-12 * abs(w)):
                converged = True
                break
        if not ok or not converged or w <= 0:
            continue
        mo = _monomials(x, y)
        Mw = M0 + w * M1 + w * w * M2
        scale = np.abs(Mw) @ np.abs(mo)
        if (np.abs(Mw @ mo) > 1e-8 * scale).any():
            continue
        if any(abs(x - s[0]) + abs(y - s[1]) < 1e-6 * (1 + abs(x) + abs(y)) and abs(w - s[2]) < 1e-6 * w for s in sols):
            continue
        sols.append((x, y, w))
    return sols


def so3exp(r):
    th = np.linalg.norm(r)
    if th < 1e-10:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def so3ln(R):
    """src/so3.cpp:25-69 semantics via a robust axis-angle extraction"""
    c = (np.trace(R) - 1) / 2
    c = min(1.0, max(-1.0, c))
    th = np.arccos(c)
    if th < 1e-10:
        return np.zeros(3)
    if np.pi - th > 1e-6:
        ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
        return th * ax
    S = (R + np.eye(3)) / 2
    i = int(np.argmax(np.diag(S)))
    ax = S[:, i] / np.sqrt(S[i, i])
    s = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if s @ ax < 0:
        ax = -ax
    return th * ax


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def _cheirality(R, t, x1, x2):
    """PoseLib's check_cheirality for unit bearing vectors, all sample points"""
    for a1, a2 in zip(x1, x2):
        Rx1 = R @ a1
        a = -Rx1 @ a2
        b1 = -Rx1 @ t
        b2 = a2 @ t
        if not (b1 - a * b2 > 0 and -a * b1 + b2 > 0):
            return False
    return True


def minimal_solver(rays6):
    """rays6: 6 x 6 (u, v) with u, v = (x, y, 1) in pixel units about the principal point.
    Returns a list of (t[3] unit, r[3], focal), sorted by focal."""
    u = np.asarray(rays6, float)[:, :3]
    v = np.asarray(rays6, float)[:, 3:]
    s = np.sqrt((np.sum(u[:, :2] ** 2) + np.sum(v[:, :2] ** 2)) / 12.0)  # conditioning only: focal is in units of s
    s = s if s > 0 else 1.0
    D = np.diag([1 / s, 1 / s, 1.0])
    x1, x2 = u @ D, v @ D
    F0, F1, F2 = nullspace_basis(x1, x2)
    M0, M1, M2 = constraint_matrices(F0, F1, F2)
    out = []
    for x, y, w in solve_xyw(M0, M1, M2):
        f = 1.0 / np.sqrt(w)
        K = np.diag([f, f, 1.0])
        E = K @ (x * F0 + y * F1 + F2) @ K
        U, _, Vt = np.linalg.svd(E)
        if np.linalg.det(U) < 0:
            U[:, 2] *= -1
        if np.linalg.det(Vt) < 0:
            Vt[2] *= -1
        W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
        Kinv = np.diag([1 / f, 1 / f, 1.0])
        b1 = x1 @ Kinv
        b2 = x2 @ Kinv
        b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
        b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
        for R in (U @ W @ Vt, U @ W.T @ Vt):
            for t in (U[:, 2], -U[:, 2]):
                if _cheirality(R, t, b1, b2):
                    out.append((t.copy(), so3ln(R), f * s))
    out.sort(key=lambda m: m[2])
    return out


def scoring_matrix(model, focal_scoring=False):
    t, r, f = model
    E = skew(t) @ so3exp(r)
    if focal_scoring:
        Kinv = np.diag([1.0, 1.0, f])
        E = Kinv @ E @ Kinv
    return E


def sampson(E, rays):
    """squared Sampson error in the operation order of the reference (six_point_estimator.cpp:85-90)"""
    u, v = rays[:, :3], rays[:, 3:]
    Eu = [(E[i, 0] * u[:, 0] + E[i, 1] * u[:, 1]) + E[i, 2] * u[:, 2] for i in range(3)]
    Etv = [(E[0, j] * v[:, 0] + E[1, j] * v[:, 1]) + E[2, j] * v[:, 2] for j in range(2)]
    d = (v[:, 0] * Eu[0] + v[:, 1] * Eu[1]) + v[:, 2] * Eu[2]
    with np.errstate(all="ignore"):
        return (d * d) / ((Eu[0] * Eu[0] + Eu[1] * Eu[1]) + (Etv[0] * Etv[0] + Etv[1] * Etv[1]))


def required_iterations(w, eta, k, lo, hi):  # include/RansacLib/utils.h:110-140
    if w <= 0.0:
        return hi
    if w >= 1.0:
        return lo
    miss = 1.0 - w ** k
    if miss >= 0.99999999999999:
        return hi
    n = np.ceil(np.log(eta) / np.log(miss) + 0.5)
    return max(lo, min(int(n), hi))


def vanilla_msac_generic(est, sampler, thr2, min_iters=100, max_iters=10000, prob=0.9999):
    """evaluation/vanilla_ransac.h:23-99 for any estimator object (see lo_msac_generic for the interface).  Pinned against the
    header itself on a toy estimator (oracle/ref_toy.cpp, test_restated_vanilla_msac_driver_equals_reference_header)."""
    n, k_min = est.n, est.min_sample_size
    st = dict(num_iterations=0, best_num_inliers=0, best_model_score=np.finfo(float).max, inlier_ratio=0.0,
              inliers=np.zeros(0, int), model=None, evals=0, status=1 if n < k_min else 2)
    if n < k_min:
        return st
    limit = max(max_iters, min_iters)
    best_min = np.finfo(float).max
    it = 0
    while it < limit:
        models = est.minimal_solver(sampler(it))
        if models:
            st["evals"] += len(models) * n
            scores = [np.minimum(est.errors(m), thr2).sum() for m in models]  # NaN-propagating like std::min(err, thr)
            k = int(np.argmin(scores))  # first minimum wins, like the strict '<' scan (ransac.h:278-293)
            if scores[k] < best_min:
                best_min = scores[k]
                st["best_model_score"] = scores[k]
                st["model"] = models[k]
                err = est.errors(models[k])
                st["errors"] = err
                st["inliers"] = np.nonzero(err < thr2)[0]
                st["best_num_inliers"] = len(st["inliers"])
                st["inlier_ratio"] = st["best_num_inliers"] / n
                limit = required_iterations(st["inlier_ratio"], 1.0 - prob, k_min, min_iters, max_iters)
                st["status"] = 0
        it += 1
    st["num_iterations"] = it
    return st


def vanilla_msac(rays, sampler, thr2, min_iters=100, max_iters=10000, prob=0.9999, focal_scoring=False,
                 scoring_of=None):
    """evaluation/vanilla_ransac.h:23-99 with the six-point estimator.  sampler(iteration) -> 6 indices.
    scoring_of(model) -> 3x3 lets a test substitute the product's scoring matrix for its own."""
    est = SixPointEstimator(rays, focal_scoring)
    if scoring_of is not None:
        est.errors = lambda m: sampson(scoring_of(m), rays)
    return vanilla_msac_generic(est, sampler, thr2, min_iters, max_iters, prob)


# ---------------------------------------------------------------------------------------------------------------
# LocallyOptimizedMSAC (include/RansacLib/ransac.h:119-430) around the six-point estimator: the driver the estimator's
# NonMinimalSolver (examples/six_point_estimator.cpp:121-144) and LeastSquares (:146-192) exist for.
# ---------------------------------------------------------------------------------------------------------------
class Mt19937:
    """std::mt19937 seeded with rng.seed(seed) (ransac.h:143-144): numpy's legacy seeding is the same init_genrand."""

    def __init__(self, seed):
        self._bg = np.random.RandomState(int(seed) & 0xFFFFFFFF)._bit_generator

    def __call__(self):
        return int(self._bg.random_raw())


def _uniform_int(rng, lo, hi):
    """std::uniform_int_distribution<int>(lo, hi)(mt19937), libstdc++ >= 11 (Lemire's nearly divisionless method)"""
    rng_range = hi - lo + 1
    prod = rng() * rng_range
    low = prod & 0xFFFFFFFF
    if low < rng_range:
        threshold = ((1 << 32) - rng_range) % rng_range
        while low < threshold:
            prod = rng() * rng_range
            low = prod & 0xFFFFFFFF
    return lo + (prod >> 32)


def shuffle_and_resize(v, target, rng):
    """utils::RandomShuffleAndResize (include/RansacLib/utils.h:34-73); std::vector::resize pads with zeros"""
    v = list(v)
    n = len(v)
    for i in range(n - 1):
        j = _uniform_int(rng, i, n - 1)
        v[i], v[j] = v[j], v[i]
    return v[:target] + [0] * max(0, target - n)


def lo_msac_generic(est, sampler, thr2, seed=0, min_iters=100, max_iters=10000, prob=0.9999, num_lo_steps=10,
                    threshold_multiplier=np.sqrt(2.0), num_lsq_iterations=4, min_sample_multiplicator=7,
                    non_min_sample_multiplier=3, lo_starting_iterations=50, final_least_squares=False):
    """LocallyOptimizedMSAC::EstimateModel (include/RansacLib/ransac.h:127-276) with LocalOptimization (:341-407) and
    LeastSquaresFit (:409-420), for any estimator object with
        n, min_sample_size, non_minimal_sample_size,
        minimal_solver(sample) -> list of models, non_minimal_solver(sample) -> model or None,
        least_squares(sample, model) -> model, errors(model) -> squared error of every data point.
    sampler(iteration) -> minimal sample.  Pinned against the header itself on a toy estimator both sides can compute
    (oracle/ref_toy.cpp, tests/test_sixpt.py::test_restated_lo_msac_driver_equals_reference_header)."""
    n, k_min, k_nonmin = est.n, est.min_sample_size, est.non_minimal_sample_size
    BIG = np.finfo(float).max
    st = dict(num_iterations=0, best_num_inliers=0, best_model_score=BIG, inlier_ratio=0.0, inliers=np.zeros(0, int),
              model=None, number_lo_iterations=0, status=1 if n < k_min else 2, lm_calls=0)
    if n < k_min:
        return st
    rng = Mt19937(seed)

    def score_model(m):  # ScoreModel :295-303 (sequential sum)
        e = np.minimum(est.errors(m), thr2)
        s = 0.0
        for x in e:
            s += x
        return s

    def get_inliers(m, thr):
        return np.nonzero(est.errors(m) < thr)[0]

    def lsq(m, sample):
        st["lm_calls"] += 1
        return est.least_squares(sample, m)

    def least_squares_fit(thresh, m):  # :409-420
        inl = get_inliers(m, thresh)
        if len(inl) < k_min:
            return m
        k = min(min_sample_multiplicator * k_min, len(inl))
        return lsq(m, shuffle_and_resize(inl, k, rng))

    def local_optimization(m_best, score_best):  # :341-407
        if k_nonmin > n:
            return m_best, score_best
        m_init = least_squares_fit(thr2 * threshold_multiplier, m_best)
        score = score_model(m_init)
        if score < score_best:
            score_best, m_best = score, m_init
        base = get_inliers(m_init, thr2 * threshold_multiplier)
        non_min = max(k_nonmin, min(k_min * non_min_sample_multiplier, len(base) // 2))
        for _ in range(num_lo_steps):
            sample = shuffle_and_resize(base, non_min, rng)
            m = est.non_minimal_solver(np.asarray(sample, int))
            if m is None:
                continue
            score = score_model(m)
            if score < score_best:
                score_best, m_best = score, m
            m = least_squares_fit(thr2, m)
            th = threshold_multiplier * thr2
            with np.errstate(all="ignore"):
                dth = np.float64((threshold_multiplier - 1.0) * thr2) / np.float64(int(num_lsq_iterations - 1))
            for _i in range(num_lsq_iterations):
                m = least_squares_fit(th, m)
                score = score_model(m)
                if score < score_best:
                    score_best, m_best = score, m
                th -= dth
        return m_best, score_best

    def refresh(update_max):
        nonlocal limit
        st["inliers"] = get_inliers(st["model"], thr2)
        st["best_num_inliers"] = len(st["inliers"])
        st["inlier_ratio"] = st["best_num_inliers"] / n
        if update_max:
            limit = required_iterations(st["inlier_ratio"], 1.0 - prob, k_min, min_iters, max_iters)

    limit = max(max_iters, min_iters)
    best_min, best_min_score = None, BIG
    it = 0
    while it < limit:
        if it == lo_starting_iterations and best_min_score < BIG:  # :166-177
            st["number_lo_iterations"] += 1
            st["model"], st["best_model_score"] = local_optimization(st["model"], st["best_model_score"])
            refresh(True)
        models = est.minimal_solver(sampler(it))
        if models:
            scores = [score_model(m) for m in models]
            k = int(np.argmin(scores))  # first minimum, like the strict '<' scan (:278-293); NaN scores never win there
            local = scores[k] if scores[k] == scores[k] else BIG
            if local < best_min_score or it == lo_starting_iterations:
                better = local < best_min_score
                if better:
                    best_min_score, best_min = local, models[k]
                    if best_min_score < st["best_model_score"]:
                        st["best_model_score"], st["model"] = best_min_score, best_min
                    st["status"] = 0
                run_lo = it >= lo_starting_iterations and best_min_score < BIG
                if better or run_lo:
                    if run_lo:
                        st["number_lo_iterations"] += 1
                        best_min, score = local_optimization(best_min, best_min_score)
                        if score < st["best_model_score"]:
                            st["best_model_score"], st["model"] = score, best_min
                    refresh(True)
        it += 1
    st["num_iterations"] = it
    if it <= lo_starting_iterations and st["best_model_score"] < BIG:  # :246-257
        st["number_lo_iterations"] += 1
        st["model"], st["best_model_score"] = local_optimization(st["model"], st["best_model_score"])
        refresh(False)
    if final_least_squares and st["model"] is not None:  # :259-275
        refined = lsq(st["model"], st["inliers"])
        score = score_model(refined)
        if score < st["best_model_score"]:
            st["best_model_score"], st["model"] = score, refined
            refresh(False)
    return st


class SixPointEstimator:
    """examples/six_point_estimator.{h,cpp} as an estimator object for lo_msac_generic: min_sample_size 6,
    non_minimal_sample_size 7 (six_point_estimator.h:21-23)."""
    min_sample_size, non_minimal_sample_size = 6, 7

    def __init__(self, rays, focal_scoring=False, solve=None):
        self.rays, self.n, self.focal_scoring, self.solve = rays, len(rays), focal_scoring, solve or minimal_solver

    def errors(self, m):
        return sampson(scoring_matrix(m, self.focal_scoring), self.rays)

    def minimal_solver(self, sample):  # :93-119: the first six entries of the sample
        return self.solve(self.rays[np.asarray(sample, int)[:6]])

    def non_minimal_solver(self, sample):  # :121-144: MinimalSolver on the sample, summed error over the sample decides
        sols = self.minimal_solver(sample)
        if not sols:
            return None
        best, best_score = 0, np.inf
        for i, m in enumerate(sols):
            s = 0.0
            for x in sampson(scoring_matrix(m, self.focal_scoring), self.rays[sample]):
                s += x
            if s < best_score:
                best_score, best = s, i
        return sols[best]

    def least_squares(self, sample, m):  # :146-192
        return least_squares(self.rays, sample, m)[0]


def lo_msac(rays, sampler, thr2, focal_scoring=False, solve=None, **options):
    """LocallyOptimizedMSAC around SixPointEstimator.  sampler(iteration) -> 6 indices (the product's Philox stream);
    solve(rays6) substitutes the minimal solver; options as lo_msac_generic."""
    return lo_msac_generic(SixPointEstimator(rays, focal_scoring, solve), sampler, thr2, **options)


def make_problem(rng, n, focal, outlier_frac=0.0, noise_px=0.0, max_angle_deg=20.0):
    """synthetic shared-focal pair in pixel units about the principal point (config C4)"""
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(rng.uniform(-1, 1) * max_angle_deg)
    R = so3exp(ax * ang)
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(4, 8, n)], axis=1)
    Y = X @ R.T + t
    u = np.stack([focal * X[:, 0] / X[:, 2], focal * X[:, 1] / X[:, 2], np.ones(n)], axis=1)
    v = np.stack([focal * Y[:, 0] / Y[:, 2], focal * Y[:, 1] / Y[:, 2], np.ones(n)], axis=1)
    u[:, :2] += rng.normal(size=(n, 2)) * noise_px
    v[:, :2] += rng.normal(size=(n, 2)) * noise_px
    no = int(round(outlier_frac * n))
    if no:
        idx = rng.choice(n, no, replace=False)
        v[idx, :2] = rng.uniform(-0.5 * focal, 0.5 * focal, size=(no, 2))
    return np.concatenate([u, v], axis=1), R, t, focal


# ---------------------------------------------------------------------------------------------------------------
# SixPointEstimator::LeastSquares (examples/six_point_estimator.cpp:146-192).  PARITY UNPINNED (Ceres is absent): the
# trust-region LM with Solver::Options defaults, the autodiff'd SampsonError functor (:25-76) and
# ceres::SphereManifold<3> (Plus / PlusJacobian in Householder form) are restated from Ceres' documentation.
# ---------------------------------------------------------------------------------------------------------------
def _col(a):
    """broadcast a value part against a (..., 7) partials part"""
    return a[..., None] if isinstance(a, np.ndarray) and a.ndim else a


class _Dual:
    """forward-mode scalar with a 7-vector of partials (r1, t1, focal) -- the role of ceres::Jet<double, 7>.  The value may
    be an array (one entry per residual, partials (n, 7)): the same elementwise operations, evaluated for all residuals at once."""
    __slots__ = ("a", "v")

    def __init__(self, a, v=None):
        self.a = a if isinstance(a, np.ndarray) else float(a)
        self.v = np.zeros(7) if v is None else v

    @staticmethod
    def var(a, k):
        v = np.zeros(7)
        v[k] = 1.0
        return _Dual(a, v)

    def _c(self, o):
        return o if isinstance(o, _Dual) else _Dual(o)

    def __add__(self, o):
        o = self._c(o)
        return _Dual(self.a + o.a, self.v + o.v)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._c(o)
        return _Dual(self.a - o.a, self.v - o.v)

    def __rsub__(self, o):
        return self._c(o) - self

    def __neg__(self):
        return _Dual(-self.a, -self.v)

    def __mul__(self, o):
        o = self._c(o)
        return _Dual(self.a * o.a, _col(self.a) * o.v + self.v * _col(o.a))
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._c(o)
        q = self.a / o.a
        return _Dual(q, (self.v - _col(q) * o.v) / _col(o.a))


def _dsqrt(x):
    r = np.sqrt(x.a)
    return _Dual(r, x.v * (0.5 / r))


def _sampson_functor(x7, u, v):
    """SampsonError::operator() with r0 = t0 = 0 (Ri = I): returns the residual as a _Dual"""
    r = [x7[0], x7[1], x7[2]]
    t = [x7[3], x7[4], x7[5]]
    f = x7[6]
    th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2]
    if th2.a > np.finfo(float).eps:  # ceres::AngleAxisToRotationMatrix
        th = _dsqrt(th2)
        wx, wy, wz = r[0] / th, r[1] / th, r[2] / th
        ct, st = _Dual(np.cos(th.a), -np.sin(th.a) * th.v), _Dual(np.sin(th.a), np.cos(th.a) * th.v)
        oc = 1.0 - ct
        R = [[ct + wx * wx * oc, wx * wy * oc - wz * st, wy * st + wx * wz * oc],
             [wz * st + wx * wy * oc, ct + wy * wy * oc, wy * wz * oc - wx * st],
             [wx * wz * oc - wy * st, wx * st + wy * wz * oc, ct + wz * wz * oc]]
    else:
        one = _Dual(1.0)
        R = [[one, -r[2], r[1]], [r[2], one, -r[0]], [-r[1], r[0], one]]
    sk = [[_Dual(0.0), -t[2], t[1]], [t[2], _Dual(0.0), -t[0]], [-t[1], t[0], _Dual(0.0)]]
    E = [[sk[i][0] * R[0][j] + sk[i][1] * R[1][j] + sk[i][2] * R[2][j] for j in range(3)] for i in range(3)]
    k = [_Dual(1.0), _Dual(1.0), f]
    F = [[k[i] * E[i][j] * k[j] for j in range(3)] for i in range(3)]
    Fu = [F[i][0] * u[0] + F[i][1] * u[1] + F[i][2] * u[2] for i in range(3)]
    Ftv = [F[0][j] * v[0] + F[1][j] * v[1] + F[2][j] * v[2] for j in range(3)]
    d = Fu[0] * v[0] + Fu[1] * v[1] + Fu[2] * v[2]
    return (d * d) / (Fu[0] * Fu[0] + Fu[1] * Fu[1] + Ftv[0] * Ftv[0] + Ftv[1] * Ftv[1])


def _householder3(x):
    sigma = x[0] * x[0] + x[1] * x[1]
    v = np.array([x[0], x[1], 1.0])
    if sigma <= np.finfo(float).eps:
        return v, (2.0 if x[2] < 0 else 0.0)
    mu = np.sqrt(x[2] * x[2] + sigma)
    vp = x[2] - mu if x[2] <= 0 else -sigma / (x[2] + mu)
    beta = 2.0 * vp * vp / (sigma + vp * vp)
    v[:2] /= vp
    return v, beta


def sphere_plus(x, d):
    nd = np.linalg.norm(d)
    if nd == 0:
        return np.array(x, float)
    v, beta = _householder3(x)
    y = np.array([np.sin(nd) / nd * d[0], np.sin(nd) / nd * d[1], np.cos(nd)])
    return np.linalg.norm(x) * (y - v * (beta * (v @ y)))


def sphere_plus_jacobian(x):
    v, beta = _householder3(x)
    return np.linalg.norm(x) * (np.eye(3) - beta * np.outer(v, v))[:, :2]


def least_squares(rays, sample, model, max_iters=200):
    """returns (refined model, iterations, initial cost, final cost)"""
    t, r, f = model
    x = np.concatenate([np.asarray(r, float), np.asarray(t, float), [float(f)]])
    pts = np.asarray(sample, int)
    U = np.ascontiguousarray(rays[pts, :3].T)  # all residuals at once: the functor's arithmetic is elementwise
    V = np.ascontiguousarray(rays[pts, 3:].T)

    def eval_full(xx):
        X = [_Dual.var(xx[k], k) for k in range(7)]
        if len(pts) == 0:
            return np.zeros(0), np.zeros((0, 6))
        res = _sampson_functor(X, U, V)
        rv = np.array(res.a, float).reshape(-1)
        Ja = np.broadcast_to(res.v, (len(pts), 7))
        Pt = sphere_plus_jacobian(xx[3:6])
        Jl = np.concatenate([Ja[:, :3], Ja[:, 3:6] @ Pt, Ja[:, 6:7]], axis=1)
        return rv, Jl

    def eval_cost(xx):
        X = [_Dual(xx[k]) for k in range(7)]
        if len(pts) == 0:
            return 0.0
        rv = np.array(_sampson_functor(X, U, V).a, float).reshape(-1)
        return 0.5 * float(rv @ rv)

    res, J = eval_full(x)
    x_cost = 0.5 * float(res @ res)
    initial = x_cost
    if not np.isfinite(x_cost):
        return model, 0, initial, x_cost
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0))) if len(pts) else np.ones(6)
    J = J * scale
    grad_max = np.abs((J / scale).T @ res).max() if len(pts) else 0.0
    radius, decrease, reuse, invalid, it = 1e4, 2.0, False, 0, 0
    diagonal = np.zeros(6)
    while True:
        if it >= max_iters or grad_max <= 1e-10 or radius < 1e-32:
            break
        it += 1
        H = J.T @ J
        g = J.T @ res
        if not reuse:
            diagonal = np.minimum(np.maximum(np.diag(H), 1e-6), 1e32)
        reuse = True
        valid = True
        try:
            L = np.linalg.cholesky(H + np.diag(diagonal / radius))
            step = -np.linalg.solve(L.T, np.linalg.solve(L, g))
            valid = bool(np.isfinite(step).all())
        except np.linalg.LinAlgError:
            valid = False
        mcc = 0.0
        if valid:
            m = J @ step
            mcc = -float(m @ (res + m / 2.0))
            valid = mcc > 0.0
        if not valid:
            invalid += 1
            if invalid >= 10:
                break
            radius /= decrease
            decrease *= 2.0
            continue
        invalid = 0
        delta = step * scale
        cand = np.concatenate([x[:3] + delta[:3], sphere_plus(x[3:6], delta[3:5]), [x[6] + delta[5]]])
        step_norm, x_norm = np.linalg.norm(cand - x), np.linalg.norm(x)
        cand_cost = eval_cost(cand)
        if not np.isfinite(cand_cost):
            cand_cost = np.finfo(float).max
        if step_norm <= 1e-8 * (x_norm + 1e-8):
            break
        change = x_cost - cand_cost
        if abs(change) <= 1e-6 * x_cost:
            break
        rho = change / mcc
        if rho > 1e-3:
            x = cand
            res, J = eval_full(x)
            x_cost = 0.5 * float(res @ res)
            grad_max = np.abs(J.T @ res).max()
            J = J * scale
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
        else:
            radius /= decrease
            decrease *= 2.0
    return (x[3:6].copy(), x[:3].copy(), float(x[6])), it, initial, x_cost
