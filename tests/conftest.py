import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

THR2 = (2.0 / 600.0) ** 2  # evaluation/test_ransac.cpp:23,52


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def E_of(m):
    """6-parameter spherical model -> 3x3 (src/spherical_solvers.cpp:299-303)."""
    return np.array([[m[0], m[1], m[2]], [m[1], -m[0], m[3]], [m[4], m[5], 0.0]])


def model_dist(a, b):
    """Sign-invariant distance between two unit-Frobenius models."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b))


def match_models(A, B):
    """max over valid models of A of the distance to the nearest model of B (root matching)."""
    A = [a for a in A if not np.isnan(a).any()]
    B = [b for b in B if not np.isnan(b).any()]
    if not A:
        return 0.0
    if not B:
        return np.inf
    return max(min(model_dist(a, b) for b in B) for a in A)


@pytest.fixture(scope="session")
def S():
    import spherical_sfm_b200
    return spherical_sfm_b200


@pytest.fixture(scope="session")
def O():
    import oracle
    return oracle


@pytest.fixture(scope="session")
def orc(O):
    return O.load()


@pytest.fixture(scope="session")
def ref(O):
    """oracle/_ref: driver loops = the reference's own RansacLib.  None if it was never built."""
    return O.load_ref()


@pytest.fixture(scope="session")
def reffull(O):
    """oracle/_ref/libssfm_reffull.so: the reference's own sources (RansacLib, SphericalEstimator, solvers)
    compiled unmodified against Eigen/Ceres stand-ins.  None if it was never built."""
    return O.load_ref_full()


@pytest.fixture(scope="session")
def engine(S):
    import __graft_entry__ as G
    if not os.path.exists(S.LIB_PATH):
        G.build()
    eng = S.Engine(0)  # raises (loudly) without a GPU: gpu-marked tests only
    yield eng
    eng.close()


class HsParams(C.Structure):
    _fields_ = [('min_iters', C.c_uint32), ('max_iters', C.c_uint32), ('success_probability', C.c_double),
                ('thr2', C.c_double), ('seed', C.c_uint32), ('num_lo_steps', C.c_int32), ('thr_mult', C.c_double),
                ('num_lsq_iters', C.c_int32), ('min_sample_mult', C.c_int32), ('non_min_mult', C.c_int32),
                ('lo_start', C.c_uint32), ('final_lsq', C.c_int32), ('solver', C.c_int32), ('driver', C.c_int32),
                ('inward', C.c_int32), ('fixed_budget', C.c_int32), ('fixed_prob', C.c_double),
                ('cand_margin', C.c_float), ('first_round', C.c_int32), ('round_cap', C.c_int32), ('defer', C.c_int32),
                ('skip_complex', C.c_int32)]


class HsResult(C.Structure):
    _fields_ = [('E', C.c_double * 9), ('r', C.c_double * 3), ('t', C.c_double * 3), ('best_model_score', C.c_double),
                ('inlier_ratio', C.c_double), ('num_iterations', C.c_uint32), ('best_num_inliers', C.c_int32),
                ('num_lo', C.c_int32), ('status', C.c_int32), ('evals_exact', C.c_int64), ('rounds', C.c_int32),
                ('candidates', C.c_int32)]


class HostShim:
    """tests/hostshim/libhostshim.so: TEST-ONLY host build of the engine's __host__ __device__ code."""

    def __init__(self):
        import __graft_entry__ as G
        self.lib = C.CDLL(G.build_hostshim())
        self.lib.hs_solve.restype = C.c_int
        self.lib.hs_required_iterations.restype = C.c_uint32

    @staticmethod
    def dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    @staticmethod
    def ip(a):
        return a.ctypes.data_as(C.POINTER(C.c_int))

    def solve(self, rays, sample, kind, skip_complex=False):
        rays = np.ascontiguousarray(rays, np.float64)
        s = np.ascontiguousarray(sample, np.int32)
        models = np.full((4, 6), np.nan)
        nm = self.lib.hs_solve(self.dp(rays), self.ip(s), kind, int(skip_complex), self.dp(models))
        return nm, models

    def estimate_pair(self, rays, opt, pair_id, margin=2e-4, first=128, cap=256, defer=1):
        rays = np.ascontiguousarray(rays, np.float64)
        n = len(rays)
        hp = HsParams(opt.min_num_iterations, opt.max_num_iterations, opt.success_probability,
                      opt.squared_inlier_threshold, opt.random_seed, opt.num_lo_steps, opt.threshold_multiplier,
                      opt.num_lsq_iterations, opt.min_sample_multiplicator, opt.non_min_sample_multiplier,
                      opt.lo_starting_iterations, opt.final_least_squares, opt.solver_kind, opt.driver, opt.inward,
                      opt.legacy_budget, opt.legacy_prob_success, margin, first, cap, defer,
                      int(opt.complex_mode == 2))  # OrcOptions.complex_mode: 0 canonical, 2 skip
        res = HsResult()
        flags = np.zeros(max(n, 1), np.uint8)
        self.lib.hs_estimate_pair(self.dp(rays), n, C.byref(hp), C.c_uint32(pair_id), C.byref(res),
                                  flags.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return res, flags[:n]


    def triangulate(self, cam_tr, obs_xy, focal, opt, point_id, chunk=0):
        """csrc/ssfm_triangulate.cuh on the host: (X, num_inliers, iterations, num_lo)."""
        cam_tr = np.ascontiguousarray(cam_tr, np.float64)
        obs_xy = np.ascontiguousarray(obs_xy, np.float64)
        hp = HsParams(opt.min_num_iterations, opt.max_num_iterations, opt.success_probability,
                      opt.squared_inlier_threshold, opt.random_seed, opt.num_lo_steps, opt.threshold_multiplier,
                      opt.num_lsq_iterations, opt.min_sample_multiplicator, opt.non_min_sample_multiplier,
                      opt.lo_starting_iterations, opt.final_least_squares, 0, 0, 0, 0, 0.0, 0.0, 0, 0, 0, 0)
        X = np.zeros(3)
        it = C.c_uint32()
        nlo = C.c_int()
        n = self.lib.hs_triangulate(self.dp(cam_tr), self.dp(obs_xy), len(obs_xy), C.c_double(focal), C.byref(hp),
                                    C.c_uint32(point_id), self.dp(X), C.byref(it), C.byref(nlo), C.c_uint32(chunk))
        return X, n, it.value, nlo.value


@pytest.fixture(scope="session")
def shim():
    return HostShim()


def to_oracle_options(O, opt):
    """SsfmOptions -> OrcOptions."""
    o = O.default_options()
    ren = {"solver": "solver_kind", "fixed_budget": "legacy_budget", "fixed_prob_success": "legacy_prob_success"}
    for f, _ in opt._fields_:
        of = ren.get(f, f)
        if hasattr(o, of):
            setattr(o, of, getattr(opt, f))
    o.complex_mode = 2 if opt.complex_root_models == 1 else 0  # SSFM_COMPLEX_SKIP -> COMPLEX_SKIP, else canonical
    return o


def check_full_path_goldens(S, g, run_case, stride=1):
    """tests/golden/refsrc_golden.npz, full 3-point path (168 cases made by the reference's own sources in oracle/_ref).
    run_case(rays, cfg, skip) -> ((status, iterations, inliers, lo_count), r, E, flags) for the implementation under test.
      * "skip" goldens (complex action-matrix eigenvalues masked upstream): every case identical -- status, iteration
        count, LO count, inlier flags -- and the pose within 0.01 deg.
      * "up" goldens (upstream as written): identical exactly on the cases recorded in fu_same_canonical (the others are
        the ones where a complex-eigenvalue model, whose phase upstream is rounding noise, wins an iteration)."""
    import hashlib
    cfgs = g["fu_cfg"]
    followed, total = 0, 0
    for k in range(0, len(cfgs), stride):
        cfg = tuple(int(x) for x in cfgs[k])
        n, nout, pid, inward, flsq, kind = cfg
        pr = S.problems.make_problem(S.problems.make_rng(2026, pid), n, bool(inward), None, 1 / 600, nout, 20.0)
        assert hashlib.sha256(np.ascontiguousarray(pr.rays).tobytes()).hexdigest() == str(g["fu_sha"][k])
        for mode, skip in (("skip", True), ("up", False)):
            st, r, E, flags = run_case(pr.rays, cfg, skip)
            o0, o1 = g["fu_inl_off_" + mode][k], g["fu_inl_off_" + mode][k + 1]
            gflags = np.unpackbits(g["fu_inl_" + mode][o0:o1])[:n]
            same = tuple(int(x) for x in g["fu_stats_" + mode][k]) == tuple(int(x) for x in st) and (gflags == flags).all()
            if mode == "skip":
                assert same, (k, cfg, st, g["fu_stats_skip"][k])
            else:
                total += 1
                followed += bool(same)
                assert bool(same) == bool(g["fu_same_canonical"][k]), (k, cfg, st, g["fu_stats_up"][k])
            if same:
                d = S.problems.rot_error(S.problems.so3exp(g["fu_r_" + mode][k]), S.problems.so3exp(np.asarray(r)))
                assert np.rad2deg(d) < 0.01
                Eg = g["fu_E_" + mode][k]
                assert model_dist(np.asarray(E) / np.linalg.norm(E), Eg / np.linalg.norm(Eg)) < 1e-5
    print("full-path goldens: follows upstream-as-written on %d / %d cases, masked upstream on all" % (followed, total))
