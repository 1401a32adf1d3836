"""ctypes loader for the CPU ORACLE (test infrastructure -- see oracle/ssfm_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  `load()` returns the restated oracle (oracle/liboracle.so, built on demand
with g++); `load_ref()` returns oracle/_ref/libssfm_ref.so, whose RANSAC driver loops are the
reference's own RansacLib headers compiled from /root/reference (prebuilt; travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
COMPLEX_CANONICAL, COMPLEX_EIGEN, COMPLEX_SKIP = 0, 1, 2  # ssfm_oracle.hpp ComplexRootMode


class OrcOptions(C.Structure):
    _fields_ = [
        ("min_num_iterations", C.c_uint32), ("max_num_iterations", C.c_uint32),
        ("success_probability", C.c_double), ("squared_inlier_threshold", C.c_double),
        ("random_seed", C.c_uint32), ("num_lo_steps", C.c_int32),
        ("threshold_multiplier", C.c_double), ("num_lsq_iterations", C.c_int32),
        ("min_sample_multiplicator", C.c_int32), ("non_min_sample_multiplier", C.c_int32),
        ("lo_starting_iterations", C.c_uint32), ("final_least_squares", C.c_int32),
        ("solver_kind", C.c_int32), ("driver", C.c_int32), ("inward", C.c_int32),
        ("legacy_budget", C.c_int32), ("legacy_prob_success", C.c_double),
        ("preemptive_block", C.c_int32), ("complex_mode", C.c_int32),
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("E", C.c_double * 9), ("r", C.c_double * 3), ("t", C.c_double * 3),
        ("num_iterations", C.c_uint32), ("best_num_inliers", C.c_int32),
        ("best_model_score", C.c_double), ("inlier_ratio", C.c_double),
        ("number_lo_iterations", C.c_int32), ("status", C.c_int32), ("evals", C.c_int64),
    ]


def default_options(**kw):
    """RansacLib defaults (include/RansacLib/ransac.h:49-73)."""
    o = OrcOptions(100, 10000, 0.9999, 1.0, 0, 10, 2.0 ** 0.5, 4, 7, 3, 50, 0, 0, 0, 0, 512, 0.999, 10, 0)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def pipeline_options(thr2, **kw):
    """estimate_pairwise's options (examples/spherical_sfm_tools.cpp:314-318)."""
    return default_options(squared_inlier_threshold=thr2, num_lo_steps=0, num_lsq_iterations=0,
                           final_least_squares=1, **kw)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    def __init__(self, path):
        self.path = path
        L = self.lib = C.CDLL(path)
        L.orc_is_reference.restype = C.c_int
        L.orc_estimate_batch.restype = C.c_double
        L.orc_score_batch.restype = C.c_double
        L.orc_estimate_pair.restype = C.c_int
        L.orc_solve.restype = C.c_int
        L.orc_solve_mode.restype = C.c_int
        self.is_reference = bool(L.orc_is_reference())

    def philox_sample(self, seed, pair, it, k, n):
        idx = np.zeros(k, np.int32)
        self.lib.orc_philox_sample(C.c_uint32(seed), C.c_uint32(pair), C.c_uint32(it), k, n, _ip(idx))
        return idx

    def triangulate(self, cam_tr, obs_xy, focal, opt, point_id=0):
        """SfM::Retriangulate for one point: cam_tr (n,6) = t, r of each observation's camera; obs_xy (n,2)."""
        cam_tr = np.ascontiguousarray(cam_tr, np.float64)
        obs_xy = np.ascontiguousarray(obs_xy, np.float64)
        res = OrcResult()
        inl = np.zeros(max(len(obs_xy), 1), np.int32)
        n = self.lib.orc_triangulate(_dp(cam_tr), _dp(obs_xy), len(obs_xy), C.c_double(focal), C.byref(opt), C.c_uint32(point_id),
                                     C.byref(res), _ip(inl))
        return res, inl[:max(n, 0)].copy()

    def knuth_sample(self, seed, pair, hyp, n_total, k):
        idx = np.zeros(k, np.int32)
        self.lib.orc_knuth_sample(C.c_uint32(seed), C.c_uint32(pair), C.c_uint32(hyp), n_total, k, _ip(idx))
        return idx

    def solve(self, rays, sample, kind=0, complex_mode=COMPLEX_CANONICAL):
        """The minimal solver on one sample.  complex_mode: what to return for action-matrix models that come from a
        complex eigenvalue (ssfm_oracle.hpp ComplexRootMode).  Reference-source builds ignore CANONICAL (they always
        return what the reference returns) and honour SKIP through the mask in the Eigen stand-in."""
        rays = np.ascontiguousarray(rays, np.float64)
        sample = np.ascontiguousarray(sample, np.int32)
        models = np.zeros((4, 6))
        nm = self.lib.orc_solve_mode(_dp(rays), _ip(sample), len(sample), kind, int(complex_mode), _dp(models))
        return nm, models

    def eigen34(self, M):
        """Eigen::EigenSolver<Matrix4d>(M) restated: (eigenvalues[4] complex, eigenvectors[4,4] complex, ok)."""
        M = np.ascontiguousarray(M, np.float64).reshape(16)
        ev = np.zeros(8)
        V = np.zeros(32)
        ok = self.lib.orc_eigen34(_dp(M), _dp(ev), _dp(V))
        return ev[0::2] + 1j * ev[1::2], (V[0::2] + 1j * V[1::2]).reshape(4, 4), bool(ok)

    def sampson(self, E, rays):
        rays = np.ascontiguousarray(rays, np.float64)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        out = np.zeros(len(rays))
        self.lib.orc_sampson(_dp(E), _dp(rays), len(rays), _dp(out))
        return out

    def score(self, E, rays, thr2):
        rays = np.ascontiguousarray(rays, np.float64)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        s = C.c_double()
        n = C.c_int()
        self.lib.orc_score(_dp(E), _dp(rays), len(rays), C.c_double(thr2), C.byref(s), C.byref(n))
        return s.value, n.value

    def decompose(self, E, inward=False):
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        r = np.zeros(3)
        t = np.zeros(3)
        self.lib.orc_decompose(_dp(E), int(inward), _dp(r), _dp(t))
        return r, t

    def make_E(self, r, inward=False):
        r = np.ascontiguousarray(r, np.float64)
        E = np.zeros(9)
        self.lib.orc_make_E(_dp(r), int(inward), _dp(E))
        return E.reshape(3, 3)

    def lm_refit(self, rays, sample, E, inward=False):
        rays = np.ascontiguousarray(rays, np.float64)
        sample = np.ascontiguousarray(sample, np.int32)
        E = np.array(E, np.float64).reshape(9).copy()
        it = C.c_int()
        term = C.c_int()
        costs = np.zeros(2)
        self.lib.orc_lm_refit(_dp(rays), _ip(sample), len(sample), int(inward), _dp(E), C.byref(it), C.byref(term),
                              _dp(costs))
        return E.reshape(3, 3), it.value, term.value, costs

    def lo_shuffle(self, seed, sizes, targets):
        sizes = np.ascontiguousarray(sizes, np.int32)
        targets = np.ascontiguousarray(targets, np.int32)
        out = np.zeros(int(targets.sum()), np.int32)
        self.lib.orc_lo_shuffle(C.c_uint32(seed), len(sizes), _ip(sizes), _ip(targets), _ip(out))
        return out

    def estimate_pair(self, rays, opt, pair_id=0):
        rays = np.ascontiguousarray(rays, np.float64)
        res = OrcResult()
        inl = np.zeros(max(len(rays), 1), np.int32)
        n = self.lib.orc_estimate_pair(_dp(rays), len(rays), C.byref(opt), C.c_uint32(pair_id), C.byref(res), _ip(inl))
        return res, inl[:max(n, 0)].copy()

    def estimate_batch(self, rays, offsets, opt, first_pair_id=0, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = len(offsets) - 1
        res = (OrcResult * P)()
        secs = self.lib.orc_estimate_batch(_dp(rays), offsets.ctypes.data_as(C.POINTER(C.c_int64)), P, C.byref(opt),
                                           C.c_uint32(first_pair_id), nthreads, res)
        return res, secs

    def estimate_batch_flags(self, rays, offsets, opt, first_pair_id=0, nthreads=0):
        """estimate_batch + the final inlier set as one byte per correspondence.  Returns (results, flags, seconds)."""
        rays = np.ascontiguousarray(rays, np.float64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        P = len(offsets) - 1
        res = (OrcResult * P)()
        flags = np.zeros(max(int(offsets[-1]), 1), np.uint8)
        self.lib.orc_estimate_batch_flags.restype = C.c_double
        secs = self.lib.orc_estimate_batch_flags(_dp(rays), offsets.ctypes.data_as(C.POINTER(C.c_int64)), P, C.byref(opt),
                                                 C.c_uint32(first_pair_id), nthreads, res,
                                                 flags.ctypes.data_as(C.POINTER(C.c_uint8)))
        return res, flags[:int(offsets[-1])], secs

    def score_batch(self, models6, rays, thr2, nthreads=0):
        models6 = np.ascontiguousarray(models6, np.float64)
        rays = np.ascontiguousarray(rays, np.float64)
        M = len(models6)
        scores = np.zeros(M)
        ninl = np.zeros(M, np.int32)
        secs = self.lib.orc_score_batch(_dp(models6), M, _dp(rays), len(rays), C.c_double(thr2), nthreads, _dp(scores),
                                        _ip(ninl))
        return scores, ninl, secs


def _compiler():
    for c in ("/usr/bin/g++", "g++"):
        if os.path.exists(c) or c == "g++":
            return c


def build(ref=True, force=False):
    """Compile liboracle.so and (when /root/reference is present) _ref/libssfm_ref.so."""
    srcs = [os.path.join(_DIR, f) for f in ("oracle_capi.cpp", "oracle_capi.h", "ssfm_oracle.hpp", "lomsac.hpp", "tri_oracle.hpp")]
    newest = max(os.path.getmtime(s) for s in srcs)
    flags = ["-O3", "-std=c++17", "-fPIC", "-pthread", "-shared"]
    out = os.path.join(_DIR, "liboracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call([_compiler()] + flags + ["-o", out, srcs[0]])
    refroot = os.environ.get("SSFM_REFERENCE_ROOT", "/root/reference")
    if ref and os.path.isdir(os.path.join(refroot, "include", "RansacLib")):
        os.makedirs(os.path.join(_DIR, "_ref"), exist_ok=True)
        rout = os.path.join(_DIR, "_ref", "libssfm_ref.so")
        if force or not os.path.exists(rout) or os.path.getmtime(rout) < newest:
            subprocess.check_call([_compiler()] + flags + ["-DSSFM_USE_REFERENCE_RANSACLIB",
                                                         "-I" + os.path.join(refroot, "include"),
                                                         "-I" + os.path.join(refroot, "evaluation"),
                                                         "-o", rout, srcs[0]])
        # the reference's own sources against the Eigen / Ceres stand-ins (see Makefile `ref`)
        shim = [os.path.join(_DIR, "eigen_shim", "ssfm_mini_eigen.hpp"), os.path.join(_DIR, "ceres_shim", "ssfm_mini_ceres.hpp")]
        refsrc = [os.path.join(refroot, "src", f) for f in ("spherical_solvers.cpp", "so3.cpp", "spherical_utils.cpp")]
        inc = ["-I" + _DIR, "-I" + os.path.join(_DIR, "eigen_shim"), "-I" + os.path.join(_DIR, "ceres_shim"),
               "-I" + os.path.join(refroot, "include"), "-I" + os.path.join(refroot, "evaluation")]
        for name, main, extra in (("libssfm_refsolvers.so", "ref_solvers.cpp", []),
                                  ("libssfm_reffull.so", "ref_full.cpp", [os.path.join(refroot, "src", "spherical_estimator.cpp")])):
            out2 = os.path.join(_DIR, "_ref", name)
            dep = max([newest, os.path.getmtime(os.path.join(_DIR, main))] + [os.path.getmtime(x) for x in shim])
            if force or not os.path.exists(out2) or os.path.getmtime(out2) < dep:
                subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread"] + inc +
                                      ["-o", out2, os.path.join(_DIR, main)] + extra + refsrc)
        # the reference's orphan SphericalFastEstimator (src/spherical_fast_estimator.cpp) with the fast_shim stand-ins
        out4 = os.path.join(_DIR, "_ref", "libssfm_reffast.so")
        fast_dep = [os.path.join(_DIR, "ref_fast.cpp"), os.path.join(_DIR, "fast_shim", "sphericalsfm", "estimator.h"),
                    os.path.join(_DIR, "fast_shim", "Polynomial", "Polynomial.hpp")]
        dep4 = max([newest] + [os.path.getmtime(x) for x in fast_dep + shim])
        if force or not os.path.exists(out4) or os.path.getmtime(out4) < dep4:
            subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + _DIR,
                                   "-I" + os.path.join(_DIR, "fast_shim"), "-I" + os.path.join(_DIR, "eigen_shim"),
                                   "-I" + os.path.join(refroot, "include"), "-o", out4, os.path.join(_DIR, "ref_fast.cpp"),
                                   os.path.join(refroot, "src", "spherical_fast_estimator.cpp"), os.path.join(refroot, "src", "so3.cpp")])
        # config C2 as upstream wrote it: msac.h / preemptive_ransac.h around SphericalFastEstimator
        out5 = os.path.join(_DIR, "_ref", "libssfm_reflegacy.so")
        leg = [os.path.join(_DIR, f) for f in ("ref_legacy_msac.cpp", "ref_legacy_preemptive.cpp", "ref_legacy_common.hpp", "pinned_rand.hpp")]
        dep5 = max([newest] + [os.path.getmtime(x) for x in leg + fast_dep + shim])
        if force or not os.path.exists(out5) or os.path.getmtime(out5) < dep5:
            subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + _DIR,
                                   "-I" + os.path.join(_DIR, "fast_shim"), "-I" + os.path.join(_DIR, "eigen_shim"),
                                   "-I" + os.path.join(refroot, "include"), "-o", out5, leg[0], leg[1],
                                   os.path.join(refroot, "src", "spherical_fast_estimator.cpp"), os.path.join(refroot, "src", "so3.cpp")])
        # the reference's synthetic-problem generator + error metrics (evaluation/problem_generator)
        out6 = os.path.join(_DIR, "_ref", "libssfm_refgen.so")
        dep6 = max([newest, os.path.getmtime(os.path.join(_DIR, "ref_gen.cpp"))] + [os.path.getmtime(x) for x in shim])
        if force or not os.path.exists(out6) or os.path.getmtime(out6) < dep6:
            subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w"] + inc +
                                  ["-o", out6, os.path.join(_DIR, "ref_gen.cpp"),
                                   os.path.join(refroot, "evaluation", "problem_generator", "problem_generator.cpp"),
                                   os.path.join(refroot, "src", "so3.cpp"), os.path.join(refroot, "src", "spherical_utils.cpp")])
        # the reference's triangulation path (src/triangulation_estimator.cpp, sfm_types.cpp, so3.cpp + RansacLib)
        out3 = os.path.join(_DIR, "_ref", "libssfm_reftri.so")
        dep = max([newest, os.path.getmtime(os.path.join(_DIR, "ref_tri.cpp"))] + [os.path.getmtime(x) for x in shim])
        if force or not os.path.exists(out3) or os.path.getmtime(out3) < dep:
            subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread"] + inc +
                                  ["-o", out3, os.path.join(_DIR, "ref_tri.cpp")] +
                                  [os.path.join(refroot, "src", f) for f in ("triangulation_estimator.cpp", "sfm_types.cpp", "so3.cpp")])
        # the reference's LocallyOptimizedMSAC header around a toy line estimator: pins sixpt_oracle.lo_msac_generic
        out7 = os.path.join(_DIR, "_ref", "libssfm_reftoy.so")
        toy = os.path.join(_DIR, "ref_toy.cpp")
        if force or not os.path.exists(out7) or os.path.getmtime(out7) < os.path.getmtime(toy):
            subprocess.check_call([_compiler(), "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(refroot, "include"),
                                   "-I" + os.path.join(refroot, "evaluation"), "-o", out7, toy])


_cache = {}


def load():
    if "o" not in _cache:
        build(ref=False)
        _cache["o"] = Oracle(os.path.join(_DIR, "liboracle.so"))
    return _cache["o"]


class TriReference:
    """oracle/_ref/libssfm_reftri.so: only orc_triangulate (the reference's TriangulationEstimator + RansacLib)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.orc_triangulate.restype = C.c_int

    triangulate = Oracle.triangulate


class FastReference:
    """oracle/_ref/libssfm_reffast.so: the reference's SphericalFastEstimator::compute / score / decomposeE."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.orc_fast_compute.restype = C.c_int

    def compute(self, rays3):
        r = np.ascontiguousarray(rays3, np.float64).reshape(-1)
        E = np.zeros(36)
        n = self.lib.orc_fast_compute(_dp(r), _dp(E))
        return [E[9 * i:9 * i + 9].reshape(3, 3).copy() for i in range(n)]

    def score(self, E, rays):
        rays = np.ascontiguousarray(rays, np.float64)
        out = np.zeros(len(rays))
        self.lib.orc_fast_score(_dp(np.ascontiguousarray(E, np.float64).reshape(-1)), _dp(rays.reshape(-1)), len(rays), _dp(out))
        return out

    def decompose(self, E, inward=False):
        r, t = np.zeros(3), np.zeros(3)
        self.lib.orc_fast_decompose(_dp(np.ascontiguousarray(E, np.float64).reshape(-1)), int(inward), _dp(r), _dp(t))
        return r, t


class GeneratorReference:
    """oracle/_ref/libssfm_refgen.so: ProblemGenerator::make_random_problem and RelativePoseSolution's error metrics
    (evaluation/problem_generator/*).  The random engine is process-wide and seeded by default, as upstream."""

    def __init__(self, path):
        self.lib = C.CDLL(path)

    def make_random_problem(self, num_corr, inward=False, rotation_deg=-1.0, point_noise=0.0):
        rays, E, R, t = np.zeros((num_corr, 6)), np.zeros(9), np.zeros(9), np.zeros(3)
        self.lib.orc_make_random_problem(num_corr, int(inward), C.c_double(rotation_deg), C.c_double(point_noise), _dp(rays),
                                         _dp(E), _dp(R), _dp(t))
        return rays, E.reshape(3, 3), R.reshape(3, 3), t

    def errors(self, E, R, t, Es, Rs, ts):
        out = np.zeros(3)
        a = [np.ascontiguousarray(x, np.float64).reshape(-1) for x in (E, R, t, Es, Rs, ts)]
        self.lib.orc_solution_errors(*[_dp(x) for x in a], _dp(out))
        return out  # frob, rot, trans


def load_ref_gen():
    if "g" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_refgen.so")
        _cache["g"] = GeneratorReference(p) if os.path.exists(p) else None
    return _cache["g"]


class LegacyReference:
    """oracle/_ref/libssfm_reflegacy.so: the reference's MSAC / PreemptiveRANSAC drivers around its SphericalFastEstimator."""

    def __init__(self, path):
        self.lib = C.CDLL(path)

    def estimate_pair(self, rays, opt, pair_id=0):
        rays = np.ascontiguousarray(rays, np.float64)
        res = OrcResult()
        inl = np.zeros(max(len(rays), 1), np.int32)
        fn = self.lib.orc_legacy_msac if opt.driver == 2 else self.lib.orc_legacy_preemptive
        fn.restype = C.c_int
        n = fn(_dp(rays), len(rays), C.byref(opt), C.c_uint32(pair_id), C.byref(res), _ip(inl))
        return res, inl[:max(n, 0)].copy()


def load_ref_legacy():
    if "l" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_reflegacy.so")
        _cache["l"] = LegacyReference(p) if os.path.exists(p) else None
    return _cache["l"]


def load_ref_fast():
    if "s" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_reffast.so")
        _cache["s"] = FastReference(p) if os.path.exists(p) else None
    return _cache["s"]


def load_ref_tri():
    if "t" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_reftri.so")
        _cache["t"] = TriReference(p) if os.path.exists(p) else None
    return _cache["t"]


def load_ref_full():
    """oracle/_ref/libssfm_reffull.so: the reference's own RansacLib + SphericalEstimator + solver sources,
    compiled unmodified against the Eigen/Ceres stand-ins.  None when it was never built."""
    if "f" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_reffull.so")
        _cache["f"] = Oracle(p) if os.path.exists(p) else None
    return _cache["f"]


def load_ref():
    """The reference-RansacLib-driven oracle, or None when it was never built (no /root/reference)."""
    if "r" not in _cache:
        build(ref=True)
        p = os.path.join(_DIR, "_ref", "libssfm_ref.so")
        _cache["r"] = Oracle(p) if os.path.exists(p) else None
    return _cache["r"]
