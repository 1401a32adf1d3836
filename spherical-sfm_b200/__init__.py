"""spherical-sfm_b200 -- host-side Python mirror of the C ABI in include/ssfm.h.

The product is libssfm_b200.so (hand-written sm_100a CUDA + a thin extern "C" layer, built from
csrc/ by `build_extension()`); this module only binds it with ctypes so tests, bench.py and
Python callers can drive the batched relative-pose engine.  There is no CPU fallback: loading
fails loudly when the library is missing, and every compute call raises SsfmError when CUDA is
unavailable.

Import name: `spherical_sfm_b200` (see spherical_sfm_b200.py at the repo root; the directory
name carries the reference's hyphen).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_DIR)
LIB_PATH = os.environ.get("SSFM_LIB_PATH", os.path.join(_DIR, "libssfm_b200.so"))  # override: profiling builds


def _sources():
    """Everything libssfm_b200.so is compiled from (the staleness check looks at all of it)."""
    import glob
    return sorted(glob.glob(os.path.join(_DIR, "csrc", "*.cu")) + glob.glob(os.path.join(_DIR, "csrc", "*.cuh")) +
                  glob.glob(os.path.join(_ROOT, "include", "*.h")))

SSFM_OK, SSFM_ERR_INVALID, SSFM_ERR_NO_DEVICE, SSFM_ERR_CUDA, SSFM_ERR_OOM = 0, 1, 2, 3, 4
PAIR_OK, PAIR_TOO_FEW_POINTS, PAIR_NO_MODEL, PAIR_SKIPPED = 0, 1, 2, 3
SOLVER_ACTION_MATRIX, SOLVER_POLYNOMIAL, SOLVER_FAST_STURM, SOLVER_SIXPT_FOCAL = 0, 1, 2, 3
DRIVER_LO_MSAC, DRIVER_VANILLA_MSAC, DRIVER_MSAC_FIXED, DRIVER_PREEMPTIVE = 0, 1, 2, 3
COMPLEX_CANONICAL, COMPLEX_SKIP = 0, 1


class SsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ssfm error %d: %s" % (code, msg))
        self.code = code


class SsfmOptions(C.Structure):
    """RansacOptions + LORansacOptions (include/RansacLib/ransac.h:47-92) + estimator arguments."""
    _fields_ = [
        ("min_num_iterations", C.c_uint32), ("max_num_iterations", C.c_uint32),
        ("success_probability", C.c_double), ("squared_inlier_threshold", C.c_double),
        ("random_seed", C.c_uint32), ("num_lo_steps", C.c_int32),
        ("threshold_multiplier", C.c_double), ("num_lsq_iterations", C.c_int32),
        ("min_sample_multiplicator", C.c_int32), ("non_min_sample_multiplier", C.c_int32),
        ("lo_starting_iterations", C.c_uint32), ("final_least_squares", C.c_int32),
        ("solver", C.c_int32), ("driver", C.c_int32), ("inward", C.c_int32),
        ("fixed_budget", C.c_int32), ("fixed_prob_success", C.c_double), ("first_pair_id", C.c_uint32),
        ("min_num_points", C.c_int32), ("preemptive_block", C.c_int32), ("sixpt_focal_scoring", C.c_int32),
        ("complex_root_models", C.c_int32),
    ]


class SsfmBatch(C.Structure):
    _fields_ = [("num_pairs", C.c_int32), ("offsets", C.POINTER(C.c_int64)), ("rays", C.c_void_p),
                ("rays_on_device", C.c_int32), ("ray_format", C.c_int32)]


RAYS_F64, RAYS_F32 = 0, 1


def _ray_array(rays):
    """(contiguous array, ray_format): float32 input stays float32 (SSFM_RAYS_F32, half the bytes), anything else is float64."""
    if isinstance(rays, np.ndarray) and rays.dtype == np.float32:
        return np.ascontiguousarray(rays), RAYS_F32
    return np.ascontiguousarray(rays, np.float64), RAYS_F64


class SsfmMatchBatch(C.Structure):
    """Keypoints + matches + Kinv, as estimate_pairwise holds them (examples/spherical_sfm_tools.cpp:335-376)."""
    _fields_ = [("num_images", C.c_int32), ("keypoint_offsets", C.POINTER(C.c_int64)), ("keypoints_xy", C.POINTER(C.c_float)),
                ("num_pairs", C.c_int32), ("pair_images", C.POINTER(C.c_int32)), ("match_offsets", C.POINTER(C.c_int64)),
                ("matches", C.POINTER(C.c_int32)), ("Kinv", C.c_double * 9)]


class SsfmDescriptorBatch(C.Structure):
    """Descriptors of every image + the pair list, as match_exhaustive holds them (examples/spherical_sfm_tools.cpp:575-600)."""
    _fields_ = [("num_images", C.c_int32), ("descriptor_length", C.c_int32), ("desc_offsets", C.POINTER(C.c_int64)),
                ("descriptors", C.POINTER(C.c_float)), ("num_pairs", C.c_int32), ("pair_images", C.POINTER(C.c_int32)),
                ("ratio", C.c_double)]


class SsfmMatchStats(C.Structure):
    _fields_ = [("pack_ms", C.c_double), ("knn_ms", C.c_double), ("compact_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("distance_evaluations", C.c_int64), ("mma_tiles", C.c_int64),
                ("knn_launches", C.c_int32), ("ctas", C.c_int32)]


class SsfmPairResult(C.Structure):
    """Best model + RansacStatistics (ransac.h:94-101) + pose."""
    _fields_ = [
        ("E", C.c_double * 9), ("r", C.c_double * 3), ("t", C.c_double * 3),
        ("best_model_score", C.c_double), ("inlier_ratio", C.c_double),
        ("num_iterations", C.c_uint32), ("best_num_inliers", C.c_int32),
        ("number_lo_iterations", C.c_int32), ("status", C.c_int32), ("evals", C.c_int64), ("focal", C.c_double),
    ]


class SsfmTrackBatch(C.Structure):
    _fields_ = [
        ("num_cameras", C.c_int32), ("camera_tr", C.POINTER(C.c_double)), ("num_points", C.c_int32),
        ("obs_offsets", C.POINTER(C.c_int64)), ("obs_camera", C.POINTER(C.c_int32)), ("obs_xy", C.POINTER(C.c_double)),
        ("focal", C.c_double),
    ]


class SsfmRunStats(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double), ("pack_ms", C.c_double), ("solve_ms", C.c_double), ("score_ms", C.c_double),
        ("chain_ms", C.c_double), ("rounds", C.c_int32), ("kernel_launches", C.c_int32),
        ("evals_useful", C.c_int64), ("evals_executed", C.c_int64), ("evals_exact", C.c_int64),
        ("score_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("refit_waves", C.c_int64), ("workers", C.c_int32),
    ]


RESULT_DTYPE = np.dtype([
    ("E", np.float64, (9,)), ("r", np.float64, (3,)), ("t", np.float64, (3,)),
    ("best_model_score", np.float64), ("inlier_ratio", np.float64),
    ("num_iterations", np.uint32), ("best_num_inliers", np.int32),
    ("number_lo_iterations", np.int32), ("status", np.int32), ("evals", np.int64), ("focal", np.float64)], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(SsfmPairResult)


def build_extension(force=False, verbose=False):
    """Compile csrc/ into libssfm_b200.so for sm_100a (nvcc cross-compiles without a GPU): runs the Makefile next to this
    file -- one object per translation unit, rebuilt when it or any header changed -- which a C++ consumer can run directly."""
    cmd = ["make", "-C", _DIR, "-j", "3", "LIB=" + LIB_PATH] + (["-B"] if force else []) + \
          (["EXTRA_NVCCFLAGS=-Xptxas -v"] if verbose else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C ABI.  Raises if libssfm_b200.so has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libssfm_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is the product; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ssfm_last_error.restype = C.c_char_p
        L.ssfm_abi_version.restype = C.c_int
        for name in EXPORTED_SYMBOLS:
            getattr(L, name)
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "ssfm_abi_version", "ssfm_last_error", "ssfm_default_options", "ssfm_create", "ssfm_destroy",
    "ssfm_estimate_pairs", "ssfm_upload_matches", "ssfm_estimate_pairs_from_matches", "ssfm_upload", "ssfm_run", "ssfm_download", "ssfm_get_stats", "ssfm_device_results",
    "ssfm_sample", "ssfm_selection_sample", "ssfm_sixpt_solve", "ssfm_sixpt_least_squares", "ssfm_retriangulate", "ssfm_minimal_solve", "ssfm_minimal_solve_opt", "ssfm_score", "ssfm_score_pairs", "ssfm_score_exact", "ssfm_least_squares", "ssfm_non_minimal_solve", "ssfm_decompose", "ssfm_decompose_rescaled",
    "ssfm_lo_shuffle", "ssfm_measure_fp32_peak", "ssfm_measure_fp32_peaks",
    "ssfm_multi_create", "ssfm_multi_destroy", "ssfm_multi_num_devices", "ssfm_partition_pairs", "ssfm_estimate_pairs_multi",
    "ssfm_multi_get_stats", "ssfm_multi_allgather_results", "ssfm_match_pairs", "ssfm_match_get_stats",
]


def _check(rc):
    if rc != SSFM_OK:
        raise SsfmError(rc, lib().ssfm_last_error().decode())


def default_options(**kw):
    """RansacLib defaults (ransac.h:49-73)."""
    o = SsfmOptions()
    lib().ssfm_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def pipeline_options(squared_inlier_threshold, **kw):
    """The options estimate_pairwise uses (examples/spherical_sfm_tools.cpp:314-318)."""
    return default_options(squared_inlier_threshold=squared_inlier_threshold, num_lo_steps=0, num_lsq_iterations=0,
                           final_least_squares=1, **kw)


def sample(seed, pair, it, k, n):
    idx = np.zeros(k, np.int32)
    _check(lib().ssfm_sample(C.c_uint32(seed), C.c_uint32(pair), C.c_uint32(it), k, n,
                             idx.ctypes.data_as(C.POINTER(C.c_int32))))
    return idx


def selection_sample(seed, pair, hypothesis, n_total, k):
    """random_sample of the legacy drivers (preemptive_ransac.h:8-28) on the Philox-backed rand()."""
    idx = np.zeros(k, np.int32)
    _check(lib().ssfm_selection_sample(C.c_uint32(seed), C.c_uint32(pair), C.c_uint32(hypothesis), n_total, k,
                                       idx.ctypes.data_as(C.POINTER(C.c_int32))))
    return idx


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Engine:
    """One handle = one GPU + one stream (include/ssfm.h)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().ssfm_create(int(device), C.byref(self._h)))
        self.device = device
        self._num_pairs = 0
        self._num_corr = 0

    def close(self):
        if self._h:
            lib().ssfm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the batched entry point --------------------------------------------------------
    def upload(self, rays, offsets, device_ptr=None, device_format=RAYS_F64):
        """rays: (M, 6) float64 host array (RayPair memory) or float32 (SSFM_RAYS_F32), or device_ptr=int for rays already
        in HBM (device_format says which of the two layouts they are in)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        b = SsfmBatch()
        b.num_pairs = len(offsets) - 1
        b.offsets = _p(offsets, C.c_int64)
        if device_ptr is not None:
            b.rays = C.c_void_p(int(device_ptr))
            b.rays_on_device = 1
            b.ray_format = device_format
        else:
            rays, b.ray_format = _ray_array(rays)
            assert rays.size == 6 * int(offsets[-1])
            self._keep = rays
            b.rays = C.c_void_p(rays.ctypes.data)
            b.rays_on_device = 0
        _check(lib().ssfm_upload(self._h, C.byref(b)))
        self._num_pairs = b.num_pairs
        self._num_corr = int(offsets[-1])
        self._keep = None

    def run(self, opt):
        _check(lib().ssfm_run(self._h, C.byref(opt)))

    def download(self, want_flags=True):
        res = np.zeros(self._num_pairs, RESULT_DTYPE)
        flags = np.zeros(self._num_corr, np.uint8) if want_flags else None
        _check(lib().ssfm_download(self._h, C.c_void_p(res.ctypes.data),
                                   C.c_void_p(flags.ctypes.data) if want_flags and self._num_corr else None))
        return res, flags

    def estimate_pairs(self, rays, offsets, opt, want_flags=True, out_results=None, out_flags=None):
        """ssfm_estimate_pairs: upload + run + download in one call (host buffers in, host buffers out).
        out_results / out_flags: optional preallocated (e.g. pinned) output arrays."""
        rays, fmt = _ray_array(rays)
        offsets = np.ascontiguousarray(offsets, np.int64)
        b = SsfmBatch(len(offsets) - 1, _p(offsets, C.c_int64), C.c_void_p(rays.ctypes.data), 0, fmt)
        res = out_results if out_results is not None else np.zeros(b.num_pairs, RESULT_DTYPE)
        assert res.dtype == RESULT_DTYPE and len(res) == b.num_pairs
        flags = (out_flags if out_flags is not None else np.zeros(int(offsets[-1]), np.uint8)) if want_flags else None
        _check(lib().ssfm_estimate_pairs(self._h, C.byref(b), C.byref(opt), C.c_void_p(res.ctypes.data),
                                         C.c_void_p(flags.ctypes.data) if want_flags and len(flags) else None))
        self._num_pairs = b.num_pairs
        self._num_corr = int(offsets[-1])
        return res, flags

    def estimate_pairs_from_matches(self, keypoints_xy, keypoint_offsets, pair_images, matches, match_offsets, Kinv, opt,
                                    want_flags=True, out_results=None, out_flags=None):
        """ssfm_estimate_pairs_from_matches: rays are built on the device from keypoints, matches and Kinv."""
        kp = np.ascontiguousarray(keypoints_xy, np.float32).reshape(-1, 2)
        kpo = np.ascontiguousarray(keypoint_offsets, np.int64)
        pi = np.ascontiguousarray(pair_images, np.int32).reshape(-1, 2)
        mt = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        mo = np.ascontiguousarray(match_offsets, np.int64)
        b = SsfmMatchBatch()
        b.num_images = len(kpo) - 1
        b.keypoint_offsets = _p(kpo, C.c_int64)
        b.keypoints_xy = _p(kp, C.c_float)
        b.num_pairs = len(pi)
        b.pair_images = _p(pi, C.c_int32)
        b.match_offsets = _p(mo, C.c_int64)
        b.matches = _p(mt, C.c_int32)
        b.Kinv = (C.c_double * 9)(*np.asarray(Kinv, np.float64).reshape(9))
        res = out_results if out_results is not None else np.zeros(b.num_pairs, RESULT_DTYPE)
        flags = (out_flags if out_flags is not None else np.zeros(int(mo[-1]), np.uint8)) if want_flags else None
        _check(lib().ssfm_estimate_pairs_from_matches(self._h, C.byref(b), C.byref(opt), C.c_void_p(res.ctypes.data),
                                                      C.c_void_p(flags.ctypes.data) if want_flags and len(flags) else None))
        self._num_pairs = b.num_pairs
        self._num_corr = int(mo[-1])
        return res, flags

    def match_pairs(self, descriptors, desc_offsets, pair_images, ratio=0.75):
        """ssfm_match_pairs: BF 2-NN + ratio test for every pair.  descriptors (rows, 128) float32 (cv::SIFT integers),
        desc_offsets (num_images + 1), pair_images (P, 2).  Returns (match_offsets (P + 1), matches (M, 2))."""
        d = np.ascontiguousarray(descriptors, np.float32).reshape(-1, 128)
        do = np.ascontiguousarray(desc_offsets, np.int64)
        pi = np.ascontiguousarray(pair_images, np.int32).reshape(-1, 2)
        n = np.diff(do)
        cap = int(np.minimum(n[pi[:, 0]], n[pi[:, 1]]).sum()) if len(pi) else 0
        b = SsfmDescriptorBatch(len(do) - 1, 128, _p(do, C.c_int64), _p(d, C.c_float), len(pi), _p(pi, C.c_int32), float(ratio))
        offs = np.zeros(len(pi) + 1, np.int64)
        out = np.empty((max(cap, 1), 2), np.int32)  # capacity, not content: only the first offs[-1] rows are written and returned
        _check(lib().ssfm_match_pairs(self._h, C.byref(b), _p(offs, C.c_int64), _p(out, C.c_int32), C.c_int64(cap)))
        return offs, out[:int(offs[-1])]

    def match_stats(self):
        s = SsfmMatchStats()
        _check(lib().ssfm_match_get_stats(self._h, C.byref(s)))
        return s

    def stats(self):
        s = SsfmRunStats()
        _check(lib().ssfm_get_stats(self._h, C.byref(s)))
        return s

    def device_results(self):
        ptr = C.c_void_p()
        n = C.c_int32()
        _check(lib().ssfm_device_results(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    # ---- replay hooks -------------------------------------------------------------------
    def minimal_solve(self, rays, samples, solver=SOLVER_ACTION_MATRIX, complex_root_models=COMPLEX_CANONICAL):
        rays = np.ascontiguousarray(rays, np.float64)
        samples = np.ascontiguousarray(samples, np.int32).reshape(-1, 3)
        ns = len(samples)
        models = np.zeros((ns, 4, 6))
        nm = np.zeros(ns, np.int32)
        opt = default_options(solver=solver, complex_root_models=complex_root_models)
        _check(lib().ssfm_minimal_solve_opt(self._h, _p(rays, C.c_double), len(rays), _p(samples, C.c_int32), ns, C.byref(opt),
                                            _p(models, C.c_double), _p(nm, C.c_int32)))
        return models, nm

    def retriangulate(self, camera_tr, obs_offsets, obs_camera, obs_xy, focal, opt):
        """ssfm_retriangulate: SfM::Retriangulate for all points.  Returns (points[P,3], num_inliers, status, iterations)."""
        cam = np.ascontiguousarray(camera_tr, np.float64).reshape(-1, 6)
        offs = np.ascontiguousarray(obs_offsets, np.int64)
        oc = np.ascontiguousarray(obs_camera, np.int32)
        oxy = np.ascontiguousarray(obs_xy, np.float64).reshape(-1, 2)
        P = len(offs) - 1
        b = SsfmTrackBatch(len(cam), _p(cam, C.c_double), P, _p(offs, C.c_int64), _p(oc, C.c_int32), _p(oxy, C.c_double), float(focal))
        pts = np.zeros((max(P, 1), 3))
        ninl = np.zeros(max(P, 1), np.int32)
        status = np.zeros(max(P, 1), np.int32)
        iters = np.zeros(max(P, 1), np.uint32)
        _check(lib().ssfm_retriangulate(self._h, C.byref(b), C.byref(opt), _p(pts, C.c_double), _p(ninl, C.c_int32),
                                        _p(status, C.c_int32), _p(iters, C.c_uint32)))
        return pts[:P], ninl[:P], status[:P], iters[:P]

    def sixpt_solve(self, rays, samples6):
        """SixPointEstimator::MinimalSolver on explicit samples: (models[ns,15,7] = t, r, focal; counts[ns])."""
        rays = np.ascontiguousarray(rays, np.float64)
        samples = np.ascontiguousarray(samples6, np.int32).reshape(-1, 6)
        ns = len(samples)
        models = np.zeros((ns, 15, 7))
        nm = np.zeros(ns, np.int32)
        _check(lib().ssfm_sixpt_solve(self._h, _p(rays, C.c_double), len(rays), _p(samples, C.c_int32), ns,
                                      _p(models, C.c_double), _p(nm, C.c_int32)))
        return models, nm

    def sixpt_least_squares(self, rays, samples, models7):
        """SixPointEstimator::LeastSquares on lists of correspondence indices; models7 (n,7) = t, r, focal."""
        rays = np.ascontiguousarray(rays, np.float64)
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in samples]), np.int32)
        offs = np.concatenate([[0], np.cumsum([len(s) for s in samples])]).astype(np.int32)
        m = np.ascontiguousarray(models7, np.float64).reshape(-1, 7).copy()
        _check(lib().ssfm_sixpt_least_squares(self._h, _p(rays, C.c_double), len(rays), _p(idx, C.c_int32), _p(offs, C.c_int32),
                                              len(samples), _p(m, C.c_double)))
        return m

    def score(self, models6, rays, thr2):
        models6 = np.ascontiguousarray(models6, np.float64).reshape(-1, 6)
        rays = np.ascontiguousarray(rays, np.float64)
        M = len(models6)
        scores = np.zeros(M, np.float32)
        counts = np.zeros(M, np.int32)
        ms = C.c_float()
        _check(lib().ssfm_score(self._h, _p(models6, C.c_double), M, _p(rays, C.c_double), len(rays), C.c_double(thr2),
                                _p(scores, C.c_float), _p(counts, C.c_int32), C.byref(ms)))
        return scores, counts, ms.value

    def score_pairs(self, models6, rays, offsets, thr2):
        """ssfm_score_pairs: models6 (P, M, 6) against P pairs (CSR offsets) in one launch -> (scores (P, M), counts, ms)."""
        models6 = np.ascontiguousarray(models6, np.float64)
        Pn, M = models6.shape[0], models6.shape[1]
        rays = np.ascontiguousarray(rays, np.float64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        assert len(offsets) == Pn + 1
        scores = np.zeros((Pn, M), np.float32)
        counts = np.zeros((Pn, M), np.int32)
        ms = C.c_float()
        _check(lib().ssfm_score_pairs(self._h, _p(models6, C.c_double), M, _p(rays, C.c_double), _p(offsets, C.c_int64), Pn,
                                      C.c_double(thr2), _p(scores, C.c_float), _p(counts, C.c_int32), C.byref(ms)))
        return scores, counts, ms.value

    def score_exact(self, E9, rays, thr2):
        E9 = np.ascontiguousarray(E9, np.float64).reshape(-1, 9)
        rays = np.ascontiguousarray(rays, np.float64)
        M = len(E9)
        scores = np.zeros(M)
        counts = np.zeros(M, np.int32)
        _check(lib().ssfm_score_exact(self._h, _p(E9, C.c_double), M, _p(rays, C.c_double), len(rays), C.c_double(thr2),
                                      _p(scores, C.c_double), _p(counts, C.c_int32)))
        return scores, counts

    def least_squares(self, rays, samples, E9, inward=False):
        """samples: list of index arrays; E9: (len(samples), 9) initial models.  Returns refined (.., 9)."""
        rays = np.ascontiguousarray(rays, np.float64)
        offs = np.zeros(len(samples) + 1, np.int32)
        offs[1:] = np.cumsum([len(s) for s in samples])
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in samples]) if len(samples) else
                                   np.zeros(0, np.int32), np.int32)
        if idx.size == 0:
            idx = np.zeros(1, np.int32)
        E = np.array(E9, np.float64).reshape(-1, 9).copy()
        _check(lib().ssfm_least_squares(self._h, _p(rays, C.c_double), len(rays), _p(idx, C.c_int32), _p(offs, C.c_int32),
                                        len(samples), int(inward), _p(E, C.c_double)))
        return E

    def non_minimal_solve(self, rays, samples):
        rays = np.ascontiguousarray(rays, np.float64)
        offs = np.zeros(len(samples) + 1, np.int32)
        offs[1:] = np.cumsum([len(s) for s in samples])
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s, np.int32) for s in samples]), np.int32)
        E = np.zeros((len(samples), 9))
        ok = np.zeros(len(samples), np.int32)
        _check(lib().ssfm_non_minimal_solve(self._h, _p(rays, C.c_double), len(rays), _p(idx, C.c_int32),
                                            _p(offs, C.c_int32), len(samples), _p(E, C.c_double), _p(ok, C.c_int32)))
        return E, ok

    def decompose(self, E9, inward=False):
        E = np.ascontiguousarray(E9, np.float64).reshape(-1, 9)
        r = np.zeros((len(E), 3))
        t = np.zeros((len(E), 3))
        _check(lib().ssfm_decompose(self._h, _p(E, C.c_double), len(E), int(inward), _p(r, C.c_double), _p(t, C.c_double)))
        return r, t

    def decompose_rescaled(self, E9, scales, inward=False):
        """transform_image_matches (examples/spherical_sfm_tools.cpp:1118-1131): r of T E T for every scale."""
        E = np.ascontiguousarray(E9, np.float64).reshape(-1, 9)
        sc = np.ascontiguousarray(scales, np.float64)
        r = np.zeros((len(sc), len(E), 3))
        _check(lib().ssfm_decompose_rescaled(self._h, _p(E, C.c_double), len(E), _p(sc, C.c_double), len(sc), int(inward),
                                             _p(r, C.c_double)))
        return r

    def lo_shuffle(self, seed, sizes, targets):
        sizes = np.ascontiguousarray(sizes, np.int32)
        targets = np.ascontiguousarray(targets, np.int32)
        out = np.zeros(max(int(targets.sum()), 1), np.int32)
        _check(lib().ssfm_lo_shuffle(self._h, C.c_uint32(seed), len(sizes), _p(sizes, C.c_int32), _p(targets, C.c_int32),
                                     _p(out, C.c_int32)))
        return out[:int(targets.sum())]

    def measure_fp32_peak(self):
        t = C.c_double()
        _check(lib().ssfm_measure_fp32_peak(self._h, C.byref(t)))
        return t.value

    def measure_fp32_peaks(self):
        """(scalar FFMA chain, packed FFMA2 chain) in TFLOP/s"""
        a, b = C.c_double(), C.c_double()
        _check(lib().ssfm_measure_fp32_peaks(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def partition_pairs(offsets, num_shards):
    """ssfm_partition_pairs: num_shards + 1 pair bounds, equal shares of the correspondences (host function)."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    bounds = np.zeros(num_shards + 1, np.int32)
    _check(lib().ssfm_partition_pairs(_p(offsets, C.c_int64), len(offsets) - 1, int(num_shards), _p(bounds, C.c_int32)))
    return bounds.tolist()


class MultiEngine:
    """One batch, N devices, one process (ssfm_estimate_pairs_multi)."""

    def __init__(self, devices):
        devs = np.ascontiguousarray(list(devices), np.int32)
        self._h = C.c_void_p()
        self.devices = devs.tolist()
        _check(lib().ssfm_multi_create(_p(devs, C.c_int32), len(devs), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ssfm_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def estimate_pairs(self, rays, offsets, opt, want_flags=True, out_results=None, out_flags=None):
        rays, fmt = _ray_array(rays)
        offsets = np.ascontiguousarray(offsets, np.int64)
        b = SsfmBatch(len(offsets) - 1, _p(offsets, C.c_int64), C.c_void_p(rays.ctypes.data), 0, fmt)
        res = out_results if out_results is not None else np.zeros(b.num_pairs, RESULT_DTYPE)
        flags = (out_flags if out_flags is not None else np.zeros(int(offsets[-1]), np.uint8)) if want_flags else None
        _check(lib().ssfm_estimate_pairs_multi(self._h, C.byref(b), C.byref(opt), C.c_void_p(res.ctypes.data),
                                               C.c_void_p(flags.ctypes.data) if want_flags and len(flags) else None))
        return res, flags

    def stats(self, index):
        s = SsfmRunStats()
        p0, n = C.c_int32(), C.c_int32()
        _check(lib().ssfm_multi_get_stats(self._h, int(index), C.byref(s), C.byref(p0), C.byref(n)))
        return s, p0.value, n.value

    def allgather_results(self):
        """Device pointers (one per device) to the whole result table in global pair order, and the pair count."""
        ptrs = (C.c_void_p * len(self.devices))()
        n = C.c_int32()
        _check(lib().ssfm_multi_allgather_results(self._h, ptrs, C.byref(n)))
        return [p for p in ptrs], n.value


from . import problems  # noqa: E402,F401
from . import sharding  # noqa: E402,F401
