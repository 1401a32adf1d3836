/* ssfm.h -- C ABI of the B200 batched spherical relative-pose engine (libssfm_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of jonathanventura/spherical-sfm: robust
 * relative pose (RANSAC / LO-RANSAC with the spherical 3-point solvers, Sampson scoring and the
 * least-squares refit) over many image pairs at once.  Plain pointers and sizes only; no C++
 * or torch types.  Every entry point cites the reference interface it replaces (paths relative
 * to the reference root).  There is NO CPU fallback: every compute entry fails with
 * SSFM_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Conventions (reference): one correspondence = RayPair = two 3-vectors (u in image 0, v in
 * image 1), 6 contiguous doubles, epipolar constraint v^T E u = 0
 * (include/sphericalsfm/ray.h:8-10, src/spherical_solvers.cpp:119).  E is row-major 3x3.
 */
#ifndef SSFM_H_
#define SSFM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SSFM_ABI_VERSION 3

typedef struct ssfm_engine* ssfm_handle;

/* Return codes of every entry point (the reference's convention is "return 0 / print";
 * include/RansacLib/ransac.h:137-139, src/spherical_solvers.cpp:105-109). */
enum {
  SSFM_OK = 0,
  SSFM_ERR_INVALID = 1,   /* bad argument */
  SSFM_ERR_NO_DEVICE = 2, /* no CUDA device / wrong architecture: the engine never falls back to the CPU */
  SSFM_ERR_CUDA = 3,      /* a CUDA call failed; ssfm_last_error() has the text */
  SSFM_ERR_OOM = 4
};

/* Per-pair status. */
enum {
  SSFM_PAIR_OK = 0,
  SSFM_PAIR_TOO_FEW_POINTS = 1, /* fewer correspondences than the minimal sample (ransac.h:137-139) */
  SSFM_PAIR_NO_MODEL = 2,
  SSFM_PAIR_SKIPPED = 3         /* fewer correspondences than SsfmOptions.min_num_points (spherical_sfm_tools.cpp:351) */
};

/* Minimal solver (SphericalEstimator's use_poly_solver flag, include/sphericalsfm/spherical_estimator.h:13-15;
 * SphericalFastEstimator, include/sphericalsfm/spherical_fast_estimator.h). */
enum {
  SSFM_SOLVER_ACTION_MATRIX = 0, /* spherical_solver_action_matrix, src/spherical_solvers.cpp:102-311 */
  SSFM_SOLVER_POLYNOMIAL = 1,    /* spherical_solver_polynomial,    src/spherical_solvers.cpp:313-660 */
  SSFM_SOLVER_FAST_STURM = 2,    /* SphericalFastEstimator::compute, src/spherical_fast_estimator.cpp:44-257 */
  SSFM_SOLVER_SIXPT_FOCAL = 3    /* SixPointEstimator (examples/six_point_estimator.{h,cpp}): six-point shared-focal
                                    relative pose, <= 15 models {t, r, focal} per sample, min_sample_size 6.  Rays are
                                    (x - cx, y - cy, 1) in pixel units.  Batched drivers: SSFM_DRIVER_VANILLA_MSAC (config C4) and
                                    SSFM_DRIVER_LO_MSAC (LocallyOptimizedMSAC with the estimator's NonMinimalSolver :121-144 and
                                    LeastSquares :146-192); the whole estimator concept is also exposed through the hooks ssfm_sixpt_solve,
                                    ssfm_score_exact and ssfm_sixpt_least_squares.
                                    SsfmPairResult.E is the matrix the estimator scores with, r/t/focal the model. */
};

/* RANSAC driver. */
enum {
  SSFM_DRIVER_LO_MSAC = 0,      /* ransac_lib::LocallyOptimizedMSAC::EstimateModel, include/RansacLib/ransac.h:128-275 */
  SSFM_DRIVER_VANILLA_MSAC = 1, /* ransac_lib::VanillaMSAC::EstimateModel,          evaluation/vanilla_ransac.h:23-99 */
  SSFM_DRIVER_MSAC_FIXED = 2,   /* sphericalsfm::MSAC::compute (fixed hypothesis budget, '<=' inlier test),
                                   include/sphericalsfm/msac.h:67-131 */
  SSFM_DRIVER_PREEMPTIVE = 3    /* sphericalsfm::PreemptiveRANSAC::compute, include/sphericalsfm/preemptive_ransac.h:46-139:
                                   fixed_budget hypotheses from selection samples of m+1 correspondences (the extra
                                   one disambiguates the roots), inlier counting in blocks of preemptive_block
                                   observations, survivors halved every preemptive_block blocks, '<=' inlier test.
                                   best_model_score = the winner's legacy MSAC cost (msac.h:56-64). */
};

/* RansacOptions + LORansacOptions, field for field (include/RansacLib/ransac.h:47-92), plus the
 * estimator's constructor arguments and the Philox key.  ssfm_default_options() fills the
 * reference defaults (ransac.h:49-73). */
typedef struct SsfmOptions {
  uint32_t min_num_iterations;      /* 100 */
  uint32_t max_num_iterations;      /* 10000 */
  double success_probability;       /* 0.9999 */
  double squared_inlier_threshold;  /* 1.0 */
  uint32_t random_seed;             /* 0; Philox key word 0, and the seed of the LO shuffle generator */
  int32_t num_lo_steps;             /* 10 (estimate_pairwise sets 0, examples/spherical_sfm_tools.cpp:316) */
  double threshold_multiplier;      /* sqrt(2) */
  int32_t num_lsq_iterations;       /* 4  (estimate_pairwise sets 0, :317) */
  int32_t min_sample_multiplicator; /* 7 */
  int32_t non_min_sample_multiplier;/* 3 */
  uint32_t lo_starting_iterations;  /* 50 */
  int32_t final_least_squares;      /* 0  (estimate_pairwise sets 1, :318) */
  int32_t solver;                   /* SSFM_SOLVER_* */
  int32_t driver;                   /* SSFM_DRIVER_* */
  int32_t inward;                   /* SphericalEstimator(.., inward) */
  int32_t fixed_budget;             /* MSAC_FIXED, PREEMPTIVE: estimators.size() (msac.h:77, preemptive_ransac.h:49) */
  double fixed_prob_success;        /* MSAC_FIXED: 0.999 (msac.h:36) */
  uint32_t first_pair_id;           /* pair p of a batch draws from Philox key (random_seed, first_pair_id + p) */
  int32_t min_num_points;           /* 0; pairs with fewer correspondences are skipped, like `m01.size() < min_num_inliers`
                                       in estimate_pairwise (examples/spherical_sfm_tools.cpp:351) */
  int32_t preemptive_block;         /* 10; PREEMPTIVE: B (preemptive_ransac.h:34,40) */
  int32_t sixpt_focal_scoring;      /* 0; SIXPT_FOCAL: 0 = score with E = skew3(t) so3exp(r) on the raw rays, exactly as
                                       SixPointEstimator::EvaluateModelOnPoint does (six_point_estimator.cpp:78-91; the
                                       focal is not used there); 1 = score with F = Kinv E Kinv, Kinv = diag(1,1,focal),
                                       the model of the reference's own refit functor (:62-70) */
  int32_t complex_root_models;      /* 0; ACTION_MATRIX: what a complex eigenvalue of the action matrix yields.  Upstream keeps the
                                       real part of Eigen's complex eigenvector (src/spherical_solvers.cpp:294-297, filter commented
                                       out); its phase is decided by rounding noise, so it cannot be matched.
                                       SSFM_COMPLEX_CANONICAL (0): the major axis of the complex solution line (deterministic,
                                       basis independent); SSFM_COMPLEX_SKIP (1): no model, i.e. upstream's filter switched on */
} SsfmOptions;
enum { SSFM_COMPLEX_CANONICAL = 0, SSFM_COMPLEX_SKIP = 1 };

/* A batch of image pairs in CSR form: pair p owns correspondences [offsets[p], offsets[p+1]).
 * `rays` is the caller's RayPairList memory (6 doubles per correspondence).  The engine never
 * keeps caller pointers after a call returns (ownership as in spherical_estimator.h:11,15). */
typedef struct SsfmBatch {
  int32_t num_pairs;
  const int64_t* offsets; /* host, num_pairs + 1 entries */
  const double* rays;     /* host pointer, or device pointer (16-byte aligned) if rays_on_device != 0 */
  int32_t rays_on_device;
  int32_t ray_format;     /* SSFM_RAYS_F64 (0): 6 doubles per correspondence, the RayPair memory.  SSFM_RAYS_F32 (1): `rays`
                             points at 6 FLOATS per correspondence (u.xyz, v.xyz) -- half the bytes over PCIe (SURVEY 8b's
                             packed-float input); they are widened to double on the device, so the result is exactly the
                             SSFM_RAYS_F64 result for the same values.  (ABI 3; the field occupies former padding: zero it.) */
} SsfmBatch;
enum { SSFM_RAYS_F64 = 0, SSFM_RAYS_F32 = 1 };

/* Per-pair output: the best model + RansacStatistics (include/RansacLib/ransac.h:94-101) + the
 * pose the callers extract afterwards with decompose_spherical_essential_matrix
 * (examples/spherical_sfm_tools.cpp:414-418; src/spherical_utils.cpp:16-66). */
typedef struct SsfmPairResult {
  double E[9];
  double r[3]; /* so3ln(R) */
  double t[3];
  double best_model_score;
  double inlier_ratio;
  uint32_t num_iterations;
  int32_t best_num_inliers;
  int32_t number_lo_iterations;
  int32_t status; /* SSFM_PAIR_* */
  int64_t evals;  /* minimal-model corr-hypothesis evaluations the reference loop would have made:
                     num_iterations * models * N */
  double focal;   /* SixPointSolution::focal (SSFM_SOLVER_SIXPT_FOCAL), 0 otherwise */
} SsfmPairResult;

/* Device-side timing/accounting of the last ssfm_run (CUDA events on the engine's stream). */
typedef struct SsfmRunStats {
  double total_ms; /* host wall clock of ssfm_run (all streams drained) */
  double pack_ms, solve_ms, score_ms, chain_ms; /* CUDA-event time on the launching streams */
  int32_t rounds;
  int32_t kernel_launches;
  int64_t evals_useful;   /* sum of SsfmPairResult.evals */
  int64_t evals_executed; /* f32 scoring evaluations actually executed (includes discarded look-ahead) */
  int64_t evals_exact;    /* f64 evaluations in the certification / LO stage */
  int64_t score_launches;
  int64_t h2d_bytes, d2h_bytes;
  int64_t refit_waves; /* walk/refit alternations of the deferred least-squares protocol */
  int32_t workers;     /* concurrent streams the pair list was split over; stage times are summed over them */
} SsfmRunStats;

int ssfm_abi_version(void);
const char* ssfm_last_error(void);
void ssfm_default_options(SsfmOptions* opt);

/* Engine lifetime.  One handle = one device + one stream; use one handle per host thread / GPU. */
int ssfm_create(int device, ssfm_handle* out);
void ssfm_destroy(ssfm_handle h);

/* Replaces the whole `#pragma omp parallel for` body of estimate_pairwise
 * (examples/spherical_sfm_tools.cpp:332-420): one LocallyOptimizedMSAC::EstimateModel per pair
 * (:380-387), the caller's inlier-mask pass (:388-392) and the pose extraction (:414-418).
 * `results`: host, num_pairs entries.  `inlier_flags`: host, one byte per correspondence of the
 * batch (1 = err < thr^2), or NULL. */
int ssfm_estimate_pairs(ssfm_handle h, const SsfmBatch* batch, const SsfmOptions* opt, SsfmPairResult* results,
                        uint8_t* inlier_flags);

/* The caller side of the path: estimate_pairwise builds every ray pair on the host as
 * Kinv * (x, y, 1) from the two images' keypoints and the pair's matches (examples/spherical_sfm_tools.cpp:
 * 357-376).  This entry takes the keypoints / matches / Kinv as they are and builds the rays on the device
 * (float64, same operation order), so 8 bytes per match cross PCIe instead of 48. */
typedef struct SsfmMatchBatch {
  int32_t num_images;
  const int64_t* keypoint_offsets; /* host, num_images + 1 */
  const float* keypoints_xy;       /* host, 2 floats per keypoint (cv::Point2f, Features::points) */
  int32_t num_pairs;
  const int32_t* pair_images;      /* host, 2 per pair: index0, index1 (ImageMatch) */
  const int64_t* match_offsets;    /* host, num_pairs + 1 */
  const int32_t* matches;          /* host, 2 per match: keypoint index in image index0, in image index1, in the
                                      iteration order of the Matches map */
  double Kinv[9];                  /* Intrinsics::getKinv(), row-major */
} SsfmMatchBatch;
int ssfm_upload_matches(ssfm_handle h, const SsfmMatchBatch* batch);
int ssfm_estimate_pairs_from_matches(ssfm_handle h, const SsfmMatchBatch* batch, const SsfmOptions* opt,
                                     SsfmPairResult* results, uint8_t* inlier_flags);

/* SfM::Retriangulate (src/sfm.cpp:156-192): one LocallyOptimizedMSAC with
 * sphericalsfm::TriangulationEstimator (src/triangulation_estimator.cpp:46-127) per 3-D point over its observations,
 * all points in one call (replaces the cv::parallel_for_ over points).  Cameras are the SfM's poses (t, r);
 * observations are pixel coordinates with the principal point removed; one focal for all (intrinsics.focal).
 * `opt` = the options Retriangulate builds (:176-178): ssfm_default_options + squared_inlier_threshold = 4,
 * final_least_squares = 1; solver/driver fields are ignored.  Outputs (host, caller-allocated): points_xyz
 * num_points x 3 (zero unless status == SSFM_PAIR_OK, like SetPoint(j, Zero) at :172), num_inliers, status
 * (SSFM_PAIR_SKIPPED: fewer than 3 observations, :173; SSFM_PAIR_NO_MODEL: fewer than 3 inliers, :186),
 * num_iterations (may be NULL). */
typedef struct SsfmTrackBatch {
  int32_t num_cameras;
  const double* camera_tr;     /* host, num_cameras x 6: Pose::t, Pose::r */
  int32_t num_points;
  const int64_t* obs_offsets;  /* host, num_points + 1 */
  const int32_t* obs_camera;   /* host, camera index of each observation */
  const double* obs_xy;        /* host, 2 per observation */
  double focal;
} SsfmTrackBatch;
int ssfm_retriangulate(ssfm_handle h, const SsfmTrackBatch* tracks, const SsfmOptions* opt, double* points_xyz,
                       int32_t* num_inliers, int32_t* status, uint32_t* num_iterations);

/* ---- one batch, N devices, one process (north_star: "image pairs shard with no cross-pair dependence across the 8 GPUs") ----
 * The reference caller is a single process whose OpenMP loop walks all i<j pairs (examples/spherical_sfm_tools.cpp:321-332).
 * ssfm_estimate_pairs_multi is that call fanned out over GPUs: contiguous shards balanced by correspondence count
 * (ssfm_partition_pairs), one host thread + one engine per device, results written straight into the caller's table.
 * There is no data-path collective, and the table is byte-identical to the single-device one (pair p keeps Philox key
 * (random_seed, first_pair_id + p) wherever it runs).  Host rays only. */
typedef struct ssfm_multi* ssfm_multi_handle;
int ssfm_multi_create(const int32_t* devices, int32_t num_devices, ssfm_multi_handle* out);
void ssfm_multi_destroy(ssfm_multi_handle m);
int32_t ssfm_multi_num_devices(ssfm_multi_handle m);
/* num_shards + 1 pair bounds: shard r = pairs [bounds[r], bounds[r+1]), equal shares of the correspondences. */
int ssfm_partition_pairs(const int64_t* offsets, int32_t num_pairs, int32_t num_shards, int32_t* bounds);
int ssfm_estimate_pairs_multi(ssfm_multi_handle m, const SsfmBatch* batch, const SsfmOptions* opt, SsfmPairResult* results,
                              uint8_t* inlier_flags);
/* Stats of device `index` in the last multi call, and the shard it ran (first_pair / num_pairs may be NULL). */
int ssfm_multi_get_stats(ssfm_multi_handle m, int32_t index, SsfmRunStats* stats, int32_t* first_pair, int32_t* num_pairs);
/* For device-side consumers of the whole table: all-gather-v of the per-pair records over NVLink/NVSwitch (NCCL, loaded
 * with dlopen on first use).  dev_tables: num_devices device pointers, each to num_pairs x SsfmPairResult in global pair
 * order, owned by the handle and valid until the next multi call. */
int ssfm_multi_allgather_results(ssfm_multi_handle m, void** dev_tables, int32_t* num_pairs);

/* ---- descriptor matching: the step in front of the relative-pose path ----
 * match() (examples/spherical_sfm_tools.cpp:235-251) for every listed image pair, as match_exhaustive() (:575-600) runs it:
 * cv::BFMatcher(NORM_L2)::knnMatch(query = features1.descs, train = features0.descs, k = 2), Lowe's ratio test
 * `m[0].distance < ratio * m[1].distance`, `m01[trainIdx] = queryIdx` in query order.  Output = the Matches maps in their
 * iteration order: per pair, (index in image 0, index in image 1) sorted by the first -- exactly what
 * SsfmMatchBatch.matches takes.  Descriptors are cv::SIFT's: 128 floats holding integers 0..255 (checked; anything else is
 * SSFM_ERR_INVALID because the fp16 tensor-core product is exact only for them); results are bit-exact with OpenCV, ties
 * included.  match_offsets: num_pairs + 1.  matches: 2 ints per match, `capacity` matches (sum of min(n0, n1) suffices). */
typedef struct SsfmDescriptorBatch {
  int32_t num_images;
  int32_t descriptor_length;    /* 128 */
  const int64_t* desc_offsets;  /* host, num_images + 1: image i owns descriptor rows [desc_offsets[i], desc_offsets[i+1]) */
  const float* descriptors;     /* host, rows x 128 (Features::descs, CV_32F) */
  int32_t num_pairs;
  const int32_t* pair_images;   /* host, 2 per pair: index0, index1 (ImageMatch) */
  double ratio;                 /* 0.75 (spherical_sfm_tools.h:70) */
} SsfmDescriptorBatch;
int ssfm_match_pairs(ssfm_handle h, const SsfmDescriptorBatch* batch, int64_t* match_offsets, int32_t* matches, int64_t capacity);
/* Stage timers (CUDA events on the engine's stream) and work counts of the last ssfm_match_pairs call on this engine. */
typedef struct SsfmMatchStats {
  double pack_ms;    /* H2D of the descriptors + k_desc_pack */
  double knn_ms;     /* k_match_2nn launches only (the tcgen05 kernel) */
  double compact_ms; /* owner table -> Matches lists + D2H */
  double total_ms;   /* whole call, host clock */
  int64_t h2d_bytes, d2h_bytes;
  int64_t distance_evaluations; /* sum over pairs of n0 * n1 */
  int64_t mma_tiles;            /* 256 x 128 x 144 tensor-core tiles executed */
  int32_t knn_launches, ctas;
} SsfmMatchStats;
int ssfm_match_get_stats(ssfm_handle h, SsfmMatchStats* out);

/* The same call split into its three stages, so inputs can stay resident in HBM:
 * upload (H2D + packing into float4 SoA), run (all kernels), download (D2H of the result table). */
int ssfm_upload(ssfm_handle h, const SsfmBatch* batch);
int ssfm_run(ssfm_handle h, const SsfmOptions* opt);
int ssfm_download(ssfm_handle h, SsfmPairResult* results, uint8_t* inlier_flags);
int ssfm_get_stats(ssfm_handle h, SsfmRunStats* stats);
/* Device pointer to the result table of the last run (num_pairs x SsfmPairResult), e.g. for an
 * NCCL all-gather without a host round trip. */
int ssfm_device_results(ssfm_handle h, void** dev_ptr, int32_t* num_pairs);

/* ---- replay hooks (parity tests drive the reference-shaped pieces one at a time) ---- */

/* The minimal sample of iteration `iter` of pair `pair`: replaces UniformSampling::Sample
 * (include/RansacLib/sampling.h:58-64).  Host function; the device code uses the same routine. */
int ssfm_sample(uint32_t seed, uint32_t pair, uint32_t iter, int32_t k, int32_t n, int32_t* idx);
/* random_sample of the legacy drivers (include/sphericalsfm/preemptive_ransac.h:8-28): Knuth 3.4.2S selection
 * sampling, k of n_total records in increasing order.  The reference draws from the C library's rand(); here
 * draw j of hypothesis h is word (j%4) of Philox(counter=(h, j/4, 1, 0), key=(seed, pair)) >> 1.  Host-side,
 * pure function. */
int ssfm_selection_sample(uint32_t seed, uint32_t pair, uint32_t hypothesis, int32_t n_total, int32_t k, int32_t* idx);

/* SphericalEstimator::MinimalSolver (src/spherical_estimator.cpp:80-84) for `num_samples`
 * samples of 3 indices into `rays` (n correspondences, host).  models: num_samples x 4 x 6 doubles
 * (p0..p5 of E = [p0 p1 p2; p1 -p0 p3; p4 p5 0], ||E||_F = 1; NaN when absent); num_models: per sample. */
int ssfm_minimal_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t num_samples,
                       int32_t solver, double* models, int32_t* num_models);
/* Same with the solver kind and complex_root_models taken from an options struct. */
int ssfm_minimal_solve_opt(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t num_samples,
                           const SsfmOptions* opt, double* models, int32_t* num_models);

/* ScoreModel + GetInliers count for many models against one pair (include/RansacLib/ransac.h:295-336,
 * EvaluateModelOnPoint src/spherical_estimator.cpp:67-78): the FP32 scoring kernel.
 * models6: num_models x 6 doubles (host).  scores: MSAC cost (float), counts: err < thr^2. */
int ssfm_score(ssfm_handle h, const double* models6, int32_t num_models, const double* rays, int32_t n,
               double squared_threshold, float* scores, int32_t* counts, float* kernel_ms);
/* The same for several pairs in ONE launch (config C5: 8 pairs x 4096 hypotheses x 10k-200k correspondences): pair p owns
 * correspondences [offsets[p], offsets[p+1]) of `rays` and the models models6[p][num_models][6]; scores / counts are
 * [num_pairs][num_models].  kernel_ms: device time of the scoring + reduction launches (after a warm-up launch). */
int ssfm_score_pairs(ssfm_handle h, const double* models6, int32_t num_models, const double* rays, const int64_t* offsets,
                     int32_t num_pairs, double squared_threshold, float* scores, int32_t* counts, float* kernel_ms);
/* Same quantities from the FP64 certification path (bit-compatible with the reference's arithmetic). */
int ssfm_score_exact(ssfm_handle h, const double* E9, int32_t num_models, const double* rays, int32_t n,
                     double squared_threshold, double* scores, int32_t* counts);

/* SphericalEstimator::LeastSquares (src/spherical_estimator.cpp:110-157) for `num_problems`
 * index sets over one pair.  sample_offsets: num_problems + 1.  E9: in/out, 9 doubles each. */
int ssfm_least_squares(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx,
                       const int32_t* sample_offsets, int32_t num_problems, int32_t inward, double* E9);

/* SphericalEstimator::NonMinimalSolver (src/spherical_estimator.cpp:86-108) for `num_problems` index
 * sets over one pair.  E9: out, 9 doubles each; ok: out, 1 if a model was produced. */
int ssfm_non_minimal_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx,
                           const int32_t* sample_offsets, int32_t num_problems, double* E9, int32_t* ok);

/* decompose_spherical_essential_matrix (src/spherical_utils.cpp:16-66), batched. */
int ssfm_decompose(ssfm_handle h, const double* E9, int32_t num, int32_t inward, double* r3, double* t3);

/* The focal search's inner step: for every trial focal f, every pair's E is rescaled
 * E' = T E T, T = diag(f/f0, f/f0, 1), and decomposed again (transform_image_matches,
 * examples/spherical_sfm_tools.cpp:1118-1131; 1024 trials x P pairs in find_best_focal_length_random).
 * scales: num_scales values of f/f0.  r3: num_scales x num x 3 (so3ln of the chosen rotation). */
int ssfm_decompose_rescaled(ssfm_handle h, const double* E9, int32_t num, const double* scales, int32_t num_scales,
                            int32_t inward, double* r3);

/* SixPointEstimator::MinimalSolver (examples/six_point_estimator.cpp:93-119) on explicit samples of six
 * indices each.  models: num_samples x 15 x 7 doubles (t[3] unit, r[3] = so3ln(R), focal), sorted by focal;
 * num_models: how many of the 15 are valid for each sample. */
int ssfm_sixpt_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples6, int32_t num_samples,
                     double* models, int32_t* num_models);

/* SixPointEstimator::LeastSquares (examples/six_point_estimator.cpp:146-192): trust-region LM over r (3), t on the unit
 * sphere (ceres::SphereManifold<3>) and the focal, residuals = the SampsonError functor (:25-76) on the listed
 * correspondences.  models7: nprob x 7 (t, r, focal), refined in place. */
int ssfm_sixpt_least_squares(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx,
                             const int32_t* sample_offsets, int32_t nprob, double* models7);

/* The LO generator: `ncalls` consecutive RandomShuffleAndResize calls (include/RansacLib/utils.h:48-52)
 * on iota vectors, one std::mt19937(seed) stream (ransac.h:143-144), executed on the device. */
int ssfm_lo_shuffle(ssfm_handle h, uint32_t seed, int32_t ncalls, const int32_t* sizes, const int32_t* targets,
                    int32_t* out);

/* Peak-FP32 microbenchmarks used as the roofline denominator: independent FFMA chains on every SM, once with the scalar
 * instruction and once with the packed one (fma.rn.f32x2 -> FFMA2).  ssfm_measure_fp32_peak returns the higher of the two. */
int ssfm_measure_fp32_peaks(ssfm_handle h, double* scalar_tflops, double* packed_tflops);
int ssfm_measure_fp32_peak(ssfm_handle h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SSFM_H_ */
