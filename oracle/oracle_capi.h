/* oracle_capi.h -- C entry points of the CPU ORACLE (test infrastructure; see ssfm_oracle.hpp).
 * The same entry points are exported by oracle/_ref/libssfm_ref.so, where the driver loop is the
 * reference's own RansacLib (include/RansacLib/ransac.h, evaluation/vanilla_ransac.h) compiled
 * from /root/reference instead of the restatement in lomsac.hpp. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint32_t min_num_iterations;
  uint32_t max_num_iterations;
  double success_probability;
  double squared_inlier_threshold;
  uint32_t random_seed;
  int32_t num_lo_steps;
  double threshold_multiplier;
  int32_t num_lsq_iterations;
  int32_t min_sample_multiplicator;
  int32_t non_min_sample_multiplier;
  uint32_t lo_starting_iterations;
  int32_t final_least_squares;
  int32_t solver_kind; /* 0 action matrix, 1 polynomial, 2 fast/Sturm */
  int32_t driver;      /* 0 LO-MSAC, 1 vanilla MSAC, 2 legacy fixed-budget MSAC */
  int32_t inward;
  int32_t legacy_budget;       /* drivers 2, 3: number of hypotheses (estimators.size()) */
  double legacy_prob_success;
  int32_t preemptive_block;    /* driver 3: B (preemptive_ransac.h:40) */
  int32_t complex_mode;        /* action-matrix solver, models from complex eigenvalues: 0 canonical, 1 Eigen-restated Re(V), 2 skip (ssfm_oracle.hpp ComplexRootMode);
                                  reference-source builds: 2 masks them in the Eigen stand-in, anything else is Eigen-restated */
} OrcOptions;

typedef struct {
  double E[9];
  double r[3];
  double t[3];
  uint32_t num_iterations;
  int32_t best_num_inliers;
  double best_model_score;
  double inlier_ratio;
  int32_t number_lo_iterations;
  int32_t status; /* 0 ok, 1 too few points, 2 no model */
  int64_t evals;  /* EvaluateModelOnPoint calls made */
} OrcResult;

int orc_is_reference(void); /* 1 in oracle/_ref (driver = the reference's RansacLib) */
void orc_philox_sample(uint32_t seed, uint32_t pair, uint32_t iter, int k, int n, int* idx);
int orc_triangulate(const double* cam_tr, const double* obs_xy, int n, double focal, const OrcOptions* opt, uint32_t point_id,
                    OrcResult* out, int* inlier_idx);
void orc_knuth_sample(uint32_t seed, uint32_t pair, uint32_t hyp, int N, int n, int* idx);
int orc_solve(const double* rays, const int* sample, int n, int kind, double* models /* 4x6 */);
int orc_solve_mode(const double* rays, const int* sample, int n, int kind, int complex_mode, double* models /* 4x6 */);
int orc_eigen34(const double* M16, double* ev /* 4 x (re,im) */, double* V /* 4x4 x (re,im) */);
void orc_sampson(const double* E9, const double* rays, int n, double* out);
void orc_score(const double* E9, const double* rays, int n, double thr, double* score, int* ninl);
void orc_decompose(const double* E9, int inward, double* r, double* t);
void orc_make_E(const double* r, int inward, double* E9);
void orc_lm_refit(const double* rays, const int* sample, int n, int inward, double* E9, int* iters, int* term,
                  double* costs /* 2 */);
void orc_lo_shuffle(uint32_t seed, int ncalls, const int* sizes, const int* targets, int* out /* sum targets */);
int orc_estimate_pair(const double* rays, int n, const OrcOptions* opt, uint32_t pair_id, OrcResult* out,
                      int* inlier_idx /* n or NULL */);
/* host threads over pairs, like the OpenMP loop at examples/spherical_sfm_tools.cpp:332; returns wall seconds. */
double orc_estimate_batch(const double* rays, const int64_t* offsets, int npairs, const OrcOptions* opt,
                          uint32_t first_pair_id, int nthreads, OrcResult* out);
double orc_estimate_batch_flags(const double* rays, const int64_t* offsets, int npairs, const OrcOptions* opt,
                                uint32_t first_pair_id, int nthreads, OrcResult* out, uint8_t* flags /* one byte per correspondence */);
double orc_score_batch(const double* models6, int nmodels, const double* rays, int n, double thr, int nthreads,
                       double* scores, int* ninl);

#ifdef __cplusplus
}
#endif
