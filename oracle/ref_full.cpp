// ref_full.cpp -- oracle/_ref/libssfm_reffull.so: the reference's OWN robust-estimation path, compiled
// unmodified from /root/reference where it lies:
//     include/RansacLib/{ransac,sampling,utils}.h, evaluation/vanilla_ransac.h       (drivers)
//     src/spherical_estimator.cpp  (SphericalEstimator: MinimalSolver, EvaluateModelOnPoint,
//                                   NonMinimalSolver, LeastSquares with its autodiff'd SampsonError)
//     src/spherical_solvers.cpp, src/so3.cpp, src/spherical_utils.cpp
// Only the third-party libraries it needs are substituted: Eigen by oracle/eigen_shim and Ceres by
// oracle/ceres_shim (neither is installed here).  The sampler is the Philox Sampler (RansacLib's own
// template parameter).  TEST INFRASTRUCTURE ONLY.  It exports the same C entry points as
// oracle_capi.cpp so tests can put it beside the restated oracle and the GPU engine.
// (Not used for timing: the Eigen stand-in is heap-backed and would misrepresent the reference's speed.)
#include <RansacLib/ransac.h>
#include <sphericalsfm/so3.h>
#include <sphericalsfm/spherical_estimator.h>
#include <sphericalsfm/spherical_solvers.h>
#include <sphericalsfm/spherical_utils.h>
#include <vanilla_ransac.h>

#include <chrono>
#include <cstring>

#include "lomsac.hpp"
#include "oracle_capi.h"
#include "ssfm_oracle.hpp"

using namespace sphericalsfm;

namespace {

// Adds what the batched harness needs (a pair id for the Philox key, an evaluation counter)
// without touching the reference class.
class CountingEstimator : public SphericalEstimator {
 public:
  typedef Eigen::Matrix3d Model;
  typedef std::vector<Eigen::Matrix3d> ModelVector;
  CountingEstimator(const RayPairList& c, bool poly, bool inward, uint32_t id) : SphericalEstimator(c, poly, inward), id_(id) {}
  uint32_t pair_id() const { return id_; }
  double EvaluateModelOnPoint(const Eigen::Matrix3d& E, int i) const {
    ++evals_;
    return SphericalEstimator::EvaluateModelOnPoint(E, i);
  }
  mutable long long evals_ = 0;

 private:
  uint32_t id_;
};

RayPairList to_list(const double* rays, int n) {
  RayPairList c(n);
  for (int i = 0; i < n; ++i) {
    c[i].first = Eigen::Vector3d(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    c[i].second = Eigen::Vector3d(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
  }
  return c;
}

Eigen::Matrix3d to_mat(const double* E9) {
  Eigen::Matrix3d E;
  for (int i = 0; i < 9; ++i) E.d[i] = E9[i];
  return E;
}

int estimate_one(const double* rays, int n, const OrcOptions& o, uint32_t pair_id, OrcResult* out, int* inlier_idx) {
  ssfm_oracle::eigen_shim_skip_complex() = o.complex_mode == ssfm_oracle::COMPLEX_SKIP;  // single-threaded harness
  const RayPairList corr = to_list(rays, n);
  CountingEstimator est(corr, o.solver_kind == 1, o.inward != 0, pair_id);
  ransac_lib::LORansacOptions ro;
  ro.min_num_iterations_ = o.min_num_iterations;
  ro.max_num_iterations_ = o.max_num_iterations;
  ro.success_probability_ = o.success_probability;
  ro.squared_inlier_threshold_ = o.squared_inlier_threshold;
  ro.random_seed_ = o.random_seed;
  ro.num_lo_steps_ = o.num_lo_steps;
  ro.threshold_multiplier_ = o.threshold_multiplier;
  ro.num_lsq_iterations_ = o.num_lsq_iterations;
  ro.min_sample_multiplicator_ = o.min_sample_multiplicator;
  ro.non_min_sample_multiplier_ = o.non_min_sample_multiplier;
  ro.lo_starting_iterations_ = o.lo_starting_iterations;
  ro.final_least_squares_ = o.final_least_squares != 0;
  ransac_lib::RansacStatistics rs;
  Eigen::Matrix3d E;
  typedef ssfm_oracle::PhiloxSampling<CountingEstimator> Sampler;
  if (o.driver == 1) {
    ransac_lib::VanillaMSAC<Eigen::Matrix3d, std::vector<Eigen::Matrix3d>, CountingEstimator, Sampler> ransac;
    ransac.EstimateModel(ro, est, &E, &rs);
  } else {
    ransac_lib::LocallyOptimizedMSAC<Eigen::Matrix3d, std::vector<Eigen::Matrix3d>, CountingEstimator, Sampler> ransac;
    ransac.EstimateModel(ro, est, &E, &rs);
  }
  for (int i = 0; i < 9; ++i) out->E[i] = E.d.size() == 9 ? E.d[i] : 0.0;
  out->num_iterations = rs.num_iterations;
  out->best_num_inliers = rs.best_num_inliers;
  out->best_model_score = rs.best_model_score;
  out->inlier_ratio = rs.inlier_ratio;
  out->number_lo_iterations = rs.number_lo_iterations;
  out->evals = est.evals_;
  for (int i = 0; i < 3; ++i) out->r[i] = out->t[i] = 0.0;
  if (n < 3) {
    out->status = 1;
  } else if (!(rs.best_model_score < std::numeric_limits<double>::max())) {
    out->status = 2;
  } else {
    out->status = 0;
    Eigen::Vector3d r, t;
    decompose_spherical_essential_matrix(E, o.inward != 0, r, t);  // examples/spherical_sfm_tools.cpp:414-418
    for (int i = 0; i < 3; ++i) { out->r[i] = r(i); out->t[i] = t(i); }
  }
  if (inlier_idx)
    for (size_t i = 0; i < rs.inlier_indices.size(); ++i) inlier_idx[i] = rs.inlier_indices[i];
  return rs.best_num_inliers;
}

}  // namespace

extern "C" {

int orc_is_reference(void) { return 2; }

void orc_philox_sample(uint32_t seed, uint32_t pair, uint32_t iter, int k, int n, int* idx) {
  ssfm_oracle::philox_sample(seed, pair, iter, k, n, idx);
}

int orc_solve(const double* rays, const int* sample, int n, int kind, double* models);
int orc_solve_mode(const double* rays, const int* sample, int n, int kind, int complex_mode, double* models) {
  ssfm_oracle::eigen_shim_skip_complex() = complex_mode == ssfm_oracle::COMPLEX_SKIP;
  const int nm = orc_solve(rays, sample, n, kind, models);
  ssfm_oracle::eigen_shim_skip_complex() = 0;
  return nm;
}
int orc_solve(const double* rays, const int* sample, int n, int kind, double* models) {
  int mx = 0;
  for (int i = 0; i < n; ++i) mx = std::max(mx, sample[i]);
  const RayPairList corr = to_list(rays, mx + 1);
  std::vector<int> s(sample, sample + n);
  std::vector<Eigen::Matrix3d> Es;
  const int nm = kind == 1 ? spherical_solver_polynomial(corr, s, &Es) : spherical_solver_action_matrix(corr, s, &Es);
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 6; ++i) models[6 * k + i] = std::numeric_limits<double>::quiet_NaN();
  for (int k = 0; k < nm && k < 4; ++k) {
    const Eigen::Matrix3d& E = Es[k];
    const double p[6] = {E(0, 0), E(0, 1), E(0, 2), E(1, 2), E(2, 0), E(2, 1)};
    std::memcpy(models + 6 * k, p, sizeof(p));
  }
  return nm;
}

void orc_sampson(const double* E9, const double* rays, int n, double* out) {
  const RayPairList corr = to_list(rays, n);
  SphericalEstimator est(corr, false, false);
  const Eigen::Matrix3d E = to_mat(E9);
  for (int i = 0; i < n; ++i) out[i] = est.EvaluateModelOnPoint(E, i);
}

void orc_score(const double* E9, const double* rays, int n, double thr, double* score, int* ninl) {
  const RayPairList corr = to_list(rays, n);
  SphericalEstimator est(corr, false, false);
  const Eigen::Matrix3d E = to_mat(E9);
  double s = 0.0;
  int c = 0;
  for (int i = 0; i < n; ++i) {
    const double e = est.EvaluateModelOnPoint(E, i);
    s += std::min(e, thr);
    c += e < thr;
  }
  *score = s;
  *ninl = c;
}

void orc_decompose(const double* E9, int inward, double* r, double* t) {
  Eigen::Vector3d rr, tt;
  decompose_spherical_essential_matrix(to_mat(E9), inward != 0, rr, tt);
  for (int i = 0; i < 3; ++i) { r[i] = rr(i); t[i] = tt(i); }
}

void orc_make_E(const double* r, int inward, double* E9) {
  Eigen::Matrix3d E;
  make_spherical_essential_matrix(so3exp(Eigen::Vector3d(r[0], r[1], r[2])), inward != 0, E);
  for (int i = 0; i < 9; ++i) E9[i] = E.d[i];
}

void orc_lm_refit(const double* rays, const int* sample, int n, int inward, double* E9, int* iters, int* term, double* costs) {
  int mx = 0;
  for (int i = 0; i < n; ++i) mx = std::max(mx, sample[i]);
  const RayPairList corr = to_list(rays, mx + 1);
  SphericalEstimator est(corr, false, inward != 0);
  Eigen::Matrix3d E = to_mat(E9);
  est.LeastSquares(std::vector<int>(sample, sample + n), &E);  // src/spherical_estimator.cpp:110-157
  for (int i = 0; i < 9; ++i) E9[i] = E.d[i];
  if (iters) *iters = -1;
  if (term) *term = -1;
  if (costs) { costs[0] = costs[1] = 0.0; }
}

void orc_lo_shuffle(uint32_t seed, int ncalls, const int* sizes, const int* targets, int* out) {
  std::mt19937 rng;
  rng.seed(seed);
  int o = 0;
  for (int c = 0; c < ncalls; ++c) {
    std::vector<int> v(sizes[c]);
    for (int i = 0; i < sizes[c]; ++i) v[i] = i;
    ransac_lib::utils::RandomShuffleAndResize(targets[c], &rng, &v);  // include/RansacLib/utils.h:48-52
    for (int i = 0; i < targets[c]; ++i) out[o++] = v[i];
  }
}

int orc_estimate_pair(const double* rays, int n, const OrcOptions* opt, uint32_t pair_id, OrcResult* out, int* inlier_idx) {
  return estimate_one(rays, n, *opt, pair_id, out, inlier_idx);
}

double orc_estimate_batch(const double* rays, const int64_t* offsets, int npairs, const OrcOptions* opt, uint32_t first_pair_id,
                          int /*nthreads*/, OrcResult* out) {
  const auto t0 = std::chrono::steady_clock::now();
  for (int p = 0; p < npairs; ++p)
    estimate_one(rays + 6 * offsets[p], (int)(offsets[p + 1] - offsets[p]), *opt, first_pair_id + (uint32_t)p, &out[p], nullptr);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double orc_score_batch(const double* models6, int nmodels, const double* rays, int n, double thr, int, double* scores, int* ninl) {
  const auto t0 = std::chrono::steady_clock::now();
  for (int m = 0; m < nmodels; ++m) {
    const double* p = models6 + 6 * (size_t)m;
    const double E9[9] = {p[0], p[1], p[2], p[1], -p[0], p[3], p[4], p[5], 0.0};
    orc_score(E9, rays, n, thr, &scores[m], &ninl[m]);
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
