// ssfm_ransaclib.hpp -- header-only C++ adapters over the C ABI (ssfm.h) that keep RansacLib's
// estimator concept and driver signature, so the reference pipeline can call the B200 engine as a
// drop-in for this path:
//
//   reference type                                            replacement here
//   sphericalsfm::SphericalEstimator                          ssfm_b200::GpuSphericalEstimator<Matrix3>
//     (include/sphericalsfm/spherical_estimator.h:8-33)
//   ransac_lib::LocallyOptimizedMSAC<M,MV,Solver>::EstimateModel   ssfm_b200::LocallyOptimizedMSAC<M,MV,Solver>::EstimateModel
//     (include/RansacLib/ransac.h:128-129)                           (same signature; one GPU call per pair)
//   ransac_lib::VanillaMSAC (evaluation/vanilla_ransac.h:23)   ssfm_b200::VanillaMSAC
//   sphericalsfm::MSAC<List,Est>::compute (msac.h:67-131)     ssfm_b200::MSAC<List,Est>::compute            (legacy drivers of the
//   sphericalsfm::PreemptiveRANSAC<List,Est>::compute         ssfm_b200::PreemptiveRANSAC<List,Est>::compute  Sturm-variant estimator;
//     (preemptive_ransac.h:46-139)                            + ssfm_b200::GpuSphericalFastEstimator           same signatures)
//   sphericalsfmtools::SixPointEstimator / SixPointSolution    ssfm_b200::GpuSixPointEstimator / SixPointSolution
//     (examples/six_point_estimator.h:9-37)                      (VanillaMSAC<SixPointSolution, ..., GpuSixPointEstimator>)
//   SfM::Retriangulate (src/sfm.cpp:156-192)                  ssfm_b200::Retriangulate (ONE call for all points)
//   the `#pragma omp parallel for` over pairs                  ssfm_b200::EstimatePairs (ONE call for all pairs)
//     (examples/spherical_sfm_tools.cpp:332-420)
//
// Matrix3 is any 3x3 type with `double& operator()(int,int)` (Eigen::Matrix3d works unchanged);
// RayPairList is any contiguous container of {Vector3d first, second} (48 bytes per element), i.e.
// sphericalsfm::RayPairList (include/sphericalsfm/ray.h:8-10).  Options / statistics types are duck
// typed on RansacLib's field names, so ransac_lib::LORansacOptions / RansacStatistics can be passed
// as they are; POD mirrors are provided for builds without RansacLib.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "ssfm.h"

namespace ssfm_b200 {

struct Mat3d {  // stand-in for Eigen::Matrix3d where Eigen is absent (row-major)
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};

struct RansacOptions {  // include/RansacLib/ransac.h:47-60
  uint32_t min_num_iterations_ = 100u;
  uint32_t max_num_iterations_ = 10000u;
  double success_probability_ = 0.9999;
  double squared_inlier_threshold_ = 1.0;
  unsigned int random_seed_ = 0u;
};
struct LORansacOptions : public RansacOptions {  // ransac.h:64-92
  int num_lo_steps_ = 10;
  double threshold_multiplier_ = std::sqrt(2.0);
  int num_lsq_iterations_ = 4;
  int min_sample_multiplicator_ = 7;
  int non_min_sample_multiplier_ = 3;
  uint32_t lo_starting_iterations_ = 50u;
  bool final_least_squares_ = false;
};
struct RansacStatistics {  // ransac.h:94-101
  uint32_t num_iterations = 0;
  int best_num_inliers = 0;
  double best_model_score = std::numeric_limits<double>::max();
  double inlier_ratio = 0.0;
  std::vector<int> inlier_indices;
  int number_lo_iterations = 0;
};

class Error : public std::runtime_error {
 public:
  Error(int code, const char* what) : std::runtime_error(std::string("ssfm: ") + what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};
inline void check(int rc) {
  if (rc != SSFM_OK) throw Error(rc, ssfm_last_error());
}

// RAII handle: one GPU, one stream.
class Engine {
 public:
  explicit Engine(int device = 0) { check(ssfm_create(device, &h_)); }
  ~Engine() { ssfm_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  ssfm_handle get() const { return h_; }

 private:
  ssfm_handle h_ = nullptr;
};

// RAII handle over several GPUs of one box (ssfm_multi_create): one engine per device, one batch per call.
class MultiEngine {
 public:
  explicit MultiEngine(const std::vector<int>& devices) {
    std::vector<int32_t> d(devices.begin(), devices.end());
    check(ssfm_multi_create(d.data(), (int32_t)d.size(), &h_));
  }
  ~MultiEngine() { ssfm_multi_destroy(h_); }
  MultiEngine(const MultiEngine&) = delete;
  MultiEngine& operator=(const MultiEngine&) = delete;
  ssfm_multi_handle get() const { return h_; }
  int num_devices() const { return ssfm_multi_num_devices(h_); }

 private:
  ssfm_multi_handle h_ = nullptr;
};

namespace detail {
template <class Matrix3>
inline void to_rowmajor(const Matrix3& E, double* out) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[3 * r + c] = E(r, c);
}
template <class Matrix3>
inline void from_rowmajor(const double* in, Matrix3* E) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) (*E)(r, c) = in[3 * r + c];
}
template <class Options>
inline auto lo_fields(const Options& o, SsfmOptions* s, int) -> decltype(o.num_lo_steps_, void()) {
  s->num_lo_steps = o.num_lo_steps_;
  s->threshold_multiplier = o.threshold_multiplier_;
  s->num_lsq_iterations = o.num_lsq_iterations_;
  s->min_sample_multiplicator = o.min_sample_multiplicator_;
  s->non_min_sample_multiplier = o.non_min_sample_multiplier_;
  s->lo_starting_iterations = o.lo_starting_iterations_;
  s->final_least_squares = o.final_least_squares_ ? 1 : 0;
}
template <class Options>
inline void lo_fields(const Options&, SsfmOptions*, long) {}
template <class Options>
inline SsfmOptions to_c_options(const Options& o) {
  SsfmOptions s;
  ssfm_default_options(&s);
  s.min_num_iterations = o.min_num_iterations_;
  s.max_num_iterations = o.max_num_iterations_;
  s.success_probability = o.success_probability_;
  s.squared_inlier_threshold = o.squared_inlier_threshold_;
  s.random_seed = o.random_seed_;
  lo_fields(o, &s, 0);
  return s;
}
}  // namespace detail

// The RansacLib estimator concept (include/sphericalsfm/estimator.h:6-29) on the GPU engine.  Each
// member forwards to the engine with a batch of one, so any RansacLib-style driver still works;
// the intended fast path is the batched driver below.
template <class Matrix3 = Mat3d>
class GpuSphericalEstimator {
 public:
  // rays: n x 6 doubles = RayPairList memory; held by reference like the original (spherical_estimator.h:11,15).
  GpuSphericalEstimator(const Engine& eng, const double* rays, int n, bool use_poly_solver, bool inward,
                        uint32_t pair_id = 0)
      : h_(eng.get()), rays_(rays), n_(n), poly_(use_poly_solver), inward_(inward), pair_id_(pair_id) {}
  template <class RayPairList>
  GpuSphericalEstimator(const Engine& eng, const RayPairList& correspondences, bool use_poly_solver, bool inward,
                        uint32_t pair_id = 0)
      : GpuSphericalEstimator(eng, correspondences.empty() ? nullptr : reinterpret_cast<const double*>(&correspondences[0]),
                              (int)correspondences.size(), use_poly_solver, inward, pair_id) {
    static_assert(sizeof(correspondences[0]) == 48, "RayPair must be two packed 3-vectors of double");
  }
  inline int min_sample_size() const { return 3; }
  inline int non_minimal_sample_size() const { return 4; }
  inline int num_data() const { return n_; }
  const double* rays() const { return rays_; }
  bool inward() const { return inward_; }
  bool use_poly_solver() const { return poly_; }
  uint32_t pair_id() const { return pair_id_; }
  ssfm_handle handle() const { return h_; }

  int MinimalSolver(const std::vector<int>& sample, std::vector<Matrix3>* Es) const {  // spherical_estimator.cpp:80-84
    Es->clear();
    if (sample.size() != 3) return 0;
    double models[24];
    int nm = 0;
    check(ssfm_minimal_solve(h_, rays_, n_, sample.data(), 1, poly_ ? SSFM_SOLVER_POLYNOMIAL : SSFM_SOLVER_ACTION_MATRIX,
                             models, &nm));
    for (int k = 0; k < nm; ++k) {
      const double* p = models + 6 * k;
      const double E[9] = {p[0], p[1], p[2], p[1], -p[0], p[3], p[4], p[5], 0.0};
      Matrix3 M;
      detail::from_rowmajor(E, &M);
      Es->push_back(M);
    }
    return nm;
  }
  int NonMinimalSolver(const std::vector<int>& sample, Matrix3* E) const {  // spherical_estimator.cpp:86-108
    const int offs[2] = {0, (int)sample.size()};
    double out[9];
    int ok = 0;
    check(ssfm_non_minimal_solve(h_, rays_, n_, sample.data(), offs, 1, out, &ok));
    if (ok) detail::from_rowmajor(out, E);
    return ok;
  }
  double EvaluateModelOnPoint(const Matrix3& E, int i) const {  // spherical_estimator.cpp:67-78
    double e9[9], score = 0.0;
    int cnt = 0;
    detail::to_rowmajor(E, e9);
    check(ssfm_score_exact(h_, e9, 1, rays_ + 6 * (size_t)i, 1, std::numeric_limits<double>::max(), &score, &cnt));
    return score;  // min(err, DBL_MAX) summed over one point
  }
  void LeastSquares(const std::vector<int>& sample, Matrix3* E) const {  // spherical_estimator.cpp:110-157
    const int offs[2] = {0, (int)sample.size()};
    double e9[9];
    detail::to_rowmajor(*E, e9);
    check(ssfm_least_squares(h_, rays_, n_, sample.data(), offs, 1, inward_ ? 1 : 0, e9));
    detail::from_rowmajor(e9, E);
  }
  template <class Vector3>
  void Decompose(const Matrix3& E, const std::vector<int>& /*inliers*/, Matrix3* R, Vector3* t) const {  // :159-164
    double e9[9], r[3], tt[3];
    detail::to_rowmajor(E, e9);
    check(ssfm_decompose(h_, e9, 1, inward_ ? 1 : 0, r, tt));
    const double th = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th >= 1e-10) {  // so3exp, src/so3.cpp:16-23
      const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
      const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double kk = 0;
          for (int l = 0; l < 3; ++l) kk += K[3 * i + l] * K[3 * l + j];
          Rm[3 * i + j] += std::sin(th) * K[3 * i + j] + (1 - std::cos(th)) * kk;
        }
    }
    detail::from_rowmajor(Rm, R);
    for (int i = 0; i < 3; ++i) (*t)[i] = tt[i];
  }

 private:
  ssfm_handle h_;
  const double* rays_;
  int n_;
  bool poly_, inward_;
  uint32_t pair_id_;
};

namespace detail {
template <class Solver>
inline auto solver_fields(const Solver& s, SsfmOptions* o, int) -> decltype(s.use_poly_solver(), void()) {
  o->solver = s.use_poly_solver() ? SSFM_SOLVER_POLYNOMIAL : SSFM_SOLVER_ACTION_MATRIX;
}
template <class Solver>
inline void solver_fields(const Solver& s, SsfmOptions* o, long) {  // GpuSixPointEstimator
  o->solver = Solver::kSolver;
  o->sixpt_focal_scoring = s.focal_scoring() ? 1 : 0;
}
template <class Matrix3>
inline auto store_model(const SsfmPairResult& r, Matrix3* m, int) -> decltype((*m)(0, 0), void()) {
  from_rowmajor(r.E, m);
}
template <class Solution>
inline void store_model(const SsfmPairResult& r, Solution* m, long) {  // SixPointSolution
  for (int d = 0; d < 3; ++d) { m->t[d] = r.t[d]; m->r[d] = r.r[d]; }
  m->focal = r.focal;
}
template <class Options, class Solver, class Model, class Statistics>
int estimate_one(int driver, const Options& options, const Solver& solver, Model* best_model, Statistics* statistics) {
  SsfmOptions o = to_c_options(options);
  o.driver = driver;
  solver_fields(solver, &o, 0);
  o.inward = solver.inward() ? 1 : 0;
  o.first_pair_id = solver.pair_id();
  const int n = solver.num_data();
  const int64_t offsets[2] = {0, n};
  SsfmBatch b;
  b.num_pairs = 1;
  b.offsets = offsets;
  b.rays = solver.rays();
  b.rays_on_device = 0;
  b.ray_format = SSFM_RAYS_F64;
  SsfmPairResult r;
  std::vector<uint8_t> flags(n > 0 ? n : 1);
  check(ssfm_estimate_pairs(solver.handle(), &b, &o, &r, flags.data()));
  store_model(r, best_model, 0);
  statistics->num_iterations = r.num_iterations;
  statistics->best_num_inliers = r.best_num_inliers;
  statistics->best_model_score = r.best_model_score;
  statistics->inlier_ratio = r.inlier_ratio;
  statistics->number_lo_iterations = r.number_lo_iterations;
  statistics->inlier_indices.clear();
  for (int i = 0; i < n; ++i)
    if (flags[i]) statistics->inlier_indices.push_back(i);
  return r.best_num_inliers;
}
}  // namespace detail

// Same template parameters and EstimateModel signature as ransac_lib::LocallyOptimizedMSAC
// (include/RansacLib/ransac.h:119-129), so examples/spherical_sfm_tools.cpp:380-387 compiles after
// swapping the namespace.  Sampling is the engine's Philox stream (the Sampler parameter is ignored).
template <class Model, class ModelVector, class Solver, class Sampler = void>
class LocallyOptimizedMSAC {
 public:
  template <class Options, class Statistics>
  int EstimateModel(const Options& options, const Solver& solver, Model* best_model, Statistics* statistics) const {
    return detail::estimate_one(SSFM_DRIVER_LO_MSAC, options, solver, best_model, statistics);
  }
};
template <class Model, class ModelVector, class Solver, class Sampler = void>
class VanillaMSAC {  // evaluation/vanilla_ransac.h:17-23
 public:
  template <class Options, class Statistics>
  int EstimateModel(const Options& options, const Solver& solver, Model* best_model, Statistics* statistics) const {
    return detail::estimate_one(SSFM_DRIVER_VANILLA_MSAC, options, solver, best_model, statistics);
  }
};

// ---- the six-point shared-focal estimator (examples/six_point_estimator.h:9-37) ------------------------
struct SixPointSolution {  // :9-13 (sphericalsfm::Pose reduced to its t and r members)
  double t[3] = {0, 0, 0};
  double r[3] = {0, 0, 0};
  double focal = 0.0;
};

class GpuSixPointEstimator {
 public:
  static constexpr int kSolver = SSFM_SOLVER_SIXPT_FOCAL;
  // focal_scoring = false reproduces EvaluateModelOnPoint as written upstream (E on the raw rays, :78-91)
  GpuSixPointEstimator(const Engine& eng, const double* rays, int n, bool focal_scoring = false, uint32_t pair_id = 0)
      : h_(eng.get()), rays_(rays), n_(n), focal_scoring_(focal_scoring), pair_id_(pair_id) {}
  template <class RayPairList>
  GpuSixPointEstimator(const Engine& eng, const RayPairList& correspondences, bool focal_scoring = false, uint32_t pair_id = 0)
      : GpuSixPointEstimator(eng, correspondences.empty() ? nullptr : reinterpret_cast<const double*>(&correspondences[0]),
                             (int)correspondences.size(), focal_scoring, pair_id) {
    static_assert(sizeof(correspondences[0]) == 48, "RayPair must be two packed 3-vectors of double");
  }
  inline int min_sample_size() const { return 6; }
  inline int non_minimal_sample_size() const { return 7; }
  inline int num_data() const { return n_; }
  const double* rays() const { return rays_; }
  bool inward() const { return false; }
  bool focal_scoring() const { return focal_scoring_; }
  uint32_t pair_id() const { return pair_id_; }
  ssfm_handle handle() const { return h_; }

  int MinimalSolver(const std::vector<int>& sample, std::vector<SixPointSolution>* solns) const {  // .cpp:93-119
    solns->clear();
    if (sample.size() < 6) return 0;
    double models[15 * 7];
    int nm = 0;
    check(ssfm_sixpt_solve(h_, rays_, n_, sample.data(), 1, models, &nm));
    for (int k = 0; k < nm; ++k) {
      SixPointSolution s;
      for (int d = 0; d < 3; ++d) { s.t[d] = models[7 * k + d]; s.r[d] = models[7 * k + 3 + d]; }
      s.focal = models[7 * k + 6];
      solns->push_back(s);
    }
    return nm;
  }
  int NonMinimalSolver(const std::vector<int>& sample, SixPointSolution* soln) const {  // .cpp:121-144
    std::vector<SixPointSolution> solns;
    const int nsols = MinimalSolver(sample, &solns);
    if (nsols == 0) return 0;
    double best_score = INFINITY;
    int best_ind = 0;
    for (int i = 0; i < nsols; ++i) {
      double score = 0;
      for (size_t j = 0; j < sample.size(); ++j) score += EvaluateModelOnPoint(solns[i], sample[j]);
      if (score < best_score) { best_score = score; best_ind = i; }
    }
    *soln = solns[best_ind];
    return 1;
  }
  // The 3x3 matrix the estimator scores with: skew3(t) * so3exp(r)  (.cpp:85), optionally Kinv . Kinv.
  void ScoringMatrix(const SixPointSolution& s, double* G) const {
    const double th = std::sqrt(s.r[0] * s.r[0] + s.r[1] * s.r[1] + s.r[2] * s.r[2]);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th >= 1e-10) {
      const double k[3] = {s.r[0] / th, s.r[1] / th, s.r[2] / th};
      const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double kk = 0;
          for (int l = 0; l < 3; ++l) kk += K[3 * i + l] * K[3 * l + j];
          R[3 * i + j] += std::sin(th) * K[3 * i + j] + (1 - std::cos(th)) * kk;
        }
    }
    const double S[9] = {0, -s.t[2], s.t[1], s.t[2], 0, -s.t[0], -s.t[1], s.t[0], 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) G[3 * i + j] = S[3 * i] * R[j] + S[3 * i + 1] * R[3 + j] + S[3 * i + 2] * R[6 + j];
    if (focal_scoring_) { G[2] *= s.focal; G[5] *= s.focal; G[6] *= s.focal; G[7] *= s.focal; G[8] *= s.focal * s.focal; }
  }
  double EvaluateModelOnPoint(const SixPointSolution& soln, int i) const {  // .cpp:78-91
    double G[9], score = 0.0;
    int cnt = 0;
    ScoringMatrix(soln, G);
    check(ssfm_score_exact(h_, G, 1, rays_ + 6 * (size_t)i, 1, std::numeric_limits<double>::max(), &score, &cnt));
    return score;
  }
  void LeastSquares(const std::vector<int>& sample, SixPointSolution* soln) const {  // .cpp:146-192
    const int offs[2] = {0, (int)sample.size()};
    double m[7] = {soln->t[0], soln->t[1], soln->t[2], soln->r[0], soln->r[1], soln->r[2], soln->focal};
    check(ssfm_sixpt_least_squares(h_, rays_, n_, sample.data(), offs, 1, m));
    for (int d = 0; d < 3; ++d) { soln->t[d] = m[d]; soln->r[d] = m[3 + d]; }
    soln->focal = m[6];
  }

 private:
  ssfm_handle h_;
  const double* rays_;
  int n_;
  bool focal_scoring_;
  uint32_t pair_id_;
};

// ---- the legacy drivers (include/sphericalsfm/msac.h, preemptive_ransac.h) ----------------------------
// Upstream they take a vector of pre-allocated estimators (one per hypothesis) and return a pointer to the
// winning one.  Here the hypotheses live on the device: the vector's size is the hypothesis budget, and the
// winner's model is written into estimators[0], which is what *best_estimator points to on return.
template <class Matrix3 = Mat3d>
struct GpuSphericalFastEstimator {  // include/sphericalsfm/spherical_fast_estimator.h:8-24
  Matrix3 E;
  double r[3] = {0, 0, 0}, t[3] = {0, 0, 0};  // decomposeE(inward, r, t) of the winner (:290-341), filled by compute()
  static constexpr int kSolver = SSFM_SOLVER_FAST_STURM;
  int sampleSize() { return 3; }
  bool canRefine() { return true; }
  template <class Vector3>
  void decomposeE(bool /*inward*/, Vector3& r_out, Vector3& t_out) const {
    for (int i = 0; i < 3; ++i) { r_out[i] = r[i]; t_out[i] = t[i]; }
  }
};

namespace detail {
template <class It, class EstimatorType>
int legacy_compute(const Engine& eng, SsfmOptions o, It begin, It end, std::vector<EstimatorType*>& estimators,
                   EstimatorType** best_estimator, std::vector<bool>& inliers, int* iter) {
  static_assert(sizeof(*begin) == 48, "RayPair must be two packed 3-vectors of double");
  const int n = (int)(end - begin);
  if (estimators.empty()) throw Error(SSFM_ERR_INVALID, "no estimators (the vector's size is the hypothesis budget)");
  o.solver = EstimatorType::kSolver;
  o.fixed_budget = (int32_t)estimators.size();
  const int64_t offsets[2] = {0, n};
  SsfmBatch b;
  b.num_pairs = 1;
  b.offsets = offsets;
  b.rays = n > 0 ? reinterpret_cast<const double*>(&*begin) : nullptr;
  b.rays_on_device = 0;
  b.ray_format = SSFM_RAYS_F64;
  SsfmPairResult res;
  std::vector<uint8_t> flags(n > 0 ? n : 1);
  check(ssfm_estimate_pairs(eng.get(), &b, &o, &res, flags.data()));
  inliers.assign(n, false);
  for (int i = 0; i < n; ++i) inliers[i] = flags[i] != 0;
  if (iter) *iter = (int)res.num_iterations;
  if (res.status == SSFM_PAIR_OK) {
    from_rowmajor(res.E, &estimators[0]->E);
    for (int i = 0; i < 3; ++i) { estimators[0]->r[i] = res.r[i]; estimators[0]->t[i] = res.t[i]; }
    *best_estimator = estimators[0];
  }
  return res.best_num_inliers;
}
}  // namespace detail

template <class ListType, class EstimatorType>
struct MSAC {  // include/sphericalsfm/msac.h:29-131
  const Engine& engine;
  double inlier_threshold = 0.001;
  double prob_success;
  double init_outlier_ratio;  // kept for source compatibility; the loop overwrites it before use (msac.h:113)
  int iter = 0;               // number of iterations performed
  uint32_t random_seed = 0, pair_id = 0;
  bool inward = false;
  explicit MSAC(const Engine& eng, double _prob_success = 0.999, double _init_outlier_ratio = 0.8)
      : engine(eng), prob_success(_prob_success), init_outlier_ratio(_init_outlier_ratio) {}
  int compute(typename ListType::iterator begin, typename ListType::iterator end, std::vector<EstimatorType*>& estimators,
              EstimatorType** best_estimator, std::vector<bool>& inliers) {
    SsfmOptions o;
    ssfm_default_options(&o);
    o.driver = SSFM_DRIVER_MSAC_FIXED;
    o.squared_inlier_threshold = inlier_threshold * inlier_threshold;
    o.fixed_prob_success = prob_success;
    o.random_seed = random_seed;
    o.first_pair_id = pair_id;
    o.inward = inward ? 1 : 0;
    return detail::legacy_compute(engine, o, begin, end, estimators, best_estimator, inliers, &iter);
  }
};

template <class ListType, class EstimatorType>
struct PreemptiveRANSAC {  // include/sphericalsfm/preemptive_ransac.h:30-139
  const Engine& engine;
  double inlier_threshold = 0.001;
  size_t B;  // block size
  uint32_t random_seed = 0, pair_id = 0;
  bool inward = false;
  explicit PreemptiveRANSAC(const Engine& eng, size_t _B = 10) : engine(eng), B(_B) {}
  int compute(typename ListType::iterator begin, typename ListType::iterator end, std::vector<EstimatorType*>& estimators,
              EstimatorType** best_estimator, std::vector<bool>& inliers) {
    SsfmOptions o;
    ssfm_default_options(&o);
    o.driver = SSFM_DRIVER_PREEMPTIVE;
    o.squared_inlier_threshold = inlier_threshold * inlier_threshold;
    o.preemptive_block = (int32_t)B;
    o.random_seed = random_seed;
    o.first_pair_id = pair_id;
    o.inward = inward ? 1 : 0;
    return detail::legacy_compute(engine, o, begin, end, estimators, best_estimator, inliers, (int*)nullptr);
  }
};

// The new batched-pairs entry point: one call for all pairs (replaces the OpenMP loop).
// pair_lists: one RayPairList per image pair.  results: one SsfmPairResult per pair;
// inlier_flags (optional): concatenated 0/1 flags in pair order.
namespace detail {
template <class Options, class RayPairList, class Call>
inline void estimate_pairs_impl(const Options& options, const std::vector<RayPairList>& pair_lists, bool use_poly_solver, bool inward,
                                std::vector<SsfmPairResult>* results, std::vector<uint8_t>* inlier_flags, Call call);
}
template <class Options, class RayPairList>
inline void EstimatePairs(const Engine& eng, const Options& options, const std::vector<RayPairList>& pair_lists,
                          bool use_poly_solver, bool inward, std::vector<SsfmPairResult>* results,
                          std::vector<uint8_t>* inlier_flags = nullptr) {
  detail::estimate_pairs_impl(options, pair_lists, use_poly_solver, inward, results, inlier_flags,
                              [&](const SsfmBatch* b, const SsfmOptions* o, SsfmPairResult* r, uint8_t* f) {
                                return ssfm_estimate_pairs(eng.get(), b, o, r, f);
                              });
}
// The same over all GPUs of the handle: shards of the pair list, one per device; identical table.
template <class Options, class RayPairList>
inline void EstimatePairs(const MultiEngine& eng, const Options& options, const std::vector<RayPairList>& pair_lists,
                          bool use_poly_solver, bool inward, std::vector<SsfmPairResult>* results,
                          std::vector<uint8_t>* inlier_flags = nullptr) {
  detail::estimate_pairs_impl(options, pair_lists, use_poly_solver, inward, results, inlier_flags,
                              [&](const SsfmBatch* b, const SsfmOptions* o, SsfmPairResult* r, uint8_t* f) {
                                return ssfm_estimate_pairs_multi(eng.get(), b, o, r, f);
                              });
}
template <class Options, class RayPairList, class Call>
inline void detail::estimate_pairs_impl(const Options& options, const std::vector<RayPairList>& pair_lists, bool use_poly_solver,
                                        bool inward, std::vector<SsfmPairResult>* results, std::vector<uint8_t>* inlier_flags,
                                        Call call) {
  SsfmOptions o = detail::to_c_options(options);
  o.solver = use_poly_solver ? SSFM_SOLVER_POLYNOMIAL : SSFM_SOLVER_ACTION_MATRIX;
  o.inward = inward ? 1 : 0;
  std::vector<int64_t> offsets(pair_lists.size() + 1, 0);
  for (size_t p = 0; p < pair_lists.size(); ++p) offsets[p + 1] = offsets[p] + (int64_t)pair_lists[p].size();
  std::vector<double> rays((size_t)offsets.back() * 6);
  for (size_t p = 0; p < pair_lists.size(); ++p)
    if (!pair_lists[p].empty())
      std::memcpy(rays.data() + 6 * (size_t)offsets[p], &pair_lists[p][0], 48 * pair_lists[p].size());
  SsfmBatch b;
  b.num_pairs = (int32_t)pair_lists.size();
  b.offsets = offsets.data();
  b.rays = rays.data();
  b.rays_on_device = 0;
  b.ray_format = SSFM_RAYS_F64;
  results->resize(pair_lists.size());
  if (inlier_flags) inlier_flags->assign((size_t)offsets.back(), 0);
  check(call(&b, &o, results->data(), inlier_flags ? inlier_flags->data() : nullptr));
}

// SfM::Retriangulate (src/sfm.cpp:156-192) for all points in one call.  camera_tr[i] = {t, r} of GetPose(i);
// tracks[j] = the observations of point j as (camera index, x, y) with the principal point removed; `focal` =
// intrinsics.focal.  points[j] is zero unless status[j] == SSFM_PAIR_OK, exactly like SetPoint(j, Zero) followed by
// the conditional SetPoint(j, X).
struct TrackObservation {
  int camera;
  double x, y;
};
struct Point3d {
  double x = 0, y = 0, z = 0;
};
inline void Retriangulate(const Engine& eng, const std::vector<double>& camera_tr /* 6 per camera */,
                          const std::vector<std::vector<TrackObservation>>& tracks, double focal, std::vector<Point3d>* points,
                          std::vector<int32_t>* num_inliers = nullptr, std::vector<int32_t>* status = nullptr) {
  std::vector<int64_t> offs(tracks.size() + 1, 0);
  for (size_t j = 0; j < tracks.size(); ++j) offs[j + 1] = offs[j] + (int64_t)tracks[j].size();
  std::vector<int32_t> cam((size_t)offs.back());
  std::vector<double> xy((size_t)offs.back() * 2);
  for (size_t j = 0; j < tracks.size(); ++j)
    for (size_t k = 0; k < tracks[j].size(); ++k) {
      const size_t o = (size_t)offs[j] + k;
      cam[o] = tracks[j][k].camera;
      xy[2 * o] = tracks[j][k].x;
      xy[2 * o + 1] = tracks[j][k].y;
    }
  SsfmTrackBatch tb;
  tb.num_cameras = (int32_t)(camera_tr.size() / 6);
  tb.camera_tr = camera_tr.data();
  tb.num_points = (int32_t)tracks.size();
  tb.obs_offsets = offs.data();
  tb.obs_camera = cam.data();
  tb.obs_xy = xy.data();
  tb.focal = focal;
  SsfmOptions o;
  ssfm_default_options(&o);
  o.squared_inlier_threshold = 4.0;  // src/sfm.cpp:177
  o.final_least_squares = 1;         // :178
  std::vector<double> pts(tracks.size() * 3 + 3);
  std::vector<int32_t> ninl(tracks.size() + 1), st(tracks.size() + 1);
  check(ssfm_retriangulate(eng.get(), &tb, &o, pts.data(), ninl.data(), st.data(), nullptr));
  points->resize(tracks.size());
  for (size_t j = 0; j < tracks.size(); ++j) { (*points)[j].x = pts[3 * j]; (*points)[j].y = pts[3 * j + 1]; (*points)[j].z = pts[3 * j + 2]; }
  if (num_inliers) num_inliers->assign(ninl.begin(), ninl.begin() + tracks.size());
  if (status) status->assign(st.begin(), st.begin() + tracks.size());
}

}  // namespace ssfm_b200
