// oracle_capi.cpp -- extern "C" surface of the CPU ORACLE (test infrastructure; see ssfm_oracle.hpp).
//
// Built two ways by oracle/Makefile:
//   liboracle.so            : driver loops = the restatement in lomsac.hpp (source-only, travels)
//   _ref/libssfm_ref.so     : -DSSFM_USE_REFERENCE_RANSACLIB, driver loops = the reference's OWN
//                             headers, compiled where they lie:
//                               /root/reference/include/RansacLib/{ransac,sampling,utils}.h
//                               /root/reference/evaluation/vanilla_ransac.h
//                             (header-only, <random> only -> builds without Eigen/Ceres).
//                             The estimator plugged into them is the restated SphericalEstimator
//                             (the reference's own src/*.cpp need Eigen + Ceres: not buildable here).
#include "oracle_capi.h"

#include <chrono>
#include <cstring>

#include "lomsac.hpp"
#include "ssfm_oracle.hpp"
#include "tri_oracle.hpp"

#include <atomic>
#include <thread>

#ifdef SSFM_USE_REFERENCE_RANSACLIB
#include <RansacLib/ransac.h>
#include <vanilla_ransac.h>
// The reference's pre-emptive driver, compiled where it lies.  Its sampler calls the C library's
// rand() (preemptive_ransac.h:18); the call expression is redirected -- by a macro, the header is
// not touched -- to the Philox-backed stream the restatement uses (ssfm_oracle::philox_rand31),
// so that both draw the same samples and can be compared bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <vector>
namespace pinned_rand {
thread_local uint32_t seed = 0, pair = 0, hyp = 0, draw = 0;
inline int next() { return ssfm_oracle::philox_rand31(seed, pair, hyp, draw++); }
}  // namespace pinned_rand
#define rand() pinned_rand::next()
#include <sphericalsfm/preemptive_ransac.h>
#undef rand
#endif

using namespace ssfm_oracle;

namespace {

// Dynamic-schedule parallel for over [0, n) on `nthreads` host threads (the role of
// `#pragma omp parallel for` at examples/spherical_sfm_tools.cpp:332; std::thread so the oracle
// builds without libgomp).
template <class F>
void parallel_for(int n, int nthreads, F f) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  if (nthreads == 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&]() {
      for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i);
    });
  for (auto& th : pool) th.join();
}

Options to_options(const OrcOptions& o) {
  Options p;
  p.min_num_iterations = o.min_num_iterations;
  p.max_num_iterations = o.max_num_iterations;
  p.success_probability = o.success_probability;
  p.squared_inlier_threshold = o.squared_inlier_threshold;
  p.random_seed = o.random_seed;
  p.num_lo_steps = o.num_lo_steps;
  p.threshold_multiplier = o.threshold_multiplier;
  p.num_lsq_iterations = o.num_lsq_iterations;
  p.min_sample_multiplicator = o.min_sample_multiplicator;
  p.non_min_sample_multiplier = o.non_min_sample_multiplier;
  p.lo_starting_iterations = o.lo_starting_iterations;
  p.final_least_squares = o.final_least_squares != 0;
  return p;
}

#ifdef SSFM_USE_REFERENCE_RANSACLIB
// EstimatorType for the reference's PreemptiveRANSAC template (the interface of
// include/sphericalsfm/estimator.h:6-23 as that driver uses it): the restated minimal solver
// behind compute(), the restated Sampson error behind score().
struct RefPreemptAdapter {
  SolverKind kind = FAST_STURM;
  Mat3 Esolns[4];
  Mat3 E;
  bool has = false;
  long long* evals = nullptr;
  int sampleSize() { return 3; }
  int compute(std::vector<RayPair>::iterator b, std::vector<RayPair>::iterator e) {
    pinned_rand::hyp++;  // compute() directly follows each random_sample(): the next draws belong to the next hypothesis
    pinned_rand::draw = 0;
    const int idx[3] = {0, 1, 2};
    double models[4][6];
    const int nm = solve_spherical(&*b, idx, (int)(e - b), kind, models);
    for (int k = 0; k < nm; ++k) Esolns[k] = mat_from_p(models[k]);
    has = nm > 0;
    if (has) E = Esolns[0];
    return nm;
  }
  void chooseSolution(int j) { E = Esolns[j]; }
  double score(std::vector<RayPair>::iterator it) {
    if (!has) return INFINITY;
    ++*evals;
    return sampson_sq(E, *it);
  }
};
#endif

int estimate_one(const RayPair* corr, int n, const OrcOptions& o, uint32_t pair_id, OrcResult* out, int* inlier_idx) {
  SphericalEstimator est(corr, n, (SolverKind)o.solver_kind, o.inward != 0, pair_id, o.complex_mode);
  Mat3 E;
  for (int i = 0; i < 9; ++i) E.m[i] = 0.0;
  Statistics st;
  const Options opt = to_options(o);
  typedef PhiloxSampling<SphericalEstimator> Sampler;
#ifdef SSFM_USE_REFERENCE_RANSACLIB
  if (o.driver == 0 || o.driver == 1) {
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
    if (o.driver == 0) {
      ransac_lib::LocallyOptimizedMSAC<Mat3, std::vector<Mat3>, SphericalEstimator, Sampler> ransac;
      ransac.EstimateModel(ro, est, &E, &rs);
    } else {
      ransac_lib::VanillaMSAC<Mat3, std::vector<Mat3>, SphericalEstimator, Sampler> ransac;
      ransac.EstimateModel(ro, est, &E, &rs);
    }
    st.num_iterations = rs.num_iterations;
    st.best_num_inliers = rs.best_num_inliers;
    st.best_model_score = rs.best_model_score;
    st.inlier_ratio = rs.inlier_ratio;
    st.inlier_indices = rs.inlier_indices;
    st.number_lo_iterations = rs.number_lo_iterations;
  } else if (o.driver == 3) {
    const int M = o.legacy_budget, B = o.preemptive_block;
    if (n >= 4 && M > 0 && B > 0) {
      std::vector<RayPair> list(corr, corr + n);
      std::vector<RefPreemptAdapter> pool(M);
      std::vector<RefPreemptAdapter*> ptrs(M);
      long long evals = 0;
      for (int i = 0; i < M; ++i) {
        pool[i].kind = (SolverKind)o.solver_kind;
        pool[i].evals = &evals;
        ptrs[i] = &pool[i];
      }
      pinned_rand::seed = o.random_seed;
      pinned_rand::pair = pair_id;
      pinned_rand::hyp = 0;
      pinned_rand::draw = 0;
      sphericalsfm::PreemptiveRANSAC<std::vector<RayPair>, RefPreemptAdapter> pr((size_t)B);
      pr.inlier_threshold = std::sqrt(o.squared_inlier_threshold);
      RefPreemptAdapter* best = nullptr;
      std::vector<bool> inl;
      const int ninl = pr.compute(list.begin(), list.end(), ptrs, &best, inl);
      st.num_iterations = (uint32_t)M;
      if (best && best->has) {
        E = best->E;
        double cost = 0.0;
        for (int i = 0; i < n; ++i) {
          const double sc = sampson_sq(E, corr[i]);
          cost += (sc <= o.squared_inlier_threshold) ? sc : o.squared_inlier_threshold;
          if (inl[i]) st.inlier_indices.push_back(i);
        }
        st.best_num_inliers = ninl;
        st.best_model_score = cost;
        st.inlier_ratio = (double)ninl / (double)n;
      }
      est.evals_ += evals;
    }
  } else {
    legacy_msac(opt, o.legacy_budget, o.legacy_prob_success, est,
                [&](uint32_t it, int N, int k, int* idx) { knuth_sample(o.random_seed, pair_id, it, N, k, idx); }, &E, &st);
  }
#else
  if (o.driver == 0)
    lo_msac<SphericalEstimator, Sampler>(opt, est, &E, &st);
  else if (o.driver == 1)
    vanilla_msac<SphericalEstimator, Sampler>(opt, est, &E, &st);
  else if (o.driver == 3)
    preemptive_ransac(opt, o.legacy_budget, o.preemptive_block, est,
                      [&](uint32_t hyp, int N, int k, int* idx) { knuth_sample(o.random_seed, pair_id, hyp, N, k, idx); }, &E, &st);
  else
    legacy_msac(opt, o.legacy_budget, o.legacy_prob_success, est,
                [&](uint32_t it, int N, int k, int* idx) { knuth_sample(o.random_seed, pair_id, it, N, k, idx); }, &E, &st);
#endif
  std::memcpy(out->E, E.m, sizeof(E.m));
  out->num_iterations = st.num_iterations;
  out->best_num_inliers = st.best_num_inliers;
  out->best_model_score = st.best_model_score;
  out->inlier_ratio = st.inlier_ratio;
  out->number_lo_iterations = st.number_lo_iterations;
  out->evals = est.evals_;
  for (int i = 0; i < 3; ++i) out->r[i] = out->t[i] = 0.0;
  if (n < 3 || (o.driver == 3 && n < 4)) {
    out->status = 1;
  } else if (!(st.best_model_score < std::numeric_limits<double>::max())) {
    out->status = 2;
  } else {
    out->status = 0;
    decompose_spherical_essential_matrix(E, o.inward != 0, out->r, out->t);
  }
  if (inlier_idx) {
    for (size_t i = 0; i < st.inlier_indices.size(); ++i) inlier_idx[i] = st.inlier_indices[i];
  }
  return st.best_num_inliers;
}

}  // namespace

extern "C" {

int orc_is_reference(void) {
#ifdef SSFM_USE_REFERENCE_RANSACLIB
  return 1;
#else
  return 0;
#endif
}

void orc_philox_sample(uint32_t seed, uint32_t pair, uint32_t iter, int k, int n, int* idx) {
  philox_sample(seed, pair, iter, k, n, idx);
}
// SfM::Retriangulate for one point (src/sfm.cpp:156-192).  cam_tr: per OBSERVATION t[3], r[3] of its camera pose.
int orc_triangulate(const double* cam_tr, const double* obs_xy, int n, double focal, const OrcOptions* o, uint32_t point_id,
                    OrcResult* out, int* inlier_idx) {
  std::vector<TriObservation> obs(n > 0 ? n : 0);
  for (int i = 0; i < n; ++i) obs[i] = make_tri_observation(cam_tr + 6 * i, cam_tr + 6 * i + 3, obs_xy + 2 * i, focal);
  TriangulationEstimator est(obs.data(), n, point_id);
  Point3 X;
  Statistics st;
  for (int i = 0; i < 9; ++i) out->E[i] = 0.0;
  for (int i = 0; i < 3; ++i) out->r[i] = out->t[i] = 0.0;
  out->status = 3;  // fewer than 3 observations: the point stays at zero (sfm.cpp:172-173)
  out->num_iterations = 0; out->best_num_inliers = 0; out->best_model_score = std::numeric_limits<double>::max();
  out->inlier_ratio = 0; out->number_lo_iterations = 0; out->evals = 0;
  if (n < 3) return 0;
  const Options opt = to_options(*o);
#ifdef SSFM_USE_REFERENCE_RANSACLIB
  {
    ransac_lib::LORansacOptions ro;
    ro.min_num_iterations_ = o->min_num_iterations; ro.max_num_iterations_ = o->max_num_iterations;
    ro.success_probability_ = o->success_probability; ro.squared_inlier_threshold_ = o->squared_inlier_threshold;
    ro.random_seed_ = o->random_seed; ro.num_lo_steps_ = o->num_lo_steps; ro.threshold_multiplier_ = o->threshold_multiplier;
    ro.num_lsq_iterations_ = o->num_lsq_iterations; ro.min_sample_multiplicator_ = o->min_sample_multiplicator;
    ro.non_min_sample_multiplier_ = o->non_min_sample_multiplier; ro.lo_starting_iterations_ = o->lo_starting_iterations;
    ro.final_least_squares_ = o->final_least_squares != 0;
    ransac_lib::RansacStatistics rs;
    ransac_lib::LocallyOptimizedMSAC<Point3, std::vector<Point3>, TriangulationEstimator, PhiloxSampling<TriangulationEstimator> > ransac;
    ransac.EstimateModel(ro, est, &X, &rs);
    st.num_iterations = rs.num_iterations; st.best_num_inliers = rs.best_num_inliers; st.best_model_score = rs.best_model_score;
    st.inlier_ratio = rs.inlier_ratio; st.inlier_indices = rs.inlier_indices; st.number_lo_iterations = rs.number_lo_iterations;
  }
#else
  lo_msac<TriangulationEstimator, PhiloxSampling<TriangulationEstimator> >(opt, est, &X, &st);
#endif
  out->num_iterations = st.num_iterations;
  out->best_num_inliers = st.best_num_inliers;
  out->best_model_score = st.best_model_score;
  out->inlier_ratio = st.inlier_ratio;
  out->number_lo_iterations = st.number_lo_iterations;
  out->evals = est.evals_;
  if (st.best_num_inliers < 3) {  // sfm.cpp:186
    out->status = 2;
  } else {
    out->status = 0;
    for (int i = 0; i < 3; ++i) out->E[i] = X.v[i];
  }
  if (inlier_idx)
    for (size_t i = 0; i < st.inlier_indices.size(); ++i) inlier_idx[i] = st.inlier_indices[i];
  return st.best_num_inliers;
}

void orc_knuth_sample(uint32_t seed, uint32_t pair, uint32_t hyp, int N, int n, int* idx) {
  knuth_sample(seed, pair, hyp, N, n, idx);
}

int orc_solve_mode(const double* rays, const int* sample, int n, int kind, int complex_mode, double* models) {
  double m[4][6];
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 6; ++i) m[k][i] = std::numeric_limits<double>::quiet_NaN();
  const int nm = solve_spherical(reinterpret_cast<const RayPair*>(rays), sample, n, (SolverKind)kind, m, complex_mode);
  std::memcpy(models, m, sizeof(m));
  return nm;
}
int orc_solve(const double* rays, const int* sample, int n, int kind, double* models) {
  return orc_solve_mode(rays, sample, n, kind, COMPLEX_CANONICAL, models);
}

// Eigen::EigenSolver<Matrix4d> restated (ssfm_oracle.hpp): ev = 4 x (re, im); V = 4 x 4 x (re, im), row-major.
int orc_eigen34(const double* M16, double* ev, double* V) {
  double M[4][4];
  std::memcpy(M, M16, sizeof(M));
  std::complex<double> e[4], v[4][4];
  const bool ok = eigen34_eigensolver_4x4(M, e, v);
  for (int k = 0; k < 4; ++k) {
    ev[2 * k] = e[k].real();
    ev[2 * k + 1] = e[k].imag();
    for (int i = 0; i < 4; ++i) {
      V[2 * (4 * i + k)] = v[i][k].real();
      V[2 * (4 * i + k) + 1] = v[i][k].imag();
    }
  }
  return ok ? 1 : 0;
}

void orc_sampson(const double* E9, const double* rays, int n, double* out) {
  Mat3 E;
  std::memcpy(E.m, E9, sizeof(E.m));
  const RayPair* c = reinterpret_cast<const RayPair*>(rays);
  for (int i = 0; i < n; ++i) out[i] = sampson_sq(E, c[i]);
}

void orc_score(const double* E9, const double* rays, int n, double thr, double* score, int* ninl) {
  Mat3 E;
  std::memcpy(E.m, E9, sizeof(E.m));
  const RayPair* c = reinterpret_cast<const RayPair*>(rays);
  double s = 0.0;
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    const double e = sampson_sq(E, c[i]);
    s += std::min(e, thr);
    cnt += (e < thr);
  }
  *score = s;
  *ninl = cnt;
}

void orc_decompose(const double* E9, int inward, double* r, double* t) {
  Mat3 E;
  std::memcpy(E.m, E9, sizeof(E.m));
  decompose_spherical_essential_matrix(E, inward != 0, r, t);
}

void orc_make_E(const double* r, int inward, double* E9) {
  const Mat3 E = make_spherical_essential_matrix(so3exp(r), inward != 0);
  std::memcpy(E9, E.m, sizeof(E.m));
}

void orc_lm_refit(const double* rays, const int* sample, int n, int inward, double* E9, int* iters, int* term,
                  double* costs) {
  Mat3 E;
  std::memcpy(E.m, E9, sizeof(E.m));
  double r[3], t[3];
  decompose_spherical_essential_matrix(E, inward != 0, r, t);
  double x[6] = {r[0], r[1], r[2], 0, 0, inward ? 1.0 : -1.0};
  const LMSummary s = lm_refit(reinterpret_cast<const RayPair*>(rays), sample, n, inward != 0, x);
  E = make_spherical_essential_matrix(so3exp(x), inward != 0);
  std::memcpy(E9, E.m, sizeof(E.m));
  if (iters) *iters = s.iterations;
  if (term) *term = s.termination;
  if (costs) { costs[0] = s.initial_cost; costs[1] = s.final_cost; }
}

// The LO generator's draw sequence: ncalls consecutive shuffle_and_resize() calls on iota
// vectors of the given sizes, one std::mt19937 seeded once (ransac.h:143-144, utils.h:34-52).
void orc_lo_shuffle(uint32_t seed, int ncalls, const int* sizes, const int* targets, int* out) {
  std::mt19937 rng;
  rng.seed(seed);
  int o = 0;
  for (int c = 0; c < ncalls; ++c) {
    std::vector<int> v(sizes[c]);
    for (int i = 0; i < sizes[c]; ++i) v[i] = i;
    shuffle_and_resize(targets[c], &rng, &v);
    for (int i = 0; i < targets[c]; ++i) out[o++] = v[i];
  }
}

int orc_estimate_pair(const double* rays, int n, const OrcOptions* opt, uint32_t pair_id, OrcResult* out,
                      int* inlier_idx) {
  return estimate_one(reinterpret_cast<const RayPair*>(rays), n, *opt, pair_id, out, inlier_idx);
}

double orc_estimate_batch(const double* rays, const int64_t* offsets, int npairs, const OrcOptions* opt,
                          uint32_t first_pair_id, int nthreads, OrcResult* out) {
  const RayPair* c = reinterpret_cast<const RayPair*>(rays);
  const auto t0 = std::chrono::steady_clock::now();
  parallel_for(npairs, nthreads, [&](int p) {
    estimate_one(c + offsets[p], (int)(offsets[p + 1] - offsets[p]), *opt, first_pair_id + (uint32_t)p, &out[p], nullptr);
  });
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Same, also returning the final inlier set of every pair as one byte per correspondence (stats.inlier_indices).
double orc_estimate_batch_flags(const double* rays, const int64_t* offsets, int npairs, const OrcOptions* opt,
                                uint32_t first_pair_id, int nthreads, OrcResult* out, uint8_t* flags) {
  const RayPair* c = reinterpret_cast<const RayPair*>(rays);
  const auto t0 = std::chrono::steady_clock::now();
  parallel_for(npairs, nthreads, [&](int p) {
    const int n = (int)(offsets[p + 1] - offsets[p]);
    std::vector<int> idx(n > 0 ? n : 1);
    const int ninl = estimate_one(c + offsets[p], n, *opt, first_pair_id + (uint32_t)p, &out[p], idx.data());
    uint8_t* f = flags + offsets[p];
    for (int i = 0; i < n; ++i) f[i] = 0;
    if (out[p].status == 0)
      for (int i = 0; i < ninl; ++i) f[idx[i]] = 1;
  });
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

double orc_score_batch(const double* models6, int nmodels, const double* rays, int n, double thr, int nthreads,
                       double* scores, int* ninl) {
  const RayPair* c = reinterpret_cast<const RayPair*>(rays);
  const auto t0 = std::chrono::steady_clock::now();
  parallel_for(nmodels, nthreads, [&](int m) {
    const Mat3 E = mat_from_p(models6 + 6 * (size_t)m);
    double s = 0.0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
      const double e = sampson_sq(E, c[i]);
      s += std::min(e, thr);
      cnt += (e < thr);
    }
    scores[m] = s;
    ninl[m] = cnt;
  });
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
