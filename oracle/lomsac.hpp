// lomsac.hpp -- ORACLE restatement of the reference's robust-estimation drivers.
// TEST INFRASTRUCTURE ONLY (see ssfm_oracle.hpp header).
//
//   lo_msac()      <- ransac_lib::LocallyOptimizedMSAC::EstimateModel  include/RansacLib/ransac.h:128-275
//                     (+ LocalOptimization :341-407, LeastSquaresFit :409-420, ScoreModel :295-303,
//                        GetInliers :311-336, NumRequiredIterations include/RansacLib/utils.h:110-140)
//   vanilla_msac() <- ransac_lib::VanillaMSAC::EstimateModel           evaluation/vanilla_ransac.h:23-99
//   legacy_msac()  <- sphericalsfm::MSAC::compute                      include/sphericalsfm/msac.h:67-131
//   preemptive_ransac() <- sphericalsfm::PreemptiveRANSAC::compute     include/sphericalsfm/preemptive_ransac.h:46-139
//
// Parity: PINNED.  tests/test_oracle_vs_ref.py runs this restatement and the reference's own
// RansacLib headers (compiled into oracle/_ref from /root/reference/include) on the same
// problems and demands bit-identical models, statistics and inlier sets.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <random>
#include <vector>

namespace ssfm_oracle {

struct Options {  // RansacOptions + LORansacOptions, ransac.h:47-92 (same defaults)
  uint32_t min_num_iterations = 100u;
  uint32_t max_num_iterations = 10000u;
  double success_probability = 0.9999;
  double squared_inlier_threshold = 1.0;
  unsigned int random_seed = 0u;
  int num_lo_steps = 10;
  double threshold_multiplier = std::sqrt(2.0);
  int num_lsq_iterations = 4;
  int min_sample_multiplicator = 7;
  int non_min_sample_multiplier = 3;
  uint32_t lo_starting_iterations = 50u;
  bool final_least_squares = false;
};

struct Statistics {  // RansacStatistics, ransac.h:94-101
  uint32_t num_iterations = 0;
  int best_num_inliers = 0;
  double best_model_score = std::numeric_limits<double>::max();
  double inlier_ratio = 0.0;
  std::vector<int> inlier_indices;
  int number_lo_iterations = 0;
};

inline uint32_t required_iterations(double w, double eta, int k, uint32_t lo, uint32_t hi) {  // utils.h:110-140
  if (w <= 0.0) return hi;
  if (w >= 1.0) return lo;
  const double miss = 1.0 - std::pow(w, (double)k);
  if (miss >= 0.99999999999999) return hi;
  const double n = std::ceil(std::log(eta) / std::log(miss) + 0.5);
  uint32_t it = std::min((uint32_t)n, hi);
  return std::max(lo, it);
}

// Fisher-Yates with the LO generator, then truncate (utils.h:34-52).  std::mt19937 and
// std::uniform_int_distribution are used as-is so the draw sequence is libstdc++'s.
inline void shuffle_and_resize(int target, std::mt19937* rng, std::vector<int>* v) {
  const int n = (int)v->size();
  for (int i = 0; i < n - 1; ++i) {
    std::uniform_int_distribution<int> dist(i, n - 1);
    std::swap((*v)[i], (*v)[dist(*rng)]);
  }
  v->resize(target);
}

template <class Solver>
struct Driver {
  typedef typename Solver::Model Model;
  const Solver& solver;
  const Options& opt;
  Driver(const Solver& s, const Options& o) : solver(s), opt(o) {}

  double msac_score(const Model& m, double thr) const {  // ScoreModel :295-303
    const int n = solver.num_data();
    double score = 0.0;
    for (int i = 0; i < n; ++i) score += std::min(solver.EvaluateModelOnPoint(m, i), thr);
    return score;
  }
  int inliers(const Model& m, double thr, std::vector<int>* out) const {  // GetInliers :311-336
    const int n = solver.num_data();
    int cnt = 0;
    if (out) out->clear();
    for (int i = 0; i < n; ++i)
      if (solver.EvaluateModelOnPoint(m, i) < thr) {
        ++cnt;
        if (out) out->push_back(i);
      }
    return cnt;
  }
  static void keep_better(double s, const Model& m, double* sb, Model* mb) {  // UpdateBestModel :422-428
    if (s < *sb) { *sb = s; *mb = m; }
  }
  void lsq_fit(double thresh, std::mt19937* rng, Model* m) const {  // LeastSquaresFit :409-420
    const int cap = opt.min_sample_multiplicator * solver.min_sample_size();
    std::vector<int> inl;
    const int n = inliers(*m, thresh, &inl);
    if (n < solver.min_sample_size()) return;
    shuffle_and_resize(std::min(cap, n), rng, &inl);
    solver.LeastSquares(inl, m);
  }
  void local_optimization(std::mt19937* rng, Model* best, double* best_score) const {  // :341-407
    const int n = solver.num_data();
    const int min_non_min = solver.non_minimal_sample_size();
    if (min_non_min > n) return;
    const int k = solver.min_sample_size();
    const double thr = opt.squared_inlier_threshold, mult = opt.threshold_multiplier;
    Model m_init = *best;
    lsq_fit(thr * mult, rng, &m_init);
    double score = msac_score(m_init, thr);
    keep_better(score, m_init, best_score, best);
    std::vector<int> base;
    inliers(m_init, thr * mult, &base);
    const int non_min_size = std::max(min_non_min, std::min(k * opt.non_min_sample_multiplier, (int)base.size() / 2));
    for (int r = 0; r < opt.num_lo_steps; ++r) {
      std::vector<int> sample = base;
      shuffle_and_resize(non_min_size, rng, &sample);
      Model m;
      if (!solver.NonMinimalSolver(sample, &m)) continue;
      score = msac_score(m, thr);
      keep_better(score, m, best_score, best);
      lsq_fit(thr, rng, &m);
      double th = mult * thr;
      const double dth = (mult - 1.0) * thr / (int)(opt.num_lsq_iterations - 1);
      for (int i = 0; i < opt.num_lsq_iterations; ++i) {
        lsq_fit(th, rng, &m);
        score = msac_score(m, thr);
        keep_better(score, m, best_score, best);
        th -= dth;
      }
    }
  }
  void refresh(const Model& m, Statistics* st, uint32_t* max_iters) const {  // :231-238
    st->best_num_inliers = inliers(m, opt.squared_inlier_threshold, &st->inlier_indices);
    st->inlier_ratio = (double)st->best_num_inliers / (double)solver.num_data();
    if (max_iters)
      *max_iters = required_iterations(st->inlier_ratio, 1.0 - opt.success_probability, solver.min_sample_size(),
                                       opt.min_num_iterations, opt.max_num_iterations);
  }
};

template <class Solver, class Sampler>
int lo_msac(const Options& opt, const Solver& solver, typename Solver::Model* best_model, Statistics* st) {
  typedef typename Solver::Model Model;
  const double kMax = std::numeric_limits<double>::max();
  *st = Statistics();
  const int k = solver.min_sample_size(), n = solver.num_data();
  if (k > n || k <= 0) return 0;
  Driver<Solver> d(solver, opt);
  Sampler sampler(opt.random_seed, solver);
  std::mt19937 rng;
  rng.seed(opt.random_seed);
  uint32_t max_iters = std::max(opt.max_num_iterations, opt.min_num_iterations);
  const double thr = opt.squared_inlier_threshold;
  Model best_min_model;
  double best_min_score = kMax;
  std::vector<int> sample(k);
  typename Solver::ModelVector models;
  for (st->num_iterations = 0u; st->num_iterations < max_iters; ++st->num_iterations) {
    const uint32_t it = st->num_iterations;
    if (it == opt.lo_starting_iterations && best_min_score < kMax) {  // :163-178
      ++st->number_lo_iterations;
      d.local_optimization(&rng, best_model, &st->best_model_score);
      d.refresh(*best_model, st, &max_iters);
    }
    sampler.Sample(&sample);
    const int nm = solver.MinimalSolver(sample, &models);
    if (nm <= 0) continue;
    double local_best = kMax;
    int local_id = 0;
    for (int m = 0; m < nm; ++m) {  // GetBestEstimatedModelId :278-293
      const double s = d.msac_score(models[m], thr);
      if (s < local_best) { local_best = s; local_id = m; }
    }
    if (local_best < best_min_score || it == opt.lo_starting_iterations) {  // :195-239
      const bool is_best = local_best < best_min_score;
      if (is_best) {
        best_min_score = local_best;
        best_min_model = models[local_id];
        Driver<Solver>::keep_better(best_min_score, best_min_model, &st->best_model_score, best_model);
      }
      const bool run_lo = (it >= opt.lo_starting_iterations && best_min_score < kMax);
      if (!is_best && !run_lo) continue;
      if (run_lo) {
        ++st->number_lo_iterations;
        double score = best_min_score;
        d.local_optimization(&rng, &best_min_model, &score);
        Driver<Solver>::keep_better(score, best_min_model, &st->best_model_score, best_model);
      }
      d.refresh(*best_model, st, &max_iters);
    }
  }
  if (st->num_iterations <= opt.lo_starting_iterations && st->best_model_score < kMax) {  // :245-255
    ++st->number_lo_iterations;
    d.local_optimization(&rng, best_model, &st->best_model_score);
    d.refresh(*best_model, st, nullptr);
  }
  if (opt.final_least_squares) {  // :257-272
    Model refined = *best_model;
    solver.LeastSquares(st->inlier_indices, &refined);
    const double score = d.msac_score(refined, thr);
    if (score < st->best_model_score) {
      st->best_model_score = score;
      *best_model = refined;
      d.refresh(*best_model, st, nullptr);
    }
  }
  return st->best_num_inliers;
}

template <class Solver, class Sampler>
int vanilla_msac(const Options& opt, const Solver& solver, typename Solver::Model* best_model, Statistics* st) {
  typedef typename Solver::Model Model;
  const double kMax = std::numeric_limits<double>::max();
  *st = Statistics();
  const int k = solver.min_sample_size(), n = solver.num_data();
  if (k > n || k <= 0) return 0;
  Driver<Solver> d(solver, opt);
  Sampler sampler(opt.random_seed, solver);
  uint32_t max_iters = std::max(opt.max_num_iterations, opt.min_num_iterations);
  const double thr = opt.squared_inlier_threshold;
  double best_min_score = kMax;
  std::vector<int> sample(k);
  typename Solver::ModelVector models;
  for (st->num_iterations = 0u; st->num_iterations < max_iters; ++st->num_iterations) {
    sampler.Sample(&sample);
    const int nm = solver.MinimalSolver(sample, &models);
    if (nm <= 0) continue;
    double local_best = kMax;
    int local_id = 0;
    for (int m = 0; m < nm; ++m) {
      const double s = d.msac_score(models[m], thr);
      if (s < local_best) { local_best = s; local_id = m; }
    }
    if (local_best < best_min_score) {  // vanilla_ransac.h:68-92
      best_min_score = local_best;
      const Model m = models[local_id];
      Driver<Solver>::keep_better(best_min_score, m, &st->best_model_score, best_model);
      d.refresh(*best_model, st, &max_iters);
    }
  }
  return st->best_num_inliers;
}

// Legacy fixed-budget MSAC (include/sphericalsfm/msac.h:67-131): hypothesis budget M
// (= estimators.size()), inlier test '<=' (:60), cost = sum(score if inlier else thr^2),
// adaptive stop num_iter = log(1-p)/log(1-(1-outlier_ratio)^m) capped at M (:119-126).
// The reference samples with random_sample (:6-27, Knuth 3.4.2S on rand()); `sample_fn(iteration, N, k, idx)` supplies
// it (knuth_sample on the Philox-backed rand()).
template <class Solver, class SampleFn>
int legacy_msac(const Options& opt, int budget, double prob_success, const Solver& solver, SampleFn sample_fn,
                typename Solver::Model* best_model, Statistics* st) {
  *st = Statistics();
  const int k = solver.min_sample_size(), n = solver.num_data();
  if (k > n || k <= 0) return 0;
  const double thr = opt.squared_inlier_threshold;
  double num_iter = budget;
  double best_score = INFINITY;
  int num_inliers = 0;
  std::vector<int> sample(k);
  typename Solver::ModelVector models;
  int iter = 0;
  while (iter < num_iter && iter < budget) {
    sample_fn((uint32_t)iter, n, k, sample.data());
    const int nm = solver.MinimalSolver(sample, &models);
    for (int m = 0; m < nm; ++m) {
      double score = 0.0;
      int cnt = 0;
      for (int i = 0; i < n; ++i) {
        const double e = solver.EvaluateModelOnPoint(models[m], i);
        if (e <= thr) { score += e; ++cnt; } else { score += thr; }
      }
      if (score < best_score) {
        best_score = score;
        num_inliers = cnt;
        *best_model = models[m];
      }
    }
    const double outlier_ratio = (n - num_inliers) / (double)n;
    if (outlier_ratio < 1.0) {
      num_iter = std::log(1. - prob_success) / std::log(1. - std::pow(1. - outlier_ratio, k));
      if (num_iter > budget) num_iter = budget;
    }
    ++iter;
  }
  st->num_iterations = iter;
  st->best_num_inliers = num_inliers;
  st->best_model_score = best_score;
  st->inlier_ratio = (double)num_inliers / (double)n;
  if (std::isfinite(best_score))
    for (int i = 0; i < n; ++i)
      if (solver.EvaluateModelOnPoint(*best_model, i) <= thr) st->inlier_indices.push_back(i);
  return num_inliers;
}

// Pre-emptive RANSAC (include/sphericalsfm/preemptive_ransac.h:46-139).
//   1. M hypotheses (= estimators.size()): a selection sample of m+1 correspondences (:66), the
//      minimal solver on the first m (:69), the extra one picks the solution with the smallest
//      error, first wins ties (:75-90).  A sample without a solution leaves a hypothesis that
//      is never an inlier anywhere (upstream leaves its E unset, :72); with exactly one solution
//      upstream never calls chooseSolution (:75) so E stays unset as well -- here it is that
//      solution (the evident intent).
//   2. Observations are visited in blocks of B (:100-107): every surviving hypothesis counts its
//      inliers ('<=', :105) on the block; after block i the survivors are the
//      f_i = floor(M 2^-floor(i/B)) best by (count, index) descending (:110-113, std::greater on
//      pair<int,size_t>); stop at one survivor or at the end of the data (:116-119).
//   3. The top hypothesis is scored on everything for the inlier mask (:122-137).
// `Sample4` draws the m+1 indices of hypothesis i: knuth_sample for the reference's scheme.
template <class Solver, class SampleFn>
int preemptive_ransac(const Options& opt, int M, int B, const Solver& solver, SampleFn sample_fn,
                      typename Solver::Model* best_model, Statistics* st) {
  typedef typename Solver::Model Model;
  *st = Statistics();
  const int m = solver.min_sample_size(), N = solver.num_data();
  if (m + 1 > N || m <= 0 || M <= 0 || B <= 0) return 0;
  const double thr = opt.squared_inlier_threshold;
  std::vector<Model> hyp(M);
  std::vector<char> has(M, 0);
  std::vector<std::pair<int, size_t> > order(M);
  std::vector<int> sample(m + 1), minimal(m);
  typename Solver::ModelVector models;
  for (int i = 0; i < M; ++i) {
    order[i].first = 0;
    order[i].second = (size_t)i;
    sample_fn((uint32_t)i, N, m + 1, sample.data());
    for (int k = 0; k < m; ++k) minimal[k] = sample[k];
    const int nsolns = solver.MinimalSolver(minimal, &models);
    if (nsolns == 0) continue;
    int best_index = 0;
    double best_score = INFINITY;
    if (nsolns > 1) {
      for (int j = 0; j < nsolns; ++j) {
        const double score = solver.EvaluateModelOnPoint(models[j], sample[m]);
        if (score < best_score) {
          best_score = score;
          best_index = j;
        }
      }
    }
    hyp[i] = models[best_index];
    has[i] = 1;
  }
  int it = 0;
  size_t f = (size_t)M;
  for (size_t i = 1; i < (size_t)N; ++i) {
    const int start = it;
    for (; it != start + B && it != N; ++it)
      for (size_t j = 0; j < f; ++j) {
        const size_t h = order[j].second;
        if (has[h] && solver.EvaluateModelOnPoint(hyp[h], it) <= thr) order[j].first++;
      }
    f = (size_t)std::floor(M * std::pow(2., -std::floor((double)(i / (size_t)B))));
    std::partial_sort(order.begin(), order.begin() + f, order.end(), std::greater<std::pair<int, size_t> >());
    if (f <= 1) break;
    if (it == N) break;
  }
  const size_t top = order[0].second;
  st->num_iterations = (uint32_t)M;
  if (!has[top]) {
    st->best_model_score = std::numeric_limits<double>::max();
    return 0;
  }
  *best_model = hyp[top];
  int num_inliers = 0;
  double cost = 0.0;  // not part of the reference's output: the legacy MSAC cost of the winner (msac.h:56-64)
  for (int i = 0; i < N; ++i) {
    const double score = solver.EvaluateModelOnPoint(*best_model, i);
    if (score <= thr) {
      st->inlier_indices.push_back(i);
      num_inliers++;
      cost += score;
    } else {
      cost += thr;
    }
  }
  st->best_num_inliers = num_inliers;
  st->best_model_score = cost;
  st->inlier_ratio = (double)num_inliers / (double)N;
  return num_inliers;
}

}  // namespace ssfm_oracle
