// ssfm_sixpt_lo.cuh -- LocallyOptimizedMSAC (include/RansacLib/ransac.h:119-430) around the six-point shared-focal
// estimator (examples/six_point_estimator.{h,cpp}): the pieces the VanillaMSAC driver of ssfm_sixpt_kernels.cuh does not
// have.  SixPointEstimator supplies NonMinimalSolver (:121-144) and LeastSquares (:146-192) so that this driver can run
// it; upstream's own tools only ever call it under VanillaMSAC (examples/run_six_point_ransac.cpp), so this is the
// "estimator concept is complete" row of SURVEY section 8, not a benchmark configuration.
//
// Split of the work.  The walk over a round's look-ahead iterations (FP32 pre-filter, FP64 certification, best-minimal /
// best-model bookkeeping) is one warp per pair (k_sixpt_chain_lo).  A pair that reaches a LocalOptimization call -- up to
// 1 + num_lo_steps * (2 + num_lsq_iterations) Ceres refits of <= 42 residuals, one six-point solve per LO step, ~100
// passes over the pair's correspondences, all sequential by construction (one mt19937 stream, each step depends on the
// previous one) -- is PARKED, and all parked pairs run their LO together in k_sixpt_lo, one WARP per pair: the passes and
// the refits' residual loops are spread over the lanes, the scalar parts (6 x 6 Cholesky, the six-point solve of an LO
// step) run redundantly on every lane from identical inputs.  (First cut: one thread per pair -- measured 490 pairs/s on
// 2000 x 1000 correspondences: a sub-pass holds at most 2048 pairs, i.e. 14 threads per SM.)  The same functions compile
// for the host (tests/hostshim) with SerialCtx.
#pragma once
#include "ssfm_chain.cuh"
#include "ssfm_sixpt.cuh"

namespace ssfm {

enum { SIX_PH_NONE = 0, SIX_PH_LO_START = 1, SIX_PH_LO_BEST = 2, SIX_PH_FINAL = 3, SIX_PH_RESUME_BODY = 4 };

// A model together with its scoring matrix (EvaluateModelOnPoint's E, or Kinv E Kinv with sixpt_focal_scoring) and the
// number of correspondences with err < thr counted by the pass that scored it.
struct SixScored {
  SixPointModel m;
  double G[9];
  double score;
  int cnt;
};

struct SixScratch {
  int* list_a;   // n ints
  int* list_b;   // n ints
  uint32_t* mt;  // 625 words
};

SSFM_HD void six_keep_better(double s, int c, const SixPointModel& m, const double* G, SixScored& best) {  // UpdateBestModel :422-428
  if (s < best.score) {
    best.score = s;
    best.cnt = c;
    best.m = m;
    for (int i = 0; i < 9; ++i) best.G[i] = G[i];
  }
}

// LeastSquaresFit (ransac.h:409-420): inliers at `thresh`, shuffled, at most min_sample_multiplicator * 6 of them refitted.
template <class Ctx>
SSFM_HD_NOINLINE void six_lsq_fit(const Ctx& cx, const Params& P, const PairView& pv, const SixScratch& sc, double thresh,
                                  SixPointModel& m, double* G, long long* evals) {
  const int cap = P.min_sample_mult * 6;
  const int n = collect_inliers(cx, G, pv.stream, pv.n, thresh, false, sc.list_a, (unsigned char*)0, evals);
  if (n < 6) return;
  const int k = n < cap ? n : cap;
  shuffle_and_resize(cx, sc.mt, sc.list_a, n, k);
  sixpt_least_squares(cx, pv.rays, sc.list_a, k, m);
  sixpt_scoring_matrix(m, P.sixpt_focal_scoring, G);
}

// SixPointEstimator::NonMinimalSolver (examples/six_point_estimator.cpp:121-144): the minimal solver on the first six
// entries of the sample (MinimalSolver reads sample[0..5] only, :99-107), the solution with the smallest summed
// EvaluateModelOnPoint over the whole sample wins (strict '<', first minimum).
SSFM_HD_NOINLINE bool six_non_minimal_solver(const Params& P, const PairView& pv, const int* sample, int ns, SixPointModel& out,
                                             double* Gout) {
  if (ns < 6) return false;
  double c[6][6];
  for (int i = 0; i < 6; ++i)
    for (int q = 0; q < 6; ++q) c[i][q] = pv.rays[6 * (size_t)sample[i] + q];
  SixPointModel sol[kSixMaxModels];
  const int nm = solve_sixpt_focal(c, sol);
  if (nm == 0) return false;
  double best_score = INFINITY;
  int best_ind = 0;
  for (int i = 0; i < nm; ++i) {
    double G[9];
    sixpt_scoring_matrix(sol[i], P.sixpt_focal_scoring, G);
    double score = 0.0;
    for (int j = 0; j < ns; ++j) {
      const double* ry = pv.rays + 6 * (size_t)sample[j];
      score += sampson_exact(G, ry, ry + 3);
    }
    if (score < best_score) { best_score = score; best_ind = i; }
  }
  out = sol[best_ind];
  sixpt_scoring_matrix(out, P.sixpt_focal_scoring, Gout);
  return true;
}

// LocalOptimization (ransac.h:341-407) with min_sample_size 6 and non_minimal_sample_size 7 (six_point_estimator.h:21-23).
template <class Ctx>
SSFM_HD_NOINLINE void six_local_optimization(const Ctx& cx, const Params& P, const PairView& pv, const SixScratch& sc,
                                             SixScored& best, long long* evals) {
  if (7 > pv.n) return;
  const double thr = P.thr2, mult = P.thr_mult;
  SixPointModel m_init = best.m;
  double G_init[9];
  for (int i = 0; i < 9; ++i) G_init[i] = best.G[i];
  six_lsq_fit(cx, P, pv, sc, thr * mult, m_init, G_init, evals);
  int cnt = 0;
  double score = msac_score_exact(cx, G_init, pv.stream, pv.n, thr, &cnt, evals);
  six_keep_better(score, cnt, m_init, G_init, best);
  if (P.num_lo_steps <= 0) return;
  const int nbase = collect_inliers(cx, G_init, pv.stream, pv.n, thr * mult, false, sc.list_b, (unsigned char*)0, evals);
  int non_min = 6 * P.non_min_mult;
  if (nbase / 2 < non_min) non_min = nbase / 2;
  if (non_min < 7) non_min = 7;
  if (non_min > pv.n) non_min = pv.n;  // cannot happen (7 <= n); keeps the scratch lists in bounds
  for (int r = 0; r < P.num_lo_steps; ++r) {
    for (int i = cx.lane(); i < nbase; i += cx.width()) sc.list_a[i] = sc.list_b[i];
    cx.sync();
    shuffle_and_resize(cx, sc.mt, sc.list_a, nbase, non_min);
    if (non_min > nbase) {  // std::vector::resize grows with zeros
      if (cx.lane() == 0)
        for (int i = nbase; i < non_min; ++i) sc.list_a[i] = 0;
      cx.sync();
    }
    SixPointModel m;
    double G[9];
    if (!six_non_minimal_solver(P, pv, sc.list_a, non_min, m, G)) continue;
    score = msac_score_exact(cx, G, pv.stream, pv.n, thr, &cnt, evals);
    six_keep_better(score, cnt, m, G, best);
    six_lsq_fit(cx, P, pv, sc, thr, m, G, evals);
    double th = mult * thr;
    const double dth = (mult - 1.0) * thr / (double)(int)(P.num_lsq_iters - 1);
    for (int i = 0; i < P.num_lsq_iters; ++i) {
      six_lsq_fit(cx, P, pv, sc, th, m, G, evals);
      score = msac_score_exact(cx, G, pv.stream, pv.n, thr, &cnt, evals);
      six_keep_better(score, cnt, m, G, best);
      th -= dth;
    }
  }
}

// The loop state of one pair under LO-MSAC (RansacStatistics + EstimateModel's locals, ransac.h:127-158).
struct SixLoState {
  SixScored best;     // *best_model, stats.best_model_score
  SixScored bestmin;  // best_minimal_model (may hold an LO-refined model, :217-218) -- .score is NOT best_min_model_score
  double best_min_score;  // best_min_model_score: minimal models only
  double inlier_ratio;
  long long evals;
  uint32_t it, max_iters;
  int best_num_inliers;
  int num_lo;
  int done;
  int phase;          // SIX_PH_*
  int resume_j;       // look-ahead slot at which the walk resumes
  int lo_start_done;  // the LO at lo_starting_iterations_ (:166-177) has run
  float runmin32;
};

SSFM_HD void six_lo_init(const Params& P, int n, SixLoState& st) {
  for (int i = 0; i < 9; ++i) st.best.G[i] = st.bestmin.G[i] = 0.0;
  for (int i = 0; i < 3; ++i) st.best.m.t[i] = st.best.m.r[i] = st.bestmin.m.t[i] = st.bestmin.m.r[i] = 0.0;
  st.best.m.f = st.bestmin.m.f = 0.0;
  st.best.score = st.bestmin.score = kDblMax;
  st.best.cnt = st.bestmin.cnt = 0;
  st.best_min_score = kDblMax;
  st.inlier_ratio = 0.0;
  st.evals = 0;
  st.it = 0;
  st.max_iters = P.max_iters > P.min_iters ? P.max_iters : P.min_iters;  // :146-147
  st.best_num_inliers = 0;
  st.num_lo = 0;
  st.done = (n < 6 || n < P.min_points) ? 1 : 0;  // :134-138 (+ the caller's min_num_points skip)
  st.phase = SIX_PH_NONE;
  st.resume_j = 0;
  st.lo_start_done = 0;
  st.runmin32 = INFINITY;
}

// GetInliers(best_model) + inlier ratio (+ NumRequiredIterations), ransac.h:231-238: the count was recorded by the pass
// that scored the model at the same threshold.
SSFM_HD void six_refresh(const Params& P, int n, SixLoState& st, bool update_max) {
  st.best_num_inliers = st.best.cnt;
  st.inlier_ratio = (double)st.best.cnt / (double)n;
  if (update_max) st.max_iters = required_iterations(st.inlier_ratio, P.eta, 6, P.min_iters, P.max_iters);
}

// What a parked pair is waiting for.  Returns true when the pair is finished (result written by the caller).
template <class Ctx>
SSFM_HD_NOINLINE bool six_lo_phase(const Ctx& cx, const Params& P, const PairView& pv, const SixScratch& sc, SixLoState& st,
                                   unsigned char* flags, long long* evals) {
  if (st.phase == SIX_PH_LO_START) {  // :166-177, on *best_model
    ++st.num_lo;
    six_local_optimization(cx, P, pv, sc, st.best, evals);
    six_refresh(P, pv.n, st, true);
    st.lo_start_done = 1;
    st.phase = SIX_PH_RESUME_BODY;  // the walk continues INSIDE the same iteration (no loop-condition check)
    return false;
  }
  if (st.phase == SIX_PH_LO_BEST) {  // :214-238, on best_minimal_model with a copy of best_min_model_score
    ++st.num_lo;
    st.bestmin.score = st.best_min_score;
    six_local_optimization(cx, P, pv, sc, st.bestmin, evals);
    six_keep_better(st.bestmin.score, st.bestmin.cnt, st.bestmin.m, st.bestmin.G, st.best);
    six_refresh(P, pv.n, st, true);
    st.phase = SIX_PH_NONE;
    return false;
  }
  // SIX_PH_FINAL: the loop has ended (:243-275)
  if (st.it <= P.lo_start && st.best.score < kDblMax) {
    ++st.num_lo;
    six_local_optimization(cx, P, pv, sc, st.best, evals);
    six_refresh(P, pv.n, st, false);
  }
  if (P.final_lsq && st.best.score < kDblMax) {
    // stats.inlier_indices = GetInliers(*best_model, thr); LeastSquares over all of them, no shuffle (:259-262)
    const int ni = collect_inliers(cx, st.best.G, pv.stream, pv.n, P.thr2, false, sc.list_a, (unsigned char*)0, evals);
    SixPointModel m = st.best.m;
    double G[9];
    sixpt_least_squares(cx, pv.rays, sc.list_a, ni, m);
    sixpt_scoring_matrix(m, P.sixpt_focal_scoring, G);
    int cnt = 0;
    const double score = msac_score_exact(cx, G, pv.stream, pv.n, P.thr2, &cnt, evals);
    if (score < st.best.score) {
      six_keep_better(score, cnt, m, G, st.best);
      six_refresh(P, pv.n, st, false);
    }
  }
  if (flags) {
    if (st.best.score < kDblMax) {
      collect_inliers(cx, st.best.G, pv.stream, pv.n, P.thr2, false, (int*)0, flags, evals);
    } else {
      for (int i = cx.lane(); i < pv.n; i += cx.width()) flags[i] = 0;
    }
  }
  st.phase = SIX_PH_NONE;
  st.done = 1;
  return true;
}

}  // namespace ssfm
