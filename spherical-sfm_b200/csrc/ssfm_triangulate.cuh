// ssfm_triangulate.cuh -- SfM::Retriangulate (src/sfm.cpp:156-192) for all points at once: one LO-MSAC per
// 3-D point over its observations with sphericalsfm::TriangulationEstimator
// (src/triangulation_estimator.cpp:46-127).  SURVEY.md 8f rank 3: the second user of the RansacLib estimator
// concept in the reference.
//
// Shape of the work: up to millions of tiny independent problems (3..~100 observations each), RansacLib's default
// LO schedule (10 LO steps x 4+1 least-squares fits at every new best model after iteration 50).  One THREAD owns
// one point and runs the whole LocallyOptimizedMSAC::EstimateModel loop (include/RansacLib/ransac.h:128-275);
// minimal samples are Philox-keyed by (seed, point, iteration), LO shuffles replay std::mt19937 exactly as in the
// pair path.  Everything is __host__ __device__ so tests/hostshim can run the same code on the CPU.
#pragma once
#include "ssfm_chain.cuh"

namespace ssfm {
namespace tri {

struct Cam {  // sphericalsfm::Pose (src/sfm_types.cpp:14-19): t, r and the rotation block of P = [so3exp(r) | t]
  double t[3], r[3], R[9];
  double Rc[9];  // the rotation ceres::AngleAxisRotatePoint(r, .) applies (the refit's model of the same pose)
};

// Fill R (so3exp, as Pose's constructor does) and Rc (ceres/rotation.h's angle-axis formula) from t, r.
SSFM_HD void make_camera(const double* t, const double* r, Cam& c) {
  for (int k = 0; k < 3; ++k) { c.t[k] = t[k]; c.r[k] = r[k]; }
  so3exp(c.r, c.R);
  double* Rc = c.Rc;
  const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2), ct = cos(th), st = sin(th), ti = 1.0 / th;
    const double wx = r[0] * ti, wy = r[1] * ti, wz = r[2] * ti, oc = 1.0 - ct;
    Rc[0] = ct + wx * wx * oc;       Rc[1] = wx * wy * oc - wz * st;  Rc[2] = wy * st + wx * wz * oc;
    Rc[3] = wz * st + wx * wy * oc;  Rc[4] = ct + wy * wy * oc;       Rc[5] = -wx * st + wy * wz * oc;
    Rc[6] = -wy * st + wx * wz * oc; Rc[7] = wx * st + wy * wz * oc;  Rc[8] = ct + wz * wz * oc;
  } else {
    Rc[0] = 1; Rc[1] = -r[2]; Rc[2] = r[1];
    Rc[3] = r[2]; Rc[4] = 1; Rc[5] = -r[0];
    Rc[6] = -r[1]; Rc[7] = r[0]; Rc[8] = 1;
  }
}

struct View {  // the TriangulationObservationList of one point
  const Cam* cams;
  const int* obs_cam;
  const double* obs_xy;
  int n;
  double focal;
};

struct Lists {  // per-point scratch, n ints each + the LO generator
  int* base;
  int* work;
  int* inl;
  int* best;
  uint32_t* mt;
};

// TriangulationEstimator::EvaluateModelOnPoint (:46-54): squared reprojection error, DBL_MAX behind the camera.
SSFM_HD double evaluate(const View& v, const double* X, int i) {
  const Cam& c = v.cams[v.obs_cam[i]];
  const double PX0 = add_rn(dot3_rn(c.R[0], X[0], c.R[1], X[1], c.R[2], X[2]), c.t[0]);
  const double PX1 = add_rn(dot3_rn(c.R[3], X[0], c.R[4], X[1], c.R[5], X[2]), c.t[1]);
  const double PX2 = add_rn(dot3_rn(c.R[6], X[0], c.R[7], X[1], c.R[8], X[2]), c.t[2]);
  if (PX2 < 0) return kDblMax;
  const double r0 = add_rn(mul_rn(v.focal, PX0 / PX2), -v.obs_xy[2 * i]);
  const double r1 = add_rn(mul_rn(v.focal, PX1 / PX2), -v.obs_xy[2 * i + 1]);
  return add_rn(mul_rn(r0, r0), mul_rn(r1, r1));
}

// NonMinimalSolver (:65-86): DLT.  The reference takes the last right-singular vector of the 2N x 4 matrix A;
// here the eigenvector of the smallest eigenvalue of A^T A (symmetric 4x4, cyclic Jacobi).
SSFM_HD_NOINLINE void non_minimal(const View& v, const int* sample, int ns, double* X) {
  double S[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i][j] = 0.0;
  for (int k = 0; k < ns; ++k) {
    const int i = sample[k];
    const Cam& c = v.cams[v.obs_cam[i]];
    const double px = v.obs_xy[2 * i] / v.focal, py = v.obs_xy[2 * i + 1] / v.focal;
    const double a0[4] = {c.R[6] * px - c.R[0], c.R[7] * px - c.R[1], c.R[8] * px - c.R[2], c.t[2] * px - c.t[0]};
    const double a1[4] = {c.R[6] * py - c.R[3], c.R[7] * py - c.R[4], c.R[8] * py - c.R[5], c.t[2] * py - c.t[1]};
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) S[a][b] += a0[a] * a0[b] + a1[a] * a1[b];
  }
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) {
      diag += S[i][i] * S[i][i];
      for (int j = i + 1; j < 4; ++j) off += S[i][j] * S[i][j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (S[p][q] == 0.0) continue;
        const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 4; ++k) {
          const double a = S[k][p], b = S[k][q];
          S[k][p] = cs * a - sn * b;
          S[k][q] = sn * a + cs * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = S[p][k], b = S[q][k];
          S[p][k] = cs * a - sn * b;
          S[q][k] = sn * a + cs * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = V[k][p], b = V[k][q];
          V[k][p] = cs * a - sn * b;
          V[k][q] = sn * a + cs * b;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (S[i][i] < S[best][best]) best = i;
  for (int k = 0; k < 3; ++k) X[k] = V[k][best] / V[3][best];
}

// Residuals of TriangulationError (:19-43) and their Jacobian w.r.t. the point (analytic: d(PX)/dX is the rotation
// ceres::AngleAxisRotatePoint applies).
SSFM_HD void residual_jac(const View& v, int i, const double* X, double* res, double (*J)[3]) {
  const Cam& c = v.cams[v.obs_cam[i]];
  const double* Rc = c.Rc;
  const double PX0 = Rc[0] * X[0] + Rc[1] * X[1] + Rc[2] * X[2] + c.t[0];
  const double PX1 = Rc[3] * X[0] + Rc[4] * X[1] + Rc[5] * X[2] + c.t[1];
  const double PX2 = Rc[6] * X[0] + Rc[7] * X[1] + Rc[8] * X[2] + c.t[2];
  const double iz = 1.0 / PX2;
  res[0] = v.focal * (PX0 * iz) - v.obs_xy[2 * i];
  res[1] = v.focal * (PX1 * iz) - v.obs_xy[2 * i + 1];
  if (J) {
    for (int k = 0; k < 3; ++k) {
      J[0][k] = v.focal * (Rc[k] - PX0 * iz * Rc[6 + k]) * iz;
      J[1][k] = v.focal * (Rc[3 + k] - PX1 * iz * Rc[6 + k]) * iz;
    }
  }
}

SSFM_HD bool cholesky_solve3(const double* H /* lower, packed 6 */, const double* diag_add, const double* g, double* x) {
  const double a00 = H[0] + diag_add[0], a10 = H[1], a11 = H[2] + diag_add[1], a20 = H[3], a21 = H[4], a22 = H[5] + diag_add[2];
  if (!(a00 > 0.0)) return false;
  const double l00 = sqrt(a00), l10 = a10 / l00, l20 = a20 / l00;
  const double d1 = a11 - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = sqrt(d1), l21 = (a21 - l20 * l10) / l11;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = sqrt(d2);
  const double y0 = g[0] / l00, y1 = (g[1] - l10 * y0) / l11, y2 = (g[2] - l20 * y0 - l21 * y1) / l22;
  x[2] = y2 / l22;
  x[1] = (y1 - l21 * x[2]) / l11;
  x[0] = (y0 - l10 * x[1] - l20 * x[2]) / l00;
  return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
}

// LeastSquares (:88-126): Ceres trust-region LM (Solver::Options defaults, max 200 iterations, 10 consecutive
// invalid steps) over the three coordinates; same restatement as lm_step in ssfm_chain.cuh.
SSFM_HD_NOINLINE void least_squares(const View& v, const int* sample, int ns, double* X) {
  double H[6], g[3], scale[3], diagonal[3] = {0, 0, 0}, x_cost = 0.0, gmax = 0.0;
  bool have_scale = false;
  auto eval_jac = [&](const double* xx) {
    for (int a = 0; a < 6; ++a) H[a] = 0.0;
    for (int a = 0; a < 3; ++a) g[a] = 0.0;
    double c = 0.0;
    for (int k = 0; k < ns; ++k) {
      double r[2], J[2][3];
      residual_jac(v, sample[k], xx, r, J);
      for (int q = 0; q < 2; ++q) {
        c += r[q] * r[q];
        g[0] += J[q][0] * r[q]; g[1] += J[q][1] * r[q]; g[2] += J[q][2] * r[q];
        H[0] += J[q][0] * J[q][0];
        H[1] += J[q][1] * J[q][0]; H[2] += J[q][1] * J[q][1];
        H[3] += J[q][2] * J[q][0]; H[4] += J[q][2] * J[q][1]; H[5] += J[q][2] * J[q][2];
      }
    }
    gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));  // gradient of the unscaled problem
    if (!have_scale) {
      scale[0] = 1.0 / (1.0 + sqrt(H[0])); scale[1] = 1.0 / (1.0 + sqrt(H[2])); scale[2] = 1.0 / (1.0 + sqrt(H[5]));
      have_scale = true;
    }
    g[0] *= scale[0]; g[1] *= scale[1]; g[2] *= scale[2];
    H[0] *= scale[0] * scale[0];
    H[1] *= scale[1] * scale[0]; H[2] *= scale[1] * scale[1];
    H[3] *= scale[2] * scale[0]; H[4] *= scale[2] * scale[1]; H[5] *= scale[2] * scale[2];
    return 0.5 * c;
  };
  auto eval_cost = [&](const double* xx) {
    double c = 0.0;
    for (int k = 0; k < ns; ++k) {
      double r[2];
      residual_jac(v, sample[k], xx, r, (double(*)[3])0);
      c += r[0] * r[0] + r[1] * r[1];
    }
    return 0.5 * c;
  };
  x_cost = eval_jac(X);
  if (!isfinite(x_cost)) return;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid = 0;
  for (int iteration = 0;;) {
    if (iteration >= 200) break;
    if (gmax <= 1e-10) break;
    if (radius < 1e-32) break;
    ++iteration;
    if (!reuse_diagonal) {
      diagonal[0] = fmin(fmax(H[0], 1e-6), 1e32);
      diagonal[1] = fmin(fmax(H[2], 1e-6), 1e32);
      diagonal[2] = fmin(fmax(H[5], 1e-6), 1e32);
    }
    const double dd[3] = {diagonal[0] / radius, diagonal[1] / radius, diagonal[2] / radius};
    double step[3];
    bool valid = cholesky_solve3(H, dd, g, step);
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (valid) {
      for (int a = 0; a < 3; ++a) step[a] = -step[a];
      const double sg = step[0] * g[0] + step[1] * g[1] + step[2] * g[2];
      const double sHs = H[0] * step[0] * step[0] + H[2] * step[1] * step[1] + H[5] * step[2] * step[2] +
                         2.0 * (H[1] * step[1] * step[0] + H[3] * step[2] * step[0] + H[4] * step[2] * step[1]);
      model_cost_change = -(sg + 0.5 * sHs);
      if (!(model_cost_change > 0.0)) valid = false;
    }
    if (!valid) {
      if (++invalid >= 10) break;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    double cand[3], step_norm = 0.0, x_norm = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double d = step[a] * scale[a];
      cand[a] = X[a] + d;
      step_norm += d * d;
      x_norm += X[a] * X[a];
    }
    step_norm = sqrt(step_norm);
    x_norm = sqrt(x_norm);
    double cand_cost = eval_cost(cand);
    if (!isfinite(cand_cost)) cand_cost = kDblMax;
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) break;
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= 1e-6 * x_cost) break;
    const double rho = cost_change / model_cost_change;
    if (rho > 1e-3) {
      for (int a = 0; a < 3; ++a) X[a] = cand[a];
      x_cost = eval_jac(X);
      const double tt = 2.0 * rho - 1.0;
      radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - tt * tt * tt));
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }
}

SSFM_HD double msac_score(const View& v, const double* X, double thr, long long* evals) {  // ScoreModel ransac.h:295-303
  double s = 0.0;
  for (int i = 0; i < v.n; ++i) {
    const double e = evaluate(v, X, i);
    s += (thr < e) ? thr : e;
  }
  *evals += v.n;
  return s;
}
SSFM_HD int inliers(const View& v, const double* X, double thr, int* out, long long* evals) {  // GetInliers :311-336
  int cnt = 0;
  for (int i = 0; i < v.n; ++i)
    if (evaluate(v, X, i) < thr) {
      if (out) out[cnt] = i;
      ++cnt;
    }
  *evals += v.n;
  return cnt;
}
SSFM_HD void keep(double s, const double* m, double* sb, double* mb) {  // UpdateBestModel :422-428
  if (s < *sb) {
    *sb = s;
    mb[0] = m[0]; mb[1] = m[1]; mb[2] = m[2];
  }
}

struct Stats {
  uint32_t num_iterations;
  int best_num_inliers, num_lo;
  double best_model_score, inlier_ratio;
  long long evals;
};

SSFM_HD void lsq_fit(const Params& P, const View& v, const Lists& L, double thresh, double* m, long long* evals) {  // :409-420
  const int cap = P.min_sample_mult * 2;
  const int n = inliers(v, m, thresh, L.inl, evals);
  if (n < 2) return;
  SerialCtx cx;
  shuffle_and_resize(cx, L.mt, L.inl, n, n < cap ? n : cap);
  least_squares(v, L.inl, n < cap ? n : cap, m);
}

SSFM_HD_NOINLINE void local_optimization(const Params& P, const View& v, const Lists& L, double* best, double* best_score,
                                         long long* evals) {  // ransac.h:341-407
  const int k = 2, min_non_min = 2;
  if (min_non_min > v.n) return;
  const double thr = P.thr2, mult = P.thr_mult;
  double m_init[3] = {best[0], best[1], best[2]};
  lsq_fit(P, v, L, thr * mult, m_init, evals);
  double score = msac_score(v, m_init, thr, evals);
  keep(score, m_init, best_score, best);
  const int nbase = inliers(v, m_init, thr * mult, L.base, evals);
  int non_min_size = k * P.non_min_mult < nbase / 2 ? k * P.non_min_mult : nbase / 2;
  if (non_min_size < min_non_min) non_min_size = min_non_min;
  SerialCtx cx;
  for (int r = 0; r < P.num_lo_steps; ++r) {
    for (int i = 0; i < nbase; ++i) L.work[i] = L.base[i];
    shuffle_and_resize(cx, L.mt, L.work, nbase, non_min_size);
    // std::vector::resize(non_min_size) on a shorter vector value-initialises the new entries (index 0)
    for (int i = nbase; i < non_min_size; ++i) L.work[i] = 0;
    double m[3];
    non_minimal(v, L.work, non_min_size, m);
    score = msac_score(v, m, thr, evals);
    keep(score, m, best_score, best);
    lsq_fit(P, v, L, thr, m, evals);
    double th = mult * thr;
    const double dth = (mult - 1.0) * thr / (double)(int)(P.num_lsq_iters - 1);
    for (int i = 0; i < P.num_lsq_iters; ++i) {
      lsq_fit(P, v, L, th, m, evals);
      score = msac_score(v, m, thr, evals);
      keep(score, m, best_score, best);
      th -= dth;
    }
  }
}

SSFM_HD void refresh(const Params& P, const View& v, const Lists& L, const double* m, Stats& st, uint32_t* max_iters) {  // :231-238
  st.best_num_inliers = inliers(v, m, P.thr2, L.best, &st.evals);
  st.inlier_ratio = (double)st.best_num_inliers / (double)v.n;
  if (max_iters) *max_iters = required_iterations(st.inlier_ratio, P.eta, 2, P.min_iters, P.max_iters);
}

// Loop state carried across launches: the kernel runs the points in phases (a point that needs many more
// iterations than its neighbours is continued in a later, densely packed launch instead of keeping its warp alive).
struct LoState {
  double X[3], best_min_model[3], best_min_score;
  Stats st;
  uint32_t max_iters;
  int started;
};

#if defined(__CUDA_ARCH__)
#define SSFM_TRI_ALL(p) __all_sync(0xffffffffu, (p))
#else
#define SSFM_TRI_ALL(p) (p)
#endif

// LocallyOptimizedMSAC::EstimateModel (ransac.h:128-275) with TriangulationEstimator, resumable: runs iterations
// until the loop ends or stats.num_iterations reaches stop_at.  Returns true when the estimate is complete
// (S.X, S.st final).
//
// Warp-synchronous form.  The expensive part of the loop is LocalOptimization (51 least-squares fits), which each
// point triggers at its own iterations; executed where it stands, one lane would refit while 31 wait.  So a lane
// that reaches a LocalOptimization call PAUSES there, the warp keeps running the cheap sample/solve/score
// iterations of the other lanes until all of them are paused (or finished), and then all pending
// LocalOptimizations run together.  Per point nothing changes: same calls, same order, same generator draws.
// ALL 32 lanes of a warp must call this function together (dummy lanes pass active = false).
SSFM_HD_NOINLINE bool lo_msac_run(const Params& P, const View& v, const Lists& L, uint32_t point_id, LoState& S, uint32_t stop_at,
                                  bool active = true) {
  Stats& st = S.st;
  double* X = S.X;
  const int n = v.n;
  if (active && !S.started) {
    S.started = 1;
    st.num_iterations = 0; st.best_num_inliers = 0; st.num_lo = 0; st.best_model_score = kDblMax; st.inlier_ratio = 0.0; st.evals = 0;
    X[0] = X[1] = X[2] = 0.0;
    S.best_min_model[0] = S.best_min_model[1] = S.best_min_model[2] = 0.0;
    S.best_min_score = kDblMax;
    S.max_iters = P.max_iters > P.min_iters ? P.max_iters : P.min_iters;
    if (n >= 2) mt19937_seed(L.mt, P.seed);
  }
  if (n < 2) active = false;  // ransac.h:137-139
  const double thr = P.thr2;
  enum { RUN = 0, LO_TOP = 1, LO_BEST = 2, END = 3, STOP = 4 };
  int pending = active ? RUN : END;
  bool sampled = false;  // the LO at the top of iteration lo_start has run; continue that iteration with its sample
  for (;;) {
    // ---- cheap section: iterate until a LocalOptimization is due, the loop ends or the phase limit is reached
    while (pending == RUN) {
      const uint32_t it = st.num_iterations;
      if (!sampled) {
        if (it >= S.max_iters) { pending = END; break; }
        if (it >= stop_at) { pending = STOP; break; }
        if (it == P.lo_start && S.best_min_score < kDblMax) { pending = LO_TOP; break; }  // :163-178
      }
      sampled = false;
      int sample[2];
      philox_sample<2>(P.seed, point_id, it, 2, n, sample);
      double m[3];
      non_minimal(v, sample, 2, m);  // MinimalSolver = NonMinimalSolver on the sample, always one model (:56-63)
      const double s = msac_score(v, m, thr, &st.evals);
      double local_best = kDblMax;
      if (s < local_best) local_best = s;  // GetBestEstimatedModelId over one model; a NaN score leaves DBL_MAX
      if (local_best < S.best_min_score || it == P.lo_start) {  // :195-239
        const bool is_best = local_best < S.best_min_score;
        if (is_best) {
          S.best_min_score = local_best;
          S.best_min_model[0] = m[0]; S.best_min_model[1] = m[1]; S.best_min_model[2] = m[2];
          keep(S.best_min_score, S.best_min_model, &st.best_model_score, X);
        }
        const bool run_lo = it >= P.lo_start && S.best_min_score < kDblMax;
        if (run_lo) { pending = LO_BEST; break; }
        if (is_best) refresh(P, v, L, X, st, &S.max_iters);
      }
      ++st.num_iterations;
    }
    // ---- every lane of the warp is paused or finished: run the pending LocalOptimizations together
    if (SSFM_TRI_ALL(pending == END || pending == STOP)) break;
    if (pending == LO_TOP) {
      ++st.num_lo;
      local_optimization(P, v, L, X, &st.best_model_score, &st.evals);
      refresh(P, v, L, X, st, &S.max_iters);
      sampled = true;  // resume inside iteration lo_start, after the check
      pending = RUN;
    } else if (pending == LO_BEST) {
      ++st.num_lo;
      double score = S.best_min_score;
      local_optimization(P, v, L, S.best_min_model, &score, &st.evals);
      keep(score, S.best_min_model, &st.best_model_score, X);
      refresh(P, v, L, X, st, &S.max_iters);
      ++st.num_iterations;
      pending = RUN;
    }
  }
  if (pending == STOP) return false;
  if (!active) return true;
  if (st.num_iterations <= P.lo_start && st.best_model_score < kDblMax) {  // :245-255
    ++st.num_lo;
    local_optimization(P, v, L, X, &st.best_model_score, &st.evals);
    refresh(P, v, L, X, st, (uint32_t*)0);
  }
  if (P.final_lsq) {  // :257-272
    double refined[3] = {X[0], X[1], X[2]};
    least_squares(v, L.best, st.best_num_inliers, refined);
    const double score = msac_score(v, refined, thr, &st.evals);
    if (score < st.best_model_score) {
      st.best_model_score = score;
      X[0] = refined[0]; X[1] = refined[1]; X[2] = refined[2];
      refresh(P, v, L, X, st, (uint32_t*)0);
    }
  }
  return true;
}

// In one go (tests/hostshim; `chunk` > 0 exercises the resumable path).  Returns best_num_inliers.
SSFM_HD int lo_msac(const Params& P, const View& v, const Lists& L, uint32_t point_id, double* X, Stats& st, uint32_t chunk = 0) {
  LoState S;
  S.started = 0;
  uint32_t stop = chunk ? chunk : 0xFFFFFFFFu;
  while (!lo_msac_run(P, v, L, point_id, S, stop)) stop += chunk;
  X[0] = S.X[0]; X[1] = S.X[1]; X[2] = S.X[2];
  st = S.st;
  return st.best_num_inliers;
}

}  // namespace tri
}  // namespace ssfm
