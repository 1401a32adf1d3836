// hostshim.cpp -- TEST-ONLY host build of the engine's __host__ __device__ arithmetic
// (spherical-sfm_b200/csrc/ssfm_math.cuh, ssfm_chain.cuh) so that `pytest -m "not gpu"` can
// check the product's solver / refit / chain logic against the oracle on a machine without a GPU.
// It is NOT part of libssfm_b200.so, is never loaded by the package, and is not a CPU fallback:
// the product entry points fail with SSFM_ERR_NO_DEVICE when CUDA is unavailable.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../spherical-sfm_b200/csrc/ssfm_chain.cuh"
#include "../../spherical-sfm_b200/csrc/ssfm_sixpt.cuh"
#include "../../spherical-sfm_b200/csrc/ssfm_sixpt_coop.cuh"
#include "../../spherical-sfm_b200/csrc/ssfm_sixpt_lo.cuh"
#include "../../spherical-sfm_b200/csrc/ssfm_triangulate.cuh"

using namespace ssfm;

namespace {
// plain-float emulation of the FP32 scoring kernel's per-iteration output
float score_iteration_f32(const double* models4x6, const double* rays, int n, float thr, float* per_model) {
  float best = INFINITY;
  for (int m = 0; m < 4; ++m) {
    float p[6];
    for (int i = 0; i < 6; ++i) p[i] = (float)models4x6[6 * m + i];
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
      const float u0 = (float)rays[6 * i], u1 = (float)rays[6 * i + 1], u2 = (float)rays[6 * i + 2];
      const float v0 = (float)rays[6 * i + 3], v1 = (float)rays[6 * i + 4], v2 = (float)rays[6 * i + 5];
      const float Eu0 = p[0] * u0 + p[1] * u1 + p[2] * u2;
      const float Eu1 = p[1] * u0 - p[0] * u1 + p[3] * u2;
      const float Eu2 = p[4] * u0 + p[5] * u1;
      const float Etv0 = p[0] * v0 + p[1] * v1 + p[4] * v2;
      const float Etv1 = p[1] * v0 - p[0] * v1 + p[5] * v2;
      const float d = v0 * Eu0 + v1 * Eu1 + v2 * Eu2;
      const float e = d * d / (Eu0 * Eu0 + Eu1 * Eu1 + Etv0 * Etv0 + Etv1 * Etv1);
      acc += fminf(e, thr);
    }
    if (!(models4x6[6 * m] == models4x6[6 * m])) acc = INFINITY;
    per_model[m] = acc;
    if (acc < best) best = acc;
  }
  return best;
}
}  // namespace

extern "C" {

struct HsParams {
  uint32_t min_iters, max_iters;
  double success_probability, thr2;
  uint32_t seed;
  int32_t num_lo_steps;
  double thr_mult;
  int32_t num_lsq_iters, min_sample_mult, non_min_mult;
  uint32_t lo_start;
  int32_t final_lsq, solver, driver, inward, fixed_budget;
  double fixed_prob;
  float cand_margin;
  int32_t first_round, round_cap;
  int32_t defer;  // 1: use the deferred-refit protocol (requires num_lo_steps == 0)
  int32_t skip_complex;  // SsfmOptions.complex_root_models == SSFM_COMPLEX_SKIP
};

struct HsResult {
  double E[9], r[3], t[3];
  double best_model_score, inlier_ratio;
  uint32_t num_iterations;
  int32_t best_num_inliers, num_lo, status;
  int64_t evals_exact;
  int32_t rounds, candidates;
};

void hs_sample(uint32_t seed, uint32_t pair, uint32_t iter, int k, int n, int* idx) {
  philox_sample<8>(seed, pair, iter, k, n, idx);
}

int hs_solve(const double* rays, const int* sample, int kind, int skip_complex, double* models) {
  const bool sk = skip_complex != 0;
  double m[4][6];
  const double* c0 = rays + 6 * (size_t)sample[0];
  const double* c1 = rays + 6 * (size_t)sample[1];
  const double* c2 = rays + 6 * (size_t)sample[2];
  int nm;
  if (kind == 0) nm = solve_minimal<0>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, sk);
  else if (kind == 1) nm = solve_minimal<1>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, sk);
  else nm = solve_minimal<2>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, sk);
  std::memcpy(models, m, sizeof(m));
  return nm;
}

void hs_quartic(const double* c, double* re, double* im) {
  Cplx r[4];
  quartic_roots(c[0], c[1], c[2], c[3], c[4], r);
  for (int i = 0; i < 4; ++i) { re[i] = r[i].re; im[i] = r[i].im; }
}

void hs_decompose(const double* E, int inward, double* r, double* t) { decompose_spherical_E(E, inward != 0, r, t); }

void hs_sampson(const double* E, const double* rays, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = sampson_exact(E, rays + 6 * (size_t)i, rays + 6 * (size_t)i + 3);
}

void hs_least_squares(const double* rays, const int* sample, int n, int inward, double* E) {
  SerialCtx cx;
  least_squares(cx, rays, sample, n, inward != 0, E);
}

// What k_refit_big / k_refit_small / k_refit_long do since round 2: gather the refit's correspondences once into a
// contiguous copy and run the same minimiser on it with sample == NULL (identity).  mode 1: plain; mode 2: the
// deferred path's arithmetic (one lane, lane-parallel after `handover` iterations -- with one lane both are the same sums).
void hs_least_squares_staged(const double* rays, const int* sample, int n, int inward, double* E, int mode, int handover) {
  SerialCtx cx;
  std::vector<double> staged((size_t)(n > 0 ? n : 1) * 6);
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < 6; ++q) staged[6 * (size_t)i + q] = rays[6 * (size_t)sample[i] + q];
  if (mode == 2) least_squares_as_deferred(cx, staged.data(), (const int*)0, n, inward != 0, E, handover);
  else least_squares(cx, staged.data(), (const int*)0, n, inward != 0, E);
}

void hs_lo_shuffle(uint32_t seed, int ncalls, const int* sizes, const int* targets, int* out) {
  std::vector<uint32_t> mt(625);
  mt19937_seed(mt.data(), seed);
  SerialCtx cx;
  int o = 0;
  for (int c = 0; c < ncalls; ++c) {
    std::vector<int> v(sizes[c]);
    for (int i = 0; i < sizes[c]; ++i) v[i] = i;
    shuffle_and_resize(cx, mt.data(), v.data(), sizes[c], targets[c]);
    for (int i = 0; i < targets[c]; ++i) out[o++] = v[i];
  }
}

uint32_t hs_required_iterations(double w, double eta, int k, uint32_t lo, uint32_t hi) {
  return required_iterations(w, eta, k, lo, hi);
}

// The engine's round structure, emulated serially for one pair.
void hs_estimate_pair(const double* rays, int n, const HsParams* hp, uint32_t pair_id, HsResult* out,
                      unsigned char* flags) {
  Params P{};
  P.min_points = 0;
  P.min_iters = hp->min_iters; P.max_iters = hp->max_iters;
  P.eta = 1.0 - hp->success_probability; P.thr2 = hp->thr2; P.seed = hp->seed;
  P.num_lo_steps = hp->num_lo_steps; P.thr_mult = hp->thr_mult; P.num_lsq_iters = hp->num_lsq_iters;
  P.min_sample_mult = hp->min_sample_mult; P.non_min_mult = hp->non_min_mult; P.lo_start = hp->lo_start;
  P.final_lsq = hp->final_lsq; P.solver = hp->solver; P.driver = hp->driver; P.inward = hp->inward;
  P.fixed_budget = hp->fixed_budget; P.fixed_prob = hp->fixed_prob; P.first_pair_id = 0;
  P.cand_margin = hp->cand_margin;
  P.skip_complex = hp->skip_complex;
  PairState st;
  init_state(P, n, st);
  std::vector<int> la(n + 16), lb(n + 16);
  std::vector<uint32_t> mt(625);
  mt19937_seed(mt.data(), P.seed);
  double lmE[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  Scratch sc{la.data(), lb.data(), mt.data(), nullptr, lmE};
  const bool defer = hp->defer != 0 && P.num_lo_steps == 0;
  PairView pv{rays, n, rays};
  SerialCtx cx;
  int rounds = 0, candidates = 0;
  std::vector<double> models;
  std::vector<float> s32, s32m;
  while (!st.done) {
    const uint32_t want = iterations_wanted(P, st);
    if (want == 0) { st.done = 1; break; }
    const int cap = rounds == 0 ? hp->first_round : hp->round_cap;
    const int na = (int)(want < (uint32_t)cap ? want : (uint32_t)cap);
    models.assign((size_t)na * 24, 0.0);
    s32.assign(na, 0.f);
    s32m.assign((size_t)na * 4, 0.f);
    for (int j = 0; j < na; ++j) {
      int idx[3];
      if (P.driver == 2) knuth_sample(P.seed, pair_id, st.it + j, n, 3, idx);
      else philox_sample<3>(P.seed, pair_id, st.it + j, 3, n, idx);
      double m[4][6];
      const double* c0 = rays + 6 * (size_t)idx[0];
      const double* c1 = rays + 6 * (size_t)idx[1];
      const double* c2 = rays + 6 * (size_t)idx[2];
      if (P.solver == 0) solve_minimal<0>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, P.skip_complex != 0);
      else if (P.solver == 1) solve_minimal<1>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, P.skip_complex != 0);
      else solve_minimal<2>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, m, P.skip_complex != 0);
      for (int k = 0; k < 24; ++k) models[(size_t)k * na + j] = (&m[0][0])[k];
      float pm[4];
      s32[j] = score_iteration_f32(&m[0][0], rays, n, (float)P.thr2, pm);
      for (int k = 0; k < 4; ++k) s32m[(size_t)k * na + j] = pm[k];
    }
    const long long before = st.evals_exact;
    if (defer) {
      for (;;) {
        process_round<SerialCtx, true>(cx, P, pv, sc, st, models.data(), na, s32.data(), s32m.data(), na);
        if (st.phase == PHASE_NONE) break;
        least_squares(cx, rays, sc.list_a, st.lm_n, P.inward != 0, sc.lm_E);  // what the refit kernel does
      }
    } else {
      process_round<SerialCtx, false>(cx, P, pv, sc, st, models.data(), na, s32.data(), s32m.data(), na);
    }
    candidates += (int)((st.evals_exact - before) / (n > 0 ? n : 1));
    ++rounds;
  }
  if (defer) {
    for (;;) {
      out->status = finalize_pair<SerialCtx, true>(cx, P, pv, sc, st, out->r, out->t, flags);
      if (out->status >= 0) break;
      least_squares(cx, rays, sc.list_a, st.lm_n, P.inward != 0, sc.lm_E);
    }
  } else {
    out->status = finalize_pair<SerialCtx, false>(cx, P, pv, sc, st, out->r, out->t, flags);
  }
  std::memcpy(out->E, st.E_best, sizeof(st.E_best));
  out->best_model_score = st.best_model_score;
  out->inlier_ratio = st.inlier_ratio;
  out->num_iterations = st.it;
  out->best_num_inliers = st.best_num_inliers;
  out->num_lo = st.num_lo;
  out->evals_exact = st.evals_exact;
  out->rounds = rounds;
  out->candidates = candidates;
}

// Retriangulate for one point (csrc/ssfm_triangulate.cuh) on the host, for tests without a GPU.
// cam_tr: per OBSERVATION t[3], r[3]; returns best_num_inliers.
int hs_triangulate(const double* cam_tr, const double* obs_xy, int n, double focal, const HsParams* hp, uint32_t point_id,
                   double* X, uint32_t* iterations, int* num_lo, uint32_t chunk) {
  Params P{};
  P.min_iters = hp->min_iters; P.max_iters = hp->max_iters; P.eta = 1.0 - hp->success_probability; P.thr2 = hp->thr2;
  P.seed = hp->seed; P.num_lo_steps = hp->num_lo_steps; P.thr_mult = hp->thr_mult; P.num_lsq_iters = hp->num_lsq_iters;
  P.min_sample_mult = hp->min_sample_mult; P.non_min_mult = hp->non_min_mult; P.lo_start = hp->lo_start;
  P.final_lsq = hp->final_lsq;
  std::vector<tri::Cam> cams(n > 0 ? n : 1);
  std::vector<int> oc(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) {
    tri::make_camera(cam_tr + 6 * i, cam_tr + 6 * i + 3, cams[i]);
    oc[i] = i;
  }
  std::vector<int> scratch(4 * (size_t)(n > 0 ? n : 1) + 16);
  std::vector<uint32_t> mt(625);
  tri::View v{cams.data(), oc.data(), obs_xy, n, focal};
  tri::Lists L{scratch.data(), scratch.data() + n, scratch.data() + 2 * n, scratch.data() + 3 * n, mt.data()};
  tri::Stats st;
  const int ninl = tri::lo_msac(P, v, L, point_id, X, st, chunk);
  *iterations = st.num_iterations;
  *num_lo = st.num_lo;
  return ninl;
}

// six-point shared-focal minimal solver (csrc/ssfm_sixpt.cuh) on the host, for tests without a GPU
int hs_sixpt_solve(const double* rays36, double* models /* 15 x 7: t, r, f */, double* G /* 15 x 9 */, int focal_scoring) {
  double c[6][6];
  for (int i = 0; i < 36; ++i) c[i / 6][i % 6] = rays36[i];
  SixPointModel out[kSixMaxModels];
  // the staged solver the batched kernel runs (here with a one-lane group); focal_scoring bit 1 selects the per-thread solver
  const int n = (focal_scoring & 2) ? solve_sixpt_focal(c, out) : sixc::solve_sixpt_focal_staged(c, out);
  focal_scoring &= 1;
  for (int k = 0; k < n; ++k) {
    for (int d = 0; d < 3; ++d) { models[7 * k + d] = out[k].t[d]; models[7 * k + 3 + d] = out[k].r[d]; }
    models[7 * k + 6] = out[k].f;
    sixpt_scoring_matrix(out[k], focal_scoring, G + 9 * k);
  }
  return n;
}

// SixPointEstimator::LeastSquares on the host (tests only): model = t[3], r[3], f; returns LM iterations
int hs_sixpt_least_squares(const double* rays, const int* sample, int n, double* model7, double* costs2) {
  SixPointModel m;
  for (int d = 0; d < 3; ++d) { m.t[d] = model7[d]; m.r[d] = model7[3 + d]; }
  m.f = model7[6];
  const SixLmSummary s = sixpt_least_squares(rays, sample, n, m);
  for (int d = 0; d < 3; ++d) { model7[d] = m.t[d]; model7[3 + d] = m.r[d]; }
  model7[6] = m.f;
  costs2[0] = s.initial_cost;
  costs2[1] = s.final_cost;
  return s.iterations;
}

// LO-MSAC around the six-point estimator on the host (tests only): the walk of k_sixpt_chain_lo, serially and without the
// FP32 pre-filter (every model of every iteration is scored in float64), around the very functions k_sixpt_lo runs
// (six_lo_phase / six_local_optimization, csrc/ssfm_sixpt_lo.cuh).  out7 = t, r, f.
int hs_sixpt_lo_msac(const double* rays, int n, const HsParams* hp, uint32_t pair_id, int focal_scoring, double* out7,
                     double* best_score, uint32_t* iterations, int* num_lo, unsigned char* flags) {
  Params P{};
  P.min_points = 0;
  P.min_iters = hp->min_iters; P.max_iters = hp->max_iters;
  P.eta = 1.0 - hp->success_probability; P.thr2 = hp->thr2; P.seed = hp->seed;
  P.num_lo_steps = hp->num_lo_steps; P.thr_mult = hp->thr_mult; P.num_lsq_iters = hp->num_lsq_iters;
  P.min_sample_mult = hp->min_sample_mult; P.non_min_mult = hp->non_min_mult; P.lo_start = hp->lo_start;
  P.final_lsq = hp->final_lsq; P.first_pair_id = 0; P.sixpt_focal_scoring = focal_scoring;
  SerialCtx cx;
  PairView pv{rays, n, rays};
  std::vector<int> la(n + 16), lb(n + 16);
  std::vector<uint32_t> mt(625);
  mt19937_seed(mt.data(), P.seed);
  SixScratch sc{la.data(), lb.data(), mt.data()};
  SixLoState st;
  six_lo_init(P, n, st);
  long long ev = 0;
  if (!st.done) {
    for (;;) {
      if (st.phase == SIX_PH_RESUME_BODY) {
        st.phase = SIX_PH_NONE;
      } else {
        if (st.it >= st.max_iters) break;
        if (st.it == P.lo_start && st.best_min_score < kDblMax && !st.lo_start_done) {
          st.phase = SIX_PH_LO_START;
          six_lo_phase(cx, P, pv, sc, st, (unsigned char*)0, &ev);
          continue;
        }
      }
      int idx[6];
      philox_sample<6>(P.seed, pair_id, st.it, 6, n, idx);
      double c[6][6];
      for (int i = 0; i < 6; ++i)
        for (int q = 0; q < 6; ++q) c[i][q] = rays[6 * (size_t)idx[i] + q];
      SixPointModel sol[kSixMaxModels];
      const int nm = sixc::solve_sixpt_focal_staged(c, sol);
      if (nm <= 0) { st.it += 1; continue; }
      double local_best = kDblMax;
      int local_id = 0, local_cnt = 0;
      double Gs[kSixMaxModels][9];
      for (int k = 0; k < nm; ++k) {
        sixpt_scoring_matrix(sol[k], focal_scoring, Gs[k]);
        int cnt = 0;
        const double s = msac_score_exact(cx, Gs[k], pv.stream, n, P.thr2, &cnt, &ev);
        if (s < local_best) { local_best = s; local_id = k; local_cnt = cnt; }
      }
      const bool better = local_best < st.best_min_score;
      if (better) {
        st.best_min_score = local_best;
        st.bestmin.m = sol[local_id];
        for (int q = 0; q < 9; ++q) st.bestmin.G[q] = Gs[local_id][q];
        st.bestmin.score = local_best;
        st.bestmin.cnt = local_cnt;
        six_keep_better(local_best, local_cnt, st.bestmin.m, st.bestmin.G, st.best);
      }
      const bool enter = better || st.it == P.lo_start;  // ransac.h:193-194
      const bool run_lo = st.it >= P.lo_start && st.best_min_score < kDblMax;
      st.it += 1;
      if (!enter || (!better && !run_lo)) continue;
      if (run_lo) {
        st.phase = SIX_PH_LO_BEST;
        six_lo_phase(cx, P, pv, sc, st, (unsigned char*)0, &ev);
        continue;
      }
      six_refresh(P, n, st, true);
    }
    st.phase = SIX_PH_FINAL;
    six_lo_phase(cx, P, pv, sc, st, flags, &ev);
  }
  for (int d = 0; d < 3; ++d) { out7[d] = st.best.m.t[d]; out7[3 + d] = st.best.m.r[d]; }
  out7[6] = st.best.m.f;
  *best_score = st.best.score;
  *iterations = st.it;
  *num_lo = st.num_lo;
  return st.best_num_inliers;
}

}  // extern "C"
