// ssfm_chain.cuh -- the float64 "certification" stage: everything in the reference's RANSAC loop
// that is sequential per pair, written once over an execution context Ctx:
//   WarpCtx   (ssfm_kernels.cu): one warp per image pair, lanes stride over correspondences,
//             sums by xor-butterfly shuffles (bit-identical on every lane -> uniform control flow)
//   SerialCtx (tests/hostshim): one host thread, sequential sums -> the reference's summation order;
//             TEST-ONLY, never part of libssfm_b200.so.
//
// Reference being restated (include/RansacLib/ransac.h): EstimateModel :128-275, ScoreModel :295-303,
// GetInliers :311-336, LocalOptimization :341-407, LeastSquaresFit :409-420, UpdateBestModel :422-428;
// VanillaMSAC (evaluation/vanilla_ransac.h:23-99); legacy MSAC (include/sphericalsfm/msac.h:67-131);
// SphericalEstimator::LeastSquares / NonMinimalSolver (src/spherical_estimator.cpp:86-157).
#pragma once
#include "ssfm_math.cuh"

namespace ssfm {

struct Params {
  uint32_t min_iters, max_iters;
  double eta;   // 1 - success_probability
  double thr2;  // squared_inlier_threshold
  uint32_t seed;
  int num_lo_steps;
  double thr_mult;
  int num_lsq_iters;
  int min_sample_mult;
  int non_min_mult;
  uint32_t lo_start;
  int final_lsq;
  int solver, driver, inward;
  int fixed_budget;
  double fixed_prob;
  uint32_t first_pair_id;
  int min_points;  // pairs with fewer correspondences are skipped
  int preempt_block;  // pre-emptive driver: B
  int sixpt_focal_scoring;  // six-point estimator: score with Kinv E Kinv instead of E
  float cand_margin;  // relative slack of the FP32 pre-filter (see process_round)
  int skip_complex;   // action-matrix solver: leave out the models of complex eigenvalues (SsfmOptions.complex_root_models)
  // Small batches run the LocalOptimization refits inline (no parked waves, no host round trips) but with the arithmetic
  // of the deferred path, so a pair's result does not depend on the size of the batch it is in: refits of at most
  // inline_small_max residuals run as ONE lane would run them (every lane redundantly: uniform) and switch to the
  // lane-parallel sums after inline_handover iterations -- exactly k_refit_small -> k_refit_long.  0: off.
  int inline_small_max;
  int inline_handover;
};

// Per-pair RANSAC state carried across rounds (RansacStatistics + the loop's locals).
struct PairState {
  double E_best[9];     // *best_model
  double E_bestmin[9];  // best_minimal_model
  double best_model_score;
  double best_min_score;
  double inlier_ratio;
  double legacy_num_iter;  // MSAC_FIXED: num_iter (msac.h:77)
  long long evals_exact;
  uint32_t it;         // stats.num_iterations
  uint32_t max_iters;  // max_num_iterations (adaptive)
  int best_num_inliers;
  int cnt_best;     // inliers (err < thr2) of E_best, known from the pass that scored it
  int cnt_bestmin;  // same for E_bestmin
  int num_lo;
  int done;
  int phase;    // PHASE_*: a deferred least-squares refit is outstanding for this pair
  int walk_j;   // slot of the current round at which process_round resumes
  int lm_n;     // number of residuals of the outstanding refit (indices in Scratch::list_a)
  float runmin32;  // running minimum of the FP32 per-iteration scores
};

// Deferred-refit protocol (pipeline options, num_lo_steps == 0).  Every LocalOptimization is then
// exactly: LeastSquaresFit (inliers at thr*mult -> shuffle -> <= 21 residuals -> LM) + one
// ScoreModel + UpdateBestModel (ransac.h:360-366).  The warp that walks a pair does the inlier
// collection and the shuffle, parks the pair with its refit described in (list_a, lm_n, lm_E),
// and a separate kernel solves ALL parked refits -- one thread per small problem -- before the
// walk resumes.  A single warp iterating one 6-parameter LM is latency bound (~10^5 cycles per
// refit); thousands of independent refits, one per lane, are not.
enum { PHASE_NONE = 0, PHASE_LO_START = 1, PHASE_LO_BEST = 2, PHASE_LO_LATE = 3, PHASE_FINAL_LSQ = 4 };

// Execution contexts.  lane()/width() stride the data-parallel loops over correspondences and sum()/sum_i()/sum_vec()
// reduce over the whole context (identical result on every thread -> uniform control flow).  slane()/swidth()/ballot()/
// prefix_min_excl()/min_f() are the scan view used on the short per-round score arrays.
struct SerialCtx {
  SSFM_HD int lane() const { return 0; }
  SSFM_HD int width() const { return 1; }
  SSFM_HD int slane() const { return 0; }
  SSFM_HD int swidth() const { return 1; }
  SSFM_HD double sum(double x) const { return x; }
  SSFM_HD int sum_i(int x) const { return x; }
  template <int K>
  SSFM_HD void sum_vec(double (&v)[K]) const { (void)v; }
  SSFM_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
  SSFM_HD float prefix_min_excl(float x, float init) const { (void)x; return init; }
  SSFM_HD float min_f(float x) const { return x; }
  SSFM_HD void sync() const {}
};

// One correspondence = 6 doubles at a 16-byte aligned address (48-byte records): three 128-bit loads
// (LDG.128) instead of six 64-bit ones.
SSFM_HD void load6(const double* p, double (&c)[6]) {
#if defined(__CUDA_ARCH__)
  const double2 a = reinterpret_cast<const double2*>(p)[0];
  const double2 b = reinterpret_cast<const double2*>(p)[1];
  const double2 d = reinterpret_cast<const double2*>(p)[2];
  c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = d.x; c[5] = d.y;
#else
  for (int i = 0; i < 6; ++i) c[i] = p[i];
#endif
}

// What the full passes over a pair read.  Pipeline rays are K^-1 (x, y, 1): z == 1 exactly, so the passes only need the
// four doubles (u.x, u.y, v.x, v.y) of a correspondence -- 32 bytes instead of 48 from HBM, and k_chain is HBM-bound
// (ncu: 4.6 TB/s).  `stream` is either the 48-byte records themselves or the compact plane with bit 0 of the address set
// (records are 16-byte aligned); load_stream rebuilds the same six doubles either way, so the arithmetic is unchanged.
SSFM_HD const double* compact_stream(const double* xy4) { return reinterpret_cast<const double*>(reinterpret_cast<uintptr_t>(xy4) | 1u); }
SSFM_HD void load_stream(const double* stream, size_t i, double (&c)[6]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(stream);
  if (a & 1u) {
    const double* p = reinterpret_cast<const double*>(a - 1u) + 4 * i;
#if defined(__CUDA_ARCH__)
    const double2 u = reinterpret_cast<const double2*>(p)[0];
    const double2 v = reinterpret_cast<const double2*>(p)[1];
    c[0] = u.x; c[1] = u.y; c[2] = 1.0; c[3] = v.x; c[4] = v.y; c[5] = 1.0;
#else
    c[0] = p[0]; c[1] = p[1]; c[2] = 1.0; c[3] = p[2]; c[4] = p[3]; c[5] = 1.0;
#endif
  } else {
    load6(stream + 6 * i, c);
  }
}

// ScoreModel (ransac.h:295-303) fused with the inlier count GetInliers would return for the same
// model and threshold (:311-336): one pass yields both, so the reference's "GetInliers(best_model)"
// refresh needs no second pass over the data.
template <class Ctx>
SSFM_HD_NOINLINE double msac_score_exact(const Ctx& cx, const double* E, const double* rays, int n, double thr, int* count,
                                         long long* evals) {
  double s = 0.0;
  int c = 0;
#pragma unroll 2
  for (int i = cx.lane(); i < n; i += cx.width()) {
    double ry[6];
    load_stream(rays, (size_t)i, ry);
    const double e = sampson_exact(E, ry, ry + 3);
    s += (thr < e) ? thr : e;  // std::min(e, thr) incl. its NaN behaviour (ransac.h:306-309)
    c += (e < thr) ? 1 : 0;
  }
  *evals += n;
  *count = cx.sum_i(c);
  return cx.sum(s);
}

// GetInliers: count (and optionally list, in index order) the points with err < thr
// (or <= thr for the legacy driver, msac.h:60).
template <class Ctx>
SSFM_HD_NOINLINE int collect_inliers(const Ctx& cx, const double* E, const double* rays, int n, double thr, bool inclusive,
                            int* list, unsigned char* flags, long long* evals) {
  int count = 0;
  for (int base = 0; base < n; base += cx.width()) {
    const int i = base + cx.lane();
    bool in = false;
    if (i < n) {
      double ry[6];
      load_stream(rays, (size_t)i, ry);
      const double e = sampson_exact(E, ry, ry + 3);
      in = inclusive ? (e <= thr) : (e < thr);
      if (flags) flags[i] = in ? 1 : 0;
    }
    const unsigned m = cx.ballot(in);
    if (list && in) {
      const unsigned below = cx.width() == 1 ? 0u : (m & ((1u << cx.lane()) - 1u));
#if defined(__CUDA_ARCH__)
      list[count + __popc(below)] = i;
#else
      list[count + __builtin_popcount(below)] = i;
#endif
    }
#if defined(__CUDA_ARCH__)
    count += __popc(m);
#else
    count += __builtin_popcount(m);
#endif
  }
  *evals += n;
  cx.sync();
  return count;
}

// mt19937 state regeneration, collectively: 32 elements at a time, reads before writes, which
// preserves the in-place sequential recurrence (element i needs old i, old i+1 and element
// (i+397) mod 624, which is old for i < 227 and already regenerated otherwise).
template <class Ctx>
SSFM_HD void mt_twist_ctx(const Ctx& cx, uint32_t* mt) {
  if (cx.width() == 1) {
    mt19937_twist(mt);
    return;
  }
  for (int base = 0; base < 624; base += 32) {
    const int i = base + cx.lane();
    uint32_t v = 0;
    if (i < 624) {
      const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
      v = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    cx.sync();
    if (i < 624) mt[i] = v;
    cx.sync();
  }
}
SSFM_HD uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// RandomShuffleAndResize (include/RansacLib/utils.h:34-52) with the LO generator: Fisher-Yates over all
// n entries, then truncation to `keep`.  Swaps at positions i >= keep only touch entries that are
// thrown away, so only the first `keep` swaps are materialised (every lane computes the same draw,
// lane 0 swaps); the remaining n-1-keep draws are just CONSUMED so that the generator ends in
// exactly the state the reference's would: lanes test 32 draws at a time for Lemire's rejection
// condition (probability ~range/2^32 each) and replay a chunk serially only if one triggers.
template <class Ctx>
SSFM_HD_NOINLINE void shuffle_and_resize(const Ctx& cx, uint32_t* mt, int* list, int n, int keep) {
  if (n < 2) return;
  uint32_t pos = mt[624];
  const int lim = keep < n - 1 ? keep : n - 1;
  // uniform_int_distribution<int>(i, n-1)(mt19937): libstdc++ (GCC >= 11), Lemire's method
  auto draw = [&](int i) -> int {
    const uint32_t range = (uint32_t)(n - 1 - i) + 1u;
    uint64_t product;
    for (;;) {
      if (pos >= 624) {
        mt_twist_ctx(cx, mt);
        pos = 0;
      }
      product = (uint64_t)mt_temper(mt[pos++]) * (uint64_t)range;
      const uint32_t low = (uint32_t)product;
      if (low >= range) break;
      if (low >= (0u - range) % range) break;
    }
    return i + (int)(uint32_t)(product >> 32);
  };
  for (int i = 0; i < lim; ++i) {
    const int j = draw(i);
    if (cx.lane() == 0) {
      const int tmp = list[i];
      list[i] = list[j];
      list[j] = tmp;
    }
  }
  cx.sync();
  int i = lim;
  const int W = cx.width();
  while (i < n - 1) {
    if (pos >= 624) {
      mt_twist_ctx(cx, mt);
      pos = 0;
    }
    int chunk = n - 1 - i;
    if (chunk > W) chunk = W;
    if ((int)(624 - pos) < chunk) chunk = (int)(624 - pos);
    bool reject = false;
    if (W > 1) {
      const int l = cx.lane();
      if (l < chunk) {
        const uint32_t y = mt_temper(mt[pos + l]);
        const uint32_t range = (uint32_t)(n - 1 - (i + l)) + 1u;
        const uint32_t low = (uint32_t)((uint64_t)y * (uint64_t)range);
        reject = low < range && low < ((0u - range) % range);
      }
    }
    if (W > 1 && cx.ballot(reject) == 0u) {
      pos += (uint32_t)chunk;
    } else {
      for (int k = 0; k < chunk; ++k) (void)draw(i + k);
    }
    i += chunk;
  }
  cx.sync();
  if (cx.lane() == 0) mt[624] = pos;
  cx.sync();
}

// SphericalEstimator::LeastSquares (src/spherical_estimator.cpp:110-157): Ceres 2.2 trust-region
// LM (dense normal Cholesky, Jacobi scaling, default tolerances, <= 200 iterations) over the
// residuals listed in `sample`; E is rebuilt from the optimised rotation only (:156).
// Written as init / step / finish so that a kernel can interleave many refits per warp (a lane
// that converges picks up the next task while its neighbours keep iterating).
struct LMState {
  double x[6], H[21], g[6], scale[6], diagonal[6];
  double x_cost, gmax, radius, decrease_factor, t0z;
  int iteration, invalid;
  bool reuse_diagonal, have_scale;
};

// cost, gradient and J^T J at S.x with Jacobi scaling.  E(x) and dE/dx are uniform across the
// context (computed once per call as jets); each residual then costs ~200 flops.
template <class Ctx>
SSFM_HD double lm_eval_jac(const Ctx& cx, const double* rays, const int* sample, int n, LMState& S) {
  Jet6 Ej[9];
  {
    const Jet6 r1[3] = {jvar(S.x[0], 0), jvar(S.x[1], 1), jvar(S.x[2], 2)};
    const Jet6 t1[3] = {jvar(S.x[3], 3), jvar(S.x[4], 4), jvar(S.x[5], 5)};
    spherical_E_of_params<Jet6>(r1, t1, S.t0z, Ej);
  }
  double acc[28];  // [0,21): J^T J (lower triangle), [21,27): J^T r, [27]: r^T r -- reduced over the context in one go
  for (int a = 0; a < 28; ++a) acc[a] = 0.0;
  for (int i = cx.lane(); i < n; i += cx.width()) {
    double ry[6];
    load6(rays + 6 * (size_t)(sample ? sample[i] : i), ry);  // sample == NULL: the residuals are rays[0..n) themselves
    double r, jr[6];
    sampson_value_grad(Ej, ry, ry + 3, r, jr);
    acc[27] += r * r;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      acc[21 + a] += jr[a] * r;
#pragma unroll
      for (int b = 0; b <= a; ++b) acc[k++] += jr[a] * jr[b];
    }
  }
  cx.sum_vec(acc);
  for (int a = 0; a < 21; ++a) S.H[a] = acc[a];
  for (int a = 0; a < 6; ++a) S.g[a] = acc[21 + a];
  double c = acc[27];
  S.gmax = 0.0;
  for (int a = 0; a < 6; ++a) S.gmax = fmax(S.gmax, fabs(S.g[a]));  // gradient of the UNSCALED problem
  if (!S.have_scale) {
    for (int a = 0; a < 6; ++a) S.scale[a] = 1.0 / (1.0 + sqrt(S.H[a * (a + 1) / 2 + a]));
    S.have_scale = true;
  }
  int k = 0;
  for (int a = 0; a < 6; ++a) {
    S.g[a] *= S.scale[a];
    for (int b = 0; b <= a; ++b) S.H[k++] *= S.scale[a] * S.scale[b];
  }
  return 0.5 * c;
}
template <class Ctx>
SSFM_HD double lm_eval_cost(const Ctx& cx, const double* rays, const int* sample, int n, double t0z, const double* xx) {
  double Ev[9];
  spherical_E_of_params<double>(xx, xx + 3, t0z, Ev);
  double c = 0.0;
  for (int i = cx.lane(); i < n; i += cx.width()) {
    double ry[6];
    load6(rays + 6 * (size_t)(sample ? sample[i] : i), ry);  // sample == NULL: the residuals are rays[0..n) themselves
    const double r = sampson_value(Ev, ry, ry + 3);
    c += r * r;
  }
  return 0.5 * cx.sum(c);
}

template <class Ctx>
SSFM_HD_NOINLINE void lm_init(const Ctx& cx, const double* rays, const int* sample, int n, bool inward, const double* E,
                              LMState& S) {
  {
    double r[3], t[3];
    decompose_spherical_E(E, inward, r, t);
    S.x[0] = r[0]; S.x[1] = r[1]; S.x[2] = r[2];
    S.x[3] = 0.0; S.x[4] = 0.0; S.x[5] = inward ? 1.0 : -1.0;
  }
  S.t0z = inward ? 1.0 : -1.0;
  S.have_scale = false;
  S.radius = 1e4;
  S.decrease_factor = 2.0;
  S.reuse_diagonal = false;
  S.invalid = 0;
  S.iteration = 0;
  S.x_cost = lm_eval_jac(cx, rays, sample, n, S);
}

// One trust-region iteration.  Returns true when the minimiser terminates.
template <class Ctx>
SSFM_HD_NOINLINE bool lm_step(const Ctx& cx, const double* rays, const int* sample, int n, LMState& S) {
  if (!isfinite(S.x_cost)) return true;
  if (S.iteration >= 200) return true;
  if (S.gmax <= 1e-10) return true;       // gradient tolerance
  if (S.radius < 1e-32) return true;
  ++S.iteration;
  if (!S.reuse_diagonal)
    for (int a = 0; a < 6; ++a) S.diagonal[a] = fmin(fmax(S.H[a * (a + 1) / 2 + a], 1e-6), 1e32);
  double Hd[21], step[6];
  for (int a = 0; a < 21; ++a) Hd[a] = S.H[a];
  for (int a = 0; a < 6; ++a) Hd[a * (a + 1) / 2 + a] += S.diagonal[a] / S.radius;
  bool valid = cholesky_solve6(Hd, S.g, step);
  S.reuse_diagonal = true;
  double model_cost_change = 0.0;
  if (valid) {
    // model_cost_change = -(J s).(r + J s / 2) = -(s.g + s^T H s / 2)
    double sg = 0.0, sHs = 0.0;
    int k = 0;
    for (int a = 0; a < 6; ++a) {
      step[a] = -step[a];
      sg += step[a] * S.g[a];
    }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b <= a; ++b) sHs += (a == b ? 1.0 : 2.0) * S.H[k++] * step[a] * step[b];
    model_cost_change = -(sg + 0.5 * sHs);
    if (!(model_cost_change > 0.0)) valid = false;
  }
  if (!valid) {
    if (++S.invalid >= 10) return true;
    S.radius /= S.decrease_factor;
    S.decrease_factor *= 2.0;
    return false;
  }
  S.invalid = 0;
  double cand[6], step_norm = 0.0, x_norm = 0.0;
  for (int a = 0; a < 6; ++a) {
    const double d = step[a] * S.scale[a];
    cand[a] = S.x[a] + d;
    step_norm += d * d;
    x_norm += S.x[a] * S.x[a];
  }
  step_norm = sqrt(step_norm);
  x_norm = sqrt(x_norm);
  double cand_cost = lm_eval_cost(cx, rays, sample, n, S.t0z, cand);
  if (!isfinite(cand_cost)) cand_cost = kDblMax;
  if (step_norm <= 1e-8 * (x_norm + 1e-8)) return true;   // parameter tolerance
  const double cost_change = S.x_cost - cand_cost;
  if (fabs(cost_change) <= 1e-6 * S.x_cost) return true;  // function tolerance
  const double rho = cost_change / model_cost_change;
  if (rho > 1e-3) {
    for (int a = 0; a < 6; ++a) S.x[a] = cand[a];
    S.x_cost = lm_eval_jac(cx, rays, sample, n, S);
    const double tt = 2.0 * rho - 1.0;
    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - tt * tt * tt);
    S.radius = fmin(1e16, S.radius);
    S.decrease_factor = 2.0;
    S.reuse_diagonal = false;
  } else {
    S.radius /= S.decrease_factor;
    S.decrease_factor *= 2.0;
    S.reuse_diagonal = true;
  }
  return false;
}

SSFM_HD void lm_finish(const LMState& S, bool inward, double* E) {
  double R[9];
  so3exp(S.x, R);
  make_spherical_E(R, inward, E);
}

template <class Ctx>
SSFM_HD_NOINLINE void least_squares(const Ctx& cx, const double* rays, const int* sample, int n, bool inward, double* E) {
  LMState S;
  lm_init(cx, rays, sample, n, inward, E, S);
  while (!lm_step(cx, rays, sample, n, S)) {
  }
  lm_finish(S, inward, E);
}

// The deferred path's small refit (k_refit_small, then k_refit_long after `handover` iterations), inline.
template <class Ctx>
SSFM_HD_NOINLINE void least_squares_as_deferred(const Ctx& cx, const double* rays, const int* sample, int n, bool inward, double* E,
                                                int handover) {
  SerialCtx one;
  LMState S;
  lm_init(one, rays, sample, n, inward, E, S);
  bool done = false;
  for (;;) {
    if (lm_step(one, rays, sample, n, S)) { done = true; break; }
    if (handover > 0 && S.iteration >= handover) break;
  }
  if (!done)
    while (!lm_step(cx, rays, sample, n, S)) {
    }
  lm_finish(S, inward, E);
}

struct Scratch {
  int* list_a;   // n ints
  int* list_b;   // n ints
  uint32_t* mt;  // 625 words
  unsigned long long* prof;  // optional per-phase cycle counters (profiling builds only)
  double* lm_E;  // 9 doubles: model handed to / returned by a deferred refit
};

// Phase timers for profiling builds (-DSSFM_PROFILE_CHAIN): cycles spent per phase, summed over warps.
#if defined(SSFM_PROFILE_CHAIN) && defined(__CUDA_ARCH__)
#define SSFM_TIC long long tic__ = clock64();
#define SSFM_TOC(sc, k) if ((sc).prof && cx.lane() == 0) atomicAdd(&(sc).prof[k], (unsigned long long)(clock64() - tic__));
#else
#define SSFM_TIC
#define SSFM_TOC(sc, k)
#endif
enum { PH_SCAN = 8, PH_RESCORE, PH_LO_COLLECT, PH_LO_SHUFFLE, PH_LO_LM, PH_LO_SCORE, PH_FINAL_LM, PH_FINAL_REST, PH_TOTAL };

struct PairView {
  const double* rays;    // 6 doubles per correspondence
  int n;
  const double* stream;  // what msac_score_exact / collect_inliers read: rays, or compact_stream(xy plane) (load_stream)
};

template <class Ctx>
SSFM_HD_NOINLINE void lsq_fit(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, double thresh, double* E,
                     long long* evals) {  // LeastSquaresFit, ransac.h:409-420
  const int cap = P.min_sample_mult * 3;
  int n;
  { SSFM_TIC n = collect_inliers(cx, E, pv.stream, pv.n, thresh, false, sc.list_a, (unsigned char*)0, evals); SSFM_TOC(sc, PH_LO_COLLECT) }
  if (n < 3) return;
  { SSFM_TIC shuffle_and_resize(cx, sc.mt, sc.list_a, n, n < cap ? n : cap); SSFM_TOC(sc, PH_LO_SHUFFLE) }
  const int nr = n < cap ? n : cap;
  SSFM_TIC
  if (P.inline_small_max > 0 && nr <= P.inline_small_max) least_squares_as_deferred(cx, pv.rays, sc.list_a, nr, P.inward != 0, E, P.inline_handover);
  else least_squares(cx, pv.rays, sc.list_a, nr, P.inward != 0, E);
  SSFM_TOC(sc, PH_LO_LM)
}

SSFM_HD void keep_better(double s, const double* m, int c, double* sb, double* mb, int* cb) {  // UpdateBestModel :422-428
  if (s < *sb) {
    *sb = s;
    *cb = c;
    for (int i = 0; i < 9; ++i) mb[i] = m[i];
  }
}

// SphericalEstimator::NonMinimalSolver (src/spherical_estimator.cpp:86-108).  The reference runs
// the action-matrix solver on the whole sample; with its pivoted QR that is the minimal solve on
// the three greedily pivoted correspondences (largest remaining epipolar-row norm), see
// nullspace_colpiv.  Picks the model with the smallest summed Sampson error over the sample.
template <class Ctx>
SSFM_HD_NOINLINE bool non_minimal_solver(const Ctx& cx, const PairView& pv, const int* sample, int ns, double* E,
                                         bool skip_complex = false) {
  if (ns < 3) return false;
  // greedy pivoting by modified Gram-Schmidt on the epipolar rows (ns <= 64 handled serially & uniformly)
  int pick[3] = {-1, -1, -1};
  double q[3][6];
  for (int k = 0; k < 3; ++k) {
    double best = -1.0;
    int bi = -1;
    for (int i = 0; i < ns; ++i) {
      if (i == pick[0] || i == pick[1]) continue;
      double a[6];
      const double* ry = pv.rays + 6 * (size_t)sample[i];
      epipolar_row(ry, ry + 3, a);
      for (int j = 0; j < k; ++j) {
        double d = 0.0;
        for (int r = 0; r < 6; ++r) d += q[j][r] * a[r];
        for (int r = 0; r < 6; ++r) a[r] -= d * q[j][r];
      }
      double s = 0.0;
      for (int r = 0; r < 6; ++r) s += a[r] * a[r];
      if (s > best) { best = s; bi = i; }
    }
    pick[k] = bi;
    double a[6];
    const double* ry = pv.rays + 6 * (size_t)sample[bi];
    epipolar_row(ry, ry + 3, a);
    for (int j = 0; j < k; ++j) {
      double d = 0.0;
      for (int r = 0; r < 6; ++r) d += q[j][r] * a[r];
      for (int r = 0; r < 6; ++r) a[r] -= d * q[j][r];
    }
    double s = 0.0;
    for (int r = 0; r < 6; ++r) s += a[r] * a[r];
    s = 1.0 / sqrt(s);
    for (int r = 0; r < 6; ++r) q[k][r] = a[r] * s;
  }
  double models[4][6];
  const double* c0 = pv.rays + 6 * (size_t)sample[pick[0]];
  const double* c1 = pv.rays + 6 * (size_t)sample[pick[1]];
  const double* c2 = pv.rays + 6 * (size_t)sample[pick[2]];
  solve_minimal<0>(c0, c0 + 3, c1, c1 + 3, c2, c2 + 3, models, skip_complex);
  double best_score = INFINITY;
  int best_ind = 0;
  for (int m = 0; m < 4; ++m) {
    double Em[9];
    E_from_p(models[m], Em);
    double score = 0.0;
    for (int j = 0; j < ns; ++j) {
      const double* ry = pv.rays + 6 * (size_t)sample[j];
      score += sampson_exact(Em, ry, ry + 3);
    }
    if (score < best_score) { best_score = score; best_ind = m; }
  }
  E_from_p(models[best_ind], E);
  (void)cx;
  return true;
}

template <class Ctx>
SSFM_HD_NOINLINE void local_optimization(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, double* E_best,
                                         double* score_best, int* cnt_best, long long* evals) {  // ransac.h:341-407
  if (4 > pv.n) return;  // non_minimal_sample_size() == 4
  const double thr = P.thr2, mult = P.thr_mult;
  double m_init[9];
  for (int i = 0; i < 9; ++i) m_init[i] = E_best[i];
  lsq_fit(cx, P, pv, sc, thr * mult, m_init, evals);
  int cnt = 0;
  double score;
  { SSFM_TIC score = msac_score_exact(cx, m_init, pv.stream, pv.n, thr, &cnt, evals); SSFM_TOC(sc, PH_LO_SCORE) }
  keep_better(score, m_init, cnt, score_best, E_best, cnt_best);
  if (P.num_lo_steps <= 0) return;  // inliers_base is only used by the steps below
  const int nbase = collect_inliers(cx, m_init, pv.stream, pv.n, thr * mult, false, sc.list_b, (unsigned char*)0, evals);
  int non_min = 3 * P.non_min_mult;
  if (nbase / 2 < non_min) non_min = nbase / 2;
  if (non_min < 4) non_min = 4;
  for (int r = 0; r < P.num_lo_steps; ++r) {
    // sample = inliers_base; RandomShuffleAndResize(non_min)   (:380-381)
    for (int i = cx.lane(); i < nbase; i += cx.width()) sc.list_a[i] = sc.list_b[i];
    cx.sync();
    shuffle_and_resize(cx, sc.mt, sc.list_a, nbase, non_min);
    // std::vector::resize(target) grows with zeros when target > size (non_min > nbase)
    const int ns = non_min;
    if (ns > nbase) {
      if (cx.lane() == 0)
        for (int i = nbase; i < ns; ++i) sc.list_a[i] = 0;
      cx.sync();
    }
    double m[9];
    if (!non_minimal_solver(cx, pv, sc.list_a, ns, m, P.skip_complex != 0)) continue;
    score = msac_score_exact(cx, m, pv.stream, pv.n, thr, &cnt, evals);
    keep_better(score, m, cnt, score_best, E_best, cnt_best);
    lsq_fit(cx, P, pv, sc, thr, m, evals);
    double th = mult * thr;
    const double dth = (mult - 1.0) * thr / (double)(int)(P.num_lsq_iters - 1);
    for (int i = 0; i < P.num_lsq_iters; ++i) {
      lsq_fit(cx, P, pv, sc, th, m, evals);
      score = msac_score_exact(cx, m, pv.stream, pv.n, thr, &cnt, evals);
      keep_better(score, m, cnt, score_best, E_best, cnt_best);
      th -= dth;
    }
  }
}

// LocalOptimization for num_lo_steps == 0, first half: the inlier collection and shuffle of
// LeastSquaresFit (ransac.h:361 -> :409-418).  Returns the number of residuals of the refit
// (0: fewer than min_sample_size inliers, no refit; < 0: LocalOptimization returns at :350).
template <class Ctx>
SSFM_HD_NOINLINE int lo_begin(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, const double* E,
                              long long* evals) {
  if (4 > pv.n) return -1;
  const int cap = P.min_sample_mult * 3;
  int n;
  { SSFM_TIC n = collect_inliers(cx, E, pv.stream, pv.n, P.thr2 * P.thr_mult, false, sc.list_a, (unsigned char*)0, evals); SSFM_TOC(sc, PH_LO_COLLECT) }
  if (n < 3) return 0;
  { SSFM_TIC shuffle_and_resize(cx, sc.mt, sc.list_a, n, n < cap ? n : cap); SSFM_TOC(sc, PH_LO_SHUFFLE) }
  return n < cap ? n : cap;
}
// second half (ransac.h:363-366): score the refitted model, keep it if it is better.
template <class Ctx>
SSFM_HD_NOINLINE void lo_end(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, const double* m_init,
                             double* E_target, double* score_target, int* cnt_target, long long* evals) {
  int cnt = 0;
  double score;
  { SSFM_TIC score = msac_score_exact(cx, m_init, pv.stream, pv.n, P.thr2, &cnt, evals); SSFM_TOC(sc, PH_LO_SCORE) }
  keep_better(score, m_init, cnt, score_target, E_target, cnt_target);
}

// GetInliers on *best_model + inlier ratio + NumRequiredIterations (ransac.h:231-238).  The count is
// the one recorded when *best_model was scored (same model, same threshold, same arithmetic).
SSFM_HD void refresh(const Params& P, const PairView& pv, PairState& st, bool update_max) {
  st.best_num_inliers = st.cnt_best;
  st.inlier_ratio = (double)st.best_num_inliers / (double)pv.n;
  if (update_max) st.max_iters = required_iterations(st.inlier_ratio, P.eta, 3, P.min_iters, P.max_iters);
}

SSFM_HD void init_state(const Params& P, int n, PairState& st) {
  for (int i = 0; i < 9; ++i) { st.E_best[i] = 0.0; st.E_bestmin[i] = 0.0; }
  st.best_model_score = kDblMax;
  st.best_min_score = P.driver == 2 ? INFINITY : kDblMax;
  st.inlier_ratio = 0.0;
  st.legacy_num_iter = (double)P.fixed_budget;
  st.evals_exact = 0;
  st.it = 0;
  st.max_iters = P.max_iters > P.min_iters ? P.max_iters : P.min_iters;
  if (P.driver == 2) st.max_iters = (uint32_t)(P.fixed_budget > 0 ? P.fixed_budget : 0);
  st.best_num_inliers = 0;
  st.cnt_best = 0;
  st.cnt_bestmin = 0;
  st.num_lo = 0;
  st.done = (n < 3 || n < P.min_points) ? 1 : 0;  // kMinSampleSize > kNumData -> return 0 (ransac.h:137-139)
  st.phase = PHASE_NONE;
  st.walk_j = 0;
  st.lm_n = 0;
  st.runmin32 = INFINITY;
}

// Number of minimal-sample iterations the pair still wants (for the next look-ahead round).
SSFM_HD uint32_t iterations_wanted(const Params& P, const PairState& st) {
  if (st.done) return 0;
  uint32_t lim = st.max_iters;
  if (P.driver == 2) {
    const double ni = st.legacy_num_iter;
    const uint32_t c = ni >= (double)P.fixed_budget ? (uint32_t)P.fixed_budget : (uint32_t)ceil(ni > 0 ? ni : 0);
    lim = c < lim ? c : lim;
  }
  return lim > st.it ? lim - st.it : 0;
}

// Consume one look-ahead round: iterations [st.it, st.it + navail) of this pair were sampled,
// solved (models: navail x 4 x 6 doubles, NaN = absent) and scored in FP32 (s32[j] = the best of the
// four FP32 MSAC costs of iteration j).  An iteration can only matter to the reference loop if
// its float64 score beats the best so far, so only iterations whose FP32 score is within
// cand_margin of the running FP32 minimum are re-scored here in float64 (bit-compatible
// arithmetic); everything the reference does at such an iteration -- best-model bookkeeping,
// local optimisation, inlier refresh, adaptive termination -- then runs exactly.
// With DEFER (only valid for num_lo_steps == 0) the function returns early with st.phase != 0
// whenever a refit is needed; calling it again after the refit was solved resumes at that point.
template <class Ctx, bool DEFER>
SSFM_HD_NOINLINE void process_round(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, PairState& st,
                           const double* models, int mstride, const float* s32, const float* s32m, int navail) {
  // model (slot j, root m, parameter i) lives at models[(m * 6 + i) * mstride + j];
  // s32[j] = min over the four roots of the FP32 cost, s32m[m * mstride + j] = the four costs.
  const bool lo_driver = P.driver == 0;
  const bool legacy = P.driver == 2;
  const float abs_slack = (float)pv.n * (float)P.thr2 * 2.4e-7f;
  int j = 0;
  bool skip_top = false;
  if (DEFER && st.phase == PHASE_LO_START) {  // the refit of the LO at lo_start came back (ransac.h:166-177)
    lo_end(cx, P, pv, sc, sc.lm_E, st.E_best, &st.best_model_score, &st.cnt_best, &st.evals_exact);
    refresh(P, pv, st, true);
    st.phase = PHASE_NONE;
    j = st.walk_j;
    skip_top = true;
  } else if (DEFER && st.phase == PHASE_LO_BEST) {  // the refit of a new best minimal model came back (:223-238)
    double score = st.best_min_score;
    lo_end(cx, P, pv, sc, sc.lm_E, st.E_bestmin, &score, &st.cnt_bestmin, &st.evals_exact);
    keep_better(score, st.E_bestmin, st.cnt_bestmin, &st.best_model_score, st.E_best, &st.cnt_best);
    refresh(P, pv, st, true);
    st.phase = PHASE_NONE;
    j = st.walk_j + 1;
    ++st.it;
  }
  while (!st.done) {
   if (!skip_top) {
    // loop condition of the reference's for/while
    if (legacy) {
      if (!((double)st.it < st.legacy_num_iter && (int)st.it < P.fixed_budget)) { st.done = 1; break; }
    } else {
      if (st.it >= st.max_iters) { st.done = 1; break; }
    }
    if (j >= navail) break;  // need another look-ahead round
    // ---- find the next iteration in this round that can matter
    int jn = navail;  // first candidate / special index >= j
    {
      SSFM_TIC
      float run = st.runmin32;
      int base = j;
      while (base < navail && jn == navail) {
        const int jj = base + cx.slane();
        const float s = jj < navail ? s32[jj] : INFINITY;
        const float before = cx.prefix_min_excl(s, run);
        // relative slack + an absolute one for scores near zero (noise-free data), where FP32 rounding noise of the sum
        // (<= n * thr * 2^-22) is all that separates the iterations
        const bool cand = s < INFINITY && s <= before * (1.0f + P.cand_margin) + abs_slack;
        const bool special = lo_driver && jj < navail && (st.it + (uint32_t)(jj - j)) == P.lo_start;
        const unsigned m = cx.ballot(cand || special);
        if (m) {
#if defined(__CUDA_ARCH__)
          jn = base + (__ffs(m) - 1);
#else
          jn = base + (__builtin_ffs(m) - 1);
#endif
        }
        // the running minimum covers every iteration consumed so far (<= jn)
        run = fminf(run, cx.min_f(jj <= jn ? s : INFINITY));
        base += cx.swidth();
      }
      st.runmin32 = run;
      SSFM_TOC(sc, PH_SCAN)
    }
    // iterations j .. jn-1 cannot matter: skip them, honouring the termination test
    {
      const uint32_t skip = (uint32_t)(jn - j);
      uint32_t lim = st.max_iters;
      if (legacy) {
        const double ni = st.legacy_num_iter;
        const uint32_t c = ni >= (double)P.fixed_budget ? (uint32_t)P.fixed_budget : (uint32_t)ceil(ni > 0 ? ni : 0);
        lim = c;
      }
      if (st.it + skip >= lim) {
        // the loop ends inside the skipped stretch (or exactly at its end)
        if (lim > st.it) st.it = lim;
        st.done = 1;
        break;
      }
      st.it += skip;
      j = jn;
    }
    if (j >= navail) break;
    // ---- iteration st.it (slot j) is a candidate and/or the lo_start iteration
    if (lo_driver && st.it == P.lo_start && st.best_min_score < kDblMax) {  // ransac.h:163-178
      ++st.num_lo;
      if (DEFER) {
        const int nr = lo_begin(cx, P, pv, sc, st.E_best, &st.evals_exact);
        if (nr > 0) {
          if (cx.lane() == 0)
            for (int i = 0; i < 9; ++i) sc.lm_E[i] = st.E_best[i];
          st.lm_n = nr;
          st.phase = PHASE_LO_START;
          st.walk_j = j;
          return;
        }
      } else {
        local_optimization(cx, P, pv, sc, st.E_best, &st.best_model_score, &st.cnt_best, &st.evals_exact);
      }
      refresh(P, pv, st, true);
    }
   }
    skip_top = false;
    // exact re-scoring of the iteration's models (GetBestEstimatedModelId, :278-293)
    // Only roots whose FP32 cost is within the pre-filter slack of the iteration's FP32 minimum can
    // be the float64 argmin; the others are skipped.
    double local_best = legacy ? INFINITY : kDblMax;
    int local_id = -1, local_cnt = 0;
    int nvalid = 0;
    const float smin = s32[j];
    SSFM_TIC
    for (int m = 0; m < 4; ++m) {
      double p[6];
      for (int i = 0; i < 6; ++i) p[i] = models[(size_t)(m * 6 + i) * mstride + j];
      if (p[0] != p[0]) continue;  // absent (or NaN) model: can never win a '<'
      ++nvalid;
      const float sm = s32m[(size_t)m * mstride + j];
      if (!(sm <= smin * (1.0f + P.cand_margin) + abs_slack)) continue;
      double Em[9];
      E_from_p(p, Em);
      int cnt = 0;
      const double s = msac_score_exact(cx, Em, pv.stream, pv.n, P.thr2, &cnt, &st.evals_exact);
      if (legacy) {
        if (s < st.best_min_score && s < local_best) { local_best = s; local_id = m; local_cnt = cnt; }
      } else if (s < local_best) {
        local_best = s;
        local_id = m;
        local_cnt = cnt;
      }
    }
    SSFM_TOC(sc, PH_RESCORE)
    // MinimalSolver returned <= 0 models -> `continue` (ransac.h:185).  The action-matrix and
    // polynomial solvers always return 4 (possibly NaN) matrices; the Sturm variant returns the
    // number of real roots it kept.
    const bool has_models = P.solver != 2 || nvalid > 0;
    if (legacy) {
      if (local_id >= 0) {  // msac.h:103-110 (running best over all models in order)
        double p[6], Em[9];
        for (int i = 0; i < 6; ++i) p[i] = models[(size_t)(local_id * 6 + i) * mstride + j];
        E_from_p(p, Em);
        st.best_min_score = local_best;
        st.best_model_score = local_best;
        for (int i = 0; i < 9; ++i) st.E_best[i] = Em[i];
        st.best_num_inliers = collect_inliers(cx, Em, pv.stream, pv.n, P.thr2, true, (int*)0, (unsigned char*)0, &st.evals_exact);
        st.inlier_ratio = (double)st.best_num_inliers / (double)pv.n;
        const double outlier_ratio = (pv.n - st.best_num_inliers) / (double)pv.n;
        if (outlier_ratio < 1.0) {  // msac.h:119-126
          double ni = log(1. - P.fixed_prob) / log(1. - pow(1. - outlier_ratio, 3.0));
          if (ni > P.fixed_budget) ni = P.fixed_budget;
          st.legacy_num_iter = ni;
        }
      }
    } else if (has_models) {
      const bool is_best = local_id >= 0 && local_best < st.best_min_score;
      const bool at_lo_start = lo_driver && st.it == P.lo_start;
      if (is_best || at_lo_start) {  // ransac.h:195-239
        if (is_best) {
          double p[6], Em[9];
          for (int i = 0; i < 6; ++i) p[i] = models[(size_t)(local_id * 6 + i) * mstride + j];
          E_from_p(p, Em);
          st.best_min_score = local_best;
          st.cnt_bestmin = local_cnt;
          for (int i = 0; i < 9; ++i) st.E_bestmin[i] = Em[i];
          keep_better(st.best_min_score, st.E_bestmin, st.cnt_bestmin, &st.best_model_score, st.E_best, &st.cnt_best);
        }
        const bool run_lo = lo_driver && st.it >= P.lo_start && st.best_min_score < kDblMax;
        if (is_best || run_lo) {
          if (run_lo) {
            ++st.num_lo;
            if (DEFER) {
              const int nr = lo_begin(cx, P, pv, sc, st.E_bestmin, &st.evals_exact);
              if (nr > 0) {
                if (cx.lane() == 0)
                  for (int i = 0; i < 9; ++i) sc.lm_E[i] = st.E_bestmin[i];
                st.lm_n = nr;
                st.phase = PHASE_LO_BEST;
                st.walk_j = j;
                return;
              }
            } else {
              double score = st.best_min_score;
              local_optimization(cx, P, pv, sc, st.E_bestmin, &score, &st.cnt_bestmin, &st.evals_exact);
              keep_better(score, st.E_bestmin, st.cnt_bestmin, &st.best_model_score, st.E_best, &st.cnt_best);
            }
          }
          refresh(P, pv, st, true);
        }
      }
    }
    ++st.it;
    ++j;
  }
}

// After the loop: the late LO (:245-255), the final least squares (:257-272), and what the
// callers do next: the inlier mask (examples/spherical_sfm_tools.cpp:388-392) and the pose
// (:414-418).  Writes r, t; returns the per-pair status.
// Returns -1 (with st.phase != 0) when DEFER parked a refit.
template <class Ctx, bool DEFER>
SSFM_HD_NOINLINE int finalize_pair(const Ctx& cx, const Params& P, const PairView& pv, const Scratch& sc, PairState& st, double* r,
                          double* t, unsigned char* flags) {
  r[0] = r[1] = r[2] = 0.0;
  t[0] = t[1] = t[2] = 0.0;
  if (pv.n < 3) {
    if (flags)
      for (int i = cx.lane(); i < pv.n; i += cx.width()) flags[i] = 0;
    return 1;
  }
  if (P.driver == 0) {
    int stage = 0;  // 0 from the top, 1 late LO done, 2 final refit done
    if (DEFER && st.phase == PHASE_LO_LATE) {
      lo_end(cx, P, pv, sc, sc.lm_E, st.E_best, &st.best_model_score, &st.cnt_best, &st.evals_exact);
      refresh(P, pv, st, false);
      st.phase = PHASE_NONE;
      stage = 1;
    } else if (DEFER && st.phase == PHASE_FINAL_LSQ) {
      st.phase = PHASE_NONE;
      stage = 2;
    }
    if (stage == 0 && st.it <= P.lo_start && st.best_model_score < kDblMax) {  // ransac.h:245-255
      ++st.num_lo;
      if (DEFER) {
        const int nr = lo_begin(cx, P, pv, sc, st.E_best, &st.evals_exact);
        if (nr > 0) {
          if (cx.lane() == 0)
            for (int i = 0; i < 9; ++i) sc.lm_E[i] = st.E_best[i];
          st.lm_n = nr;
          st.phase = PHASE_LO_LATE;
          return -1;
        }
      } else {
        local_optimization(cx, P, pv, sc, st.E_best, &st.best_model_score, &st.cnt_best, &st.evals_exact);
      }
      refresh(P, pv, st, false);
    }
    if (P.final_lsq) {  // ransac.h:257-272
      // LeastSquares on ALL current inliers of best_model (stats.inlier_indices)
      double refined[9];
      if (stage < 2) {
        const int ni = collect_inliers(cx, st.E_best, pv.stream, pv.n, P.thr2, false, sc.list_a, (unsigned char*)0, &st.evals_exact);
        if (DEFER) {
          if (cx.lane() == 0)
            for (int i = 0; i < 9; ++i) sc.lm_E[i] = st.E_best[i];
          st.lm_n = ni;
          st.phase = PHASE_FINAL_LSQ;
          return -1;
        }
        for (int i = 0; i < 9; ++i) refined[i] = st.E_best[i];
        SSFM_TIC
        if (P.inline_small_max > 0 && ni <= P.inline_small_max)  // the deferred path solves a refit this small on one lane
          least_squares_as_deferred(cx, pv.rays, sc.list_a, ni, P.inward != 0, refined, P.inline_handover);
        else
          least_squares(cx, pv.rays, sc.list_a, ni, P.inward != 0, refined);
        SSFM_TOC(sc, PH_FINAL_LM)
      } else {
        for (int i = 0; i < 9; ++i) refined[i] = sc.lm_E[i];
      }
      int cnt = 0;
      const double score = msac_score_exact(cx, refined, pv.stream, pv.n, P.thr2, &cnt, &st.evals_exact);
      if (score < st.best_model_score) {
        st.best_model_score = score;
        st.cnt_best = cnt;
        for (int i = 0; i < 9; ++i) st.E_best[i] = refined[i];
        refresh(P, pv, st, false);
      }
    }
  }
  const bool have = P.driver == 2 ? (st.best_min_score < INFINITY) : (st.best_model_score < kDblMax);
  if (!have) {
    if (P.driver == 2) st.best_model_score = INFINITY;
    if (flags)
      for (int i = cx.lane(); i < pv.n; i += cx.width()) flags[i] = 0;
    return 2;
  }
  if (flags) collect_inliers(cx, st.E_best, pv.stream, pv.n, P.thr2, P.driver == 2, (int*)0, flags, &st.evals_exact);
  decompose_spherical_E(st.E_best, P.inward != 0, r, t);
  return 0;
}

}  // namespace ssfm
