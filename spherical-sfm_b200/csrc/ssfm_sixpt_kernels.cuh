// ssfm_sixpt_kernels.cuh -- the six-point shared-focal estimator (SixPointEstimator,
// examples/six_point_estimator.{h,cpp}) under VanillaMSAC (evaluation/vanilla_ransac.h:23-99) for a batch of
// pairs (config C4).  Same look-ahead-round structure as the 3-point path:
//
//   k_sixpt_init          per-pair state
//   k_sixpt_sample_solve  one thread per (pair, look-ahead iteration): Philox sample of 6, solver (<= 15 models);
//                         the models of a pair are appended to a PACKED list (on average only ~2 of the 15
//                         possible solutions are real with positive focal and pass the cheirality test, so
//                         scoring a dense [iteration][15] grid would leave most lanes idle)
//   k_sixpt_score         FP32 MSAC cost of every packed model against every correspondence of its pair; one
//                         thread owns one model (general 3x3 matrix in registers), correspondences streamed
//                         through shared memory and read as broadcast LDS.128
//   k_sixpt_chain         one warp per pair: FP64 certification of the iterations that can matter to the
//                         reference loop + the loop's bookkeeping (best model, inliers, NumRequiredIterations)
#pragma once
#include "ssfm_kernels.cuh"
#include "ssfm_sixpt.cuh"
#include "ssfm_sixpt_coop.cuh"
#include "ssfm_sixpt_lo.cuh"

namespace ssfm {

struct SixState {
  double G[9];  // scoring matrix of the best model
  SixPointModel best;
  double best_score;  // best_min_model_score == stats.best_model_score (vanilla_ransac.h:68-79)
  double inlier_ratio;
  long long evals;
  uint32_t it, max_iters;
  int best_num_inliers;
  int done;
  float runmin32;
};

constexpr int kSixSlotModels = 16;                 // per-iteration stride of the FP32 score table
constexpr int kSixRecord = 16;                     // doubles per stored model: G[9], t[3], r[3], f
constexpr float kSixCandMargin = 5e-3f;            // FP32 pre-filter slack (pixel-unit rays: larger dynamic range)

__global__ void k_sixpt_init(Params P, const long long* __restrict__ offsets, int pair0, int npairs, SixState* states,
                             int* active, int* navail, int first_cap, int* count, uint32_t* round_it) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= npairs) return;
  round_it[a] = 0;
  const int pair = pair0 + a;
  const int n = (int)(offsets[pair + 1] - offsets[pair]);
  SixState st;
  for (int i = 0; i < 9; ++i) st.G[i] = 0.0;
  for (int i = 0; i < 3; ++i) { st.best.t[i] = 0.0; st.best.r[i] = 0.0; }
  st.best.f = 0.0;
  st.best_score = kDblMax;
  st.inlier_ratio = 0.0;
  st.evals = 0;
  st.it = 0;
  st.max_iters = P.max_iters > P.min_iters ? P.max_iters : P.min_iters;  // vanilla_ransac.h:42-43
  st.best_num_inliers = 0;
  st.done = (n < 6 || n < P.min_points || st.max_iters == 0) ? 1 : 0;      // :33-37
  st.runmin32 = INFINITY;
  states[a] = st;
  navail[a] = st.done ? 0 : lookahead(st.max_iters, first_cap);
  active[a] = a;
  if (a == 0) *count = npairs;
}

// grid (active pairs, ceil(cap / kSixSamplesPerBlock)), kSixSolveThreads threads, kSixSolveSmem bytes of dynamic shared
// memory.  Eight lanes per sample, four samples per warp (ssfm_sixpt_coop.cuh): every sample's 16 x 16 companion / Hessenberg
// matrix and its small tables live in shared memory (3.6 KB), the ten cubics in a global scratch slot (2.4 KB, read back
// through L2), so nothing of the solver goes through local memory except the per-candidate least-squares tableau.
#ifndef SSFM_SIXPT_THREADS
#define SSFM_SIXPT_THREADS 256  // 8 warps = 32 samples per block (divides the 256 look-ahead slots), two blocks per SM
#endif
constexpr int kSixSolveThreads = SSFM_SIXPT_THREADS;
constexpr int kSixSamplesPerBlock = (kSixSolveThreads / 32) * sixc::kSixSamplesPerWarp;
constexpr size_t kSixSolveSmem = (size_t)kSixSamplesPerBlock * sixc::kScratch * sizeof(double);
#ifndef SSFM_SIXPT_MINBLOCKS
#define SSFM_SIXPT_MINBLOCKS 2  // x 256 threads: registers capped at 128
#endif
__global__ void __launch_bounds__(kSixSolveThreads, SSFM_SIXPT_MINBLOCKS)
    k_sixpt_sample_solve(Params P, const double* __restrict__ rays, const long long* __restrict__ offsets, int pair0,
                         const int* __restrict__ active, const int* __restrict__ navail, const uint32_t* __restrict__ round_it,
                         int R, double* __restrict__ models, int* __restrict__ nmodels, float* __restrict__ pk_G,
                         int* __restrict__ pk_id, int* __restrict__ pk_count, float* __restrict__ s32m, double* __restrict__ scratch_M) {
  extern __shared__ double six_smem[];
  const int a = active[blockIdx.x];
  const int na = navail[a];
  const int j0 = blockIdx.y * kSixSamplesPerBlock;
  if (j0 >= na) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const int j = j0 + warp * sixc::kSixSamplesPerWarp + grp;  // this group's look-ahead slot
  const bool valid = j < na;
  double* S = six_smem + (size_t)(warp * sixc::kSixSamplesPerWarp + grp) * sixc::kScratch;
  double* Mg = scratch_M + ((size_t)a * R + j) * sixc::kMSize;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  double c[6][6];
  if (valid && gl == 0) {
    const int n = (int)(offsets[pair + 1] - off);
    const uint32_t it = round_it[a] + (uint32_t)j;  // iteration number of look-ahead slot j
    int idx[6];
    philox_sample<6>(P.seed, P.first_pair_id + (uint32_t)pair, it, 6, n, idx);
    for (int s = 0; s < 6; ++s) {
      const double2* src = reinterpret_cast<const double2*>(rays + 6 * (off + idx[s]));
      const double2 x0 = src[0], x1 = src[1], x2 = src[2];
      c[s][0] = x0.x; c[s][1] = x0.y; c[s][2] = x1.x; c[s][3] = x1.y; c[s][4] = x2.x; c[s][5] = x2.y;
    }
  }
  const int nm = sixc::six_solve_group<true>(S, Mg, c, valid);
  if (!valid) return;
  float* srow = s32m + ((size_t)a * R + j) * kSixSlotModels;
  srow[gl] = INFINITY;
  srow[gl + 8] = INFINITY;
  int base = 0;
  if (gl == 0) {
    nmodels[(size_t)a * R + j] = nm;
    if (nm > 0) base = atomicAdd(&pk_count[a], nm);
  }
  if (nm == 0) return;
  base = __shfl_sync(0xFFu << (8 * grp), base, 8 * grp);
  const SixPointModel* list = reinterpret_cast<const SixPointModel*>(S + sixc::kOffList);
  double* dst = models + ((size_t)a * R + j) * kSixMaxModels * kSixRecord;
  for (int k = gl; k < nm; k += 8) {
    const SixPointModel mdl = list[k];
    double G[9];
    sixpt_scoring_matrix(mdl, P.sixpt_focal_scoring, G);
    double* d = dst + (size_t)k * kSixRecord;
    for (int q = 0; q < 9; ++q) d[q] = G[q];
    for (int q = 0; q < 3; ++q) { d[9 + q] = mdl.t[q]; d[12 + q] = mdl.r[q]; }
    d[15] = mdl.f;
    // FP32 copy, normalised (the Sampson error is invariant to the scale of G; keeps the floats in range)
    double nrm = 0.0;
    for (int q = 0; q < 9; ++q) nrm += G[q] * G[q];
    const double inv = nrm > 0.0 ? 1.0 / sqrt(nrm) : 0.0;
    float* gq = pk_G + ((size_t)a * R * kSixMaxModels + base + k) * 9;
    for (int q = 0; q < 9; ++q) gq[q] = (float)(G[q] * inv);
    pk_id[(size_t)a * R * kSixMaxModels + base + k] = j * kSixSlotModels + k;
  }
}

// grid (active pairs, ceil(R * 15 / 128)), 128 threads: thread = one packed model of the pair
template <bool UNITZ>
__global__ void __launch_bounds__(128)
    k_sixpt_score(const float4* __restrict__ pa, const float4* __restrict__ pb, const long long* __restrict__ offsets,
                  int pair0, const int* __restrict__ active, int R, const float* __restrict__ pk_G,
                  const int* __restrict__ pk_id, const int* __restrict__ pk_count, float thr, float* __restrict__ s32m,
                  unsigned long long* __restrict__ counters) {
  constexpr int TILE = 512;
  __shared__ float4 sa[TILE];
  __shared__ float4 sb[UNITZ ? 1 : TILE];
  const int a = active[blockIdx.x];
  const int count = pk_count[a];
  const int first = blockIdx.y * blockDim.x;
  if (first >= count) return;
  const int e = first + threadIdx.x;
  const bool live = e < count;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  float g[9];
  {
    const float* src = pk_G + ((size_t)a * R * kSixMaxModels + (live ? e : first)) * 9;
#pragma unroll
    for (int q = 0; q < 9; ++q) g[q] = src[q];
  }
  float total = 0.f;
  for (int t0 = 0; t0 < n; t0 += TILE) {
    const int len = n - t0 < TILE ? n - t0 : TILE;
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      sa[i] = pa[off + t0 + i];
      if (!UNITZ) sb[i] = pb[off + t0 + i];
    }
    __syncthreads();
    float part = 0.f;
#pragma unroll 4
    for (int i = 0; i < len; ++i) {
      float ux, uy, uz, vx, vy, vz;
      const float4 A4 = sa[i];
      if (UNITZ) {
        ux = A4.x; uy = A4.y; uz = 1.f; vx = A4.z; vy = A4.w; vz = 1.f;
      } else {
        const float4 B4 = sb[i];
        ux = A4.x; uy = A4.y; uz = A4.z; vx = B4.x; vy = B4.y; vz = B4.z;
      }
      const float Eu0 = fmaf(g[2], uz, fmaf(g[1], uy, g[0] * ux));
      const float Eu1 = fmaf(g[5], uz, fmaf(g[4], uy, g[3] * ux));
      const float Eu2 = fmaf(g[8], uz, fmaf(g[7], uy, g[6] * ux));
      const float Et0 = fmaf(g[6], vz, fmaf(g[3], vy, g[0] * vx));
      const float Et1 = fmaf(g[7], vz, fmaf(g[4], vy, g[1] * vx));
      const float d = fmaf(vz, Eu2, fmaf(vy, Eu1, vx * Eu0));
      const float den = fmaf(Et1, Et1, fmaf(Et0, Et0, fmaf(Eu1, Eu1, Eu0 * Eu0)));
      const float err = __fdividef(d * d, den);
      part += fminf(err, thr);  // a NaN error counts as thr; the FP64 certification decides
    }
    total += part;
  }
  if (live) s32m[(size_t)a * R * kSixSlotModels + pk_id[(size_t)a * R * kSixMaxModels + e]] = total;
  if (threadIdx.x == 0) {
    const int lanes = count - first < (int)blockDim.x ? count - first : (int)blockDim.x;
    atomicAdd(&counters[0], (unsigned long long)lanes * (unsigned long long)n);
  }
}

struct SixChainArgs {
  const double* rays;
  const long long* offsets;
  int pair0;
  const int* list;
  int nlist;
  int* navail;
  SixState* states;
  int R;
  const double* models;
  const int* nmodels;
  const float* s32m;
  unsigned char* flags;
  SsfmPairResult* results;
  int* next_active;
  int* next_count;
  int next_cap;
  int* pk_count;
  unsigned long long* counters;
  uint32_t* round_it;  // iteration number of look-ahead slot 0 of the pair's next round
  // LO-MSAC variant only
  SixLoState* lo_states;
  int* parked;        // pairs waiting for k_sixpt_lo
  int* parked_count;
  int* list_a;        // CSR-aligned scratch lists (list_base = first correspondence of the pass)
  int* list_b;
  long long list_base;
  uint32_t* mt;       // 625 words per pair
};

constexpr int kSixChainWarps = 4;

__global__ void __launch_bounds__(kSixChainWarps * 32) k_sixpt_chain(Params P, SixChainArgs A) {
  const int wid = blockIdx.x * kSixChainWarps + (threadIdx.x >> 5);
  if (wid >= A.nlist) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = A.list[wid];
  const int pair = A.pair0 + a;
  const long long off = A.offsets[pair];
  const int n = (int)(A.offsets[pair + 1] - off);
  PairView pv{A.rays + 6 * off, n, A.rays + 6 * off};
  SixState st = A.states[a];
  const int na = A.navail[a];
  const double* mbase = A.models + (size_t)a * A.R * kSixMaxModels * kSixRecord;
  const float* sbase = A.s32m + (size_t)a * A.R * kSixSlotModels;
  long long exact = 0;
  for (int j = 0; j < na && !st.done; ++j) {
    if (st.it >= st.max_iters) { st.done = 1; break; }
    const int nm = A.nmodels[(size_t)a * A.R + j];
    st.it += 1;
    if (nm <= 0) continue;  // vanilla_ransac.h:57
    st.evals += (long long)nm * n;
    // FP32 minimum of this iteration (lanes 0..15 hold one model each)
    const float mine = cx.lane() < nm ? sbase[j * kSixSlotModels + cx.lane()] : INFINITY;
    const float smin = cx.min_f(mine);
    // An iteration matters only if its float64 score beats the best so far; its FP32 score is then within
    // the slack of the running FP32 minimum.  (A NaN FP32 score is always certified.)
    const bool nan_any = __any_sync(0xffffffffu, mine != mine);
    if (!nan_any && !(smin <= st.runmin32 * (1.0f + kSixCandMargin))) continue;
    if (smin < st.runmin32) st.runmin32 = smin;
    double local_best = kDblMax;
    int local_id = 0, local_cnt = 0;
    for (int k = 0; k < nm; ++k) {
      const float sk = __shfl_sync(0xffffffffu, mine, k);
      if (!(sk != sk) && !(sk <= smin * (1.0f + kSixCandMargin))) continue;
      const double* rec = mbase + ((size_t)j * kSixMaxModels + k) * kSixRecord;
      double G[9];
      for (int q = 0; q < 9; ++q) G[q] = rec[q];
      int cnt = 0;
      const double s = msac_score_exact(cx, G, pv.rays, pv.n, P.thr2, &cnt, &exact);
      if (s < local_best) { local_best = s; local_id = k; local_cnt = cnt; }  // GetBestEstimatedModelId, ransac.h:278-293
    }
    if (local_best < st.best_score) {  // vanilla_ransac.h:68-92
      st.best_score = local_best;
      const double* rec = mbase + ((size_t)j * kSixMaxModels + local_id) * kSixRecord;
      for (int q = 0; q < 9; ++q) st.G[q] = rec[q];
      for (int q = 0; q < 3; ++q) { st.best.t[q] = rec[9 + q]; st.best.r[q] = rec[12 + q]; }
      st.best.f = rec[15];
      st.best_num_inliers = local_cnt;  // GetInliers of the same model at the same threshold (:83-84)
      st.inlier_ratio = (double)local_cnt / (double)n;
      st.max_iters = required_iterations(st.inlier_ratio, P.eta, 6, P.min_iters, P.max_iters);
    }
  }
  if (!st.done && st.it >= st.max_iters) st.done = 1;
  if (st.done) {
    const bool have = st.best_score < kDblMax;
    unsigned char* fl = A.flags ? A.flags + off : (unsigned char*)0;
    if (have) {
      if (fl) collect_inliers(cx, st.G, pv.rays, pv.n, P.thr2, false, (int*)0, fl, &exact);
    } else if (fl) {
      for (int i = cx.lane(); i < n; i += 32) fl[i] = 0;
    }
    if (cx.lane() == 0) {
      SsfmPairResult& o = A.results[a];
      for (int i = 0; i < 9; ++i) o.E[i] = st.G[i];
      for (int i = 0; i < 3; ++i) { o.r[i] = st.best.r[i]; o.t[i] = st.best.t[i]; }
      o.best_model_score = st.best_score;
      o.inlier_ratio = st.inlier_ratio;
      o.num_iterations = st.it;
      o.best_num_inliers = st.best_num_inliers;
      o.number_lo_iterations = 0;
      o.status = n < 6 ? SSFM_PAIR_TOO_FEW_POINTS : (n < P.min_points ? SSFM_PAIR_SKIPPED : (have ? SSFM_PAIR_OK : SSFM_PAIR_NO_MODEL));
      o.evals = st.evals;
      o.focal = st.best.f;
    }
  }
  if (cx.lane() == 0) {
    A.states[a] = st;
    A.pk_count[a] = 0;
    A.round_it[a] = st.it;
    if (!st.done) {
      A.navail[a] = lookahead(st.max_iters - st.it, A.next_cap);
      const int pos = atomicAdd(A.next_count, 1);
      A.next_active[pos] = a;
    } else {
      A.navail[a] = 0;
    }
    atomicAdd(&A.counters[1], (unsigned long long)exact);
  }
}

// ---- LO-MSAC around the six-point estimator (ssfm_sixpt_lo.cuh) ----------------------------------------------

__global__ void k_sixpt_lo_init(Params P, const long long* __restrict__ offsets, int pair0, int npairs, SixLoState* states,
                                uint32_t* mt, int* active, int* navail, int first_cap, int* count, uint32_t* round_it) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= npairs) return;
  const int pair = pair0 + a;
  const int n = (int)(offsets[pair + 1] - offsets[pair]);
  SixLoState st;
  six_lo_init(P, n, st);
  if (st.max_iters == 0) st.phase = st.done ? SIX_PH_NONE : SIX_PH_FINAL;  // empty loop: straight to :243
  states[a] = st;
  round_it[a] = 0;
  mt19937_seed(mt + (size_t)a * 625, P.seed);  // rng.seed(options.random_seed_), ransac.h:143-144
  navail[a] = st.done ? 0 : lookahead(st.max_iters, first_cap);
  active[a] = a;
  if (a == 0) *count = npairs;
}

// One warp per listed pair: walks the round's look-ahead slots from st.resume_j exactly like EstimateModel's loop
// (ransac.h:160-241) until the round is used up, the loop ends, or a LocalOptimization is due -- then the pair is
// appended to `parked` for k_sixpt_lo and the walk resumes (same slot or the next one) in the following wave.
__global__ void __launch_bounds__(kSixChainWarps * 32) k_sixpt_chain_lo(Params P, SixChainArgs A) {
  const int wid = blockIdx.x * kSixChainWarps + (threadIdx.x >> 5);
  if (wid >= A.nlist) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = A.list[wid];
  SixLoState st = A.lo_states[a];
  if (st.done) return;  // finished by k_sixpt_lo in the previous wave
  const int pair = A.pair0 + a;
  const long long off = A.offsets[pair];
  const int n = (int)(A.offsets[pair + 1] - off);
  PairView pv{A.rays + 6 * off, n, A.rays + 6 * off};
  const int na = A.navail[a];
  const double* mbase = A.models + (size_t)a * A.R * kSixMaxModels * kSixRecord;
  const float* sbase = A.s32m + (size_t)a * A.R * kSixSlotModels;
  long long exact = 0;
  bool park = st.phase == SIX_PH_FINAL;  // (max_num_iterations == 0)
  int j = st.resume_j;
  for (; !park && j < na; ++j) {
    if (st.phase == SIX_PH_RESUME_BODY) {
      st.phase = SIX_PH_NONE;  // back from the LO at lo_starting_iterations_: the iteration's body continues
    } else {
      if (st.it >= st.max_iters) break;
      if (st.it == P.lo_start && st.best_min_score < kDblMax && !st.lo_start_done) {  // :166-177
        st.phase = SIX_PH_LO_START;
        st.resume_j = j;
        park = true;
        break;
      }
    }
    const int nm = A.nmodels[(size_t)a * A.R + j];
    if (nm <= 0) { st.it += 1; continue; }  // :184
    st.evals += (long long)nm * n;
    const bool force = st.it == P.lo_start;  // :194, the LO branch is entered whatever the score
    const float mine = cx.lane() < nm ? sbase[j * kSixSlotModels + cx.lane()] : INFINITY;
    const float smin = cx.min_f(mine);
    const bool nan_any = __any_sync(0xffffffffu, mine != mine);
    const bool cand = nan_any || smin <= st.runmin32 * (1.0f + kSixCandMargin);
    if (!cand && !force) { st.it += 1; continue; }
    double local_best = kDblMax;
    int local_id = 0, local_cnt = 0;
    if (cand) {
      if (smin < st.runmin32) st.runmin32 = smin;
      for (int k = 0; k < nm; ++k) {
        const float sk = __shfl_sync(0xffffffffu, mine, k);
        if (!(sk != sk) && !(sk <= smin * (1.0f + kSixCandMargin))) continue;
        const double* rec = mbase + ((size_t)j * kSixMaxModels + k) * kSixRecord;
        double G[9];
        for (int q = 0; q < 9; ++q) G[q] = rec[q];
        int cnt = 0;
        const double s = msac_score_exact(cx, G, pv.rays, pv.n, P.thr2, &cnt, &exact);
        if (s < local_best) { local_best = s; local_id = k; local_cnt = cnt; }  // GetBestEstimatedModelId, :278-293
      }
    }
    const bool better = local_best < st.best_min_score;  // kBestMinModel, :196
    if (better) {
      st.best_min_score = local_best;
      const double* rec = mbase + ((size_t)j * kSixMaxModels + local_id) * kSixRecord;
      for (int q = 0; q < 9; ++q) st.bestmin.G[q] = rec[q];
      for (int q = 0; q < 3; ++q) { st.bestmin.m.t[q] = rec[9 + q]; st.bestmin.m.r[q] = rec[12 + q]; }
      st.bestmin.m.f = rec[15];
      st.bestmin.score = local_best;
      st.bestmin.cnt = local_cnt;
      six_keep_better(local_best, local_cnt, st.bestmin.m, st.bestmin.G, st.best);  // :204-206
    }
    const bool run_lo = st.it >= P.lo_start && st.best_min_score < kDblMax;  // :209-211
    st.it += 1;
    if (!better && !force) continue;   // :193-194
    if (!better && !run_lo) continue;  // :213
    if (run_lo) {  // :219-227 and the refresh after it happen in k_sixpt_lo
      st.phase = SIX_PH_LO_BEST;
      st.resume_j = j + 1;
      park = true;
      break;
    }
    six_refresh(P, n, st, true);  // :231-238
  }
  if (!park && st.it >= st.max_iters) {  // the loop has ended: :243-275 run in k_sixpt_lo
    st.phase = SIX_PH_FINAL;
    park = true;
  }
  if (cx.lane() == 0) {
    if (park) {
      const int pos = atomicAdd(A.parked_count, 1);
      A.parked[pos] = a;
    } else {  // look-ahead used up: next round
      st.resume_j = 0;
      A.pk_count[a] = 0;
      A.round_it[a] = st.it;
      A.navail[a] = lookahead(st.max_iters - st.it, A.next_cap);
      const int pos = atomicAdd(A.next_count, 1);
      A.next_active[pos] = a;
    }
    A.lo_states[a] = st;
    atomicAdd(&A.counters[1], (unsigned long long)exact);
  }
}

// One WARP per parked pair: the LocalOptimization (or the loop's epilogue) the pair is waiting for, start to finish.
constexpr int kSixLoWarps = 2;
__global__ void __launch_bounds__(kSixLoWarps * 32) k_sixpt_lo(Params P, SixChainArgs A, int nparked) {
  const int q = blockIdx.x * kSixLoWarps + (threadIdx.x >> 5);
  if (q >= nparked) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = A.parked[q];
  const int pair = A.pair0 + a;
  const long long off = A.offsets[pair];
  const int n = (int)(A.offsets[pair + 1] - off);
  PairView pv{A.rays + 6 * off, n, A.rays + 6 * off};
  SixScratch sc{A.list_a + (off - A.list_base), A.list_b + (off - A.list_base), A.mt + (size_t)a * 625};
  SixLoState st = A.lo_states[a];
  long long exact = 0;
  const bool finished = six_lo_phase(cx, P, pv, sc, st, A.flags ? A.flags + off : (unsigned char*)0, &exact);
  if (cx.lane() != 0) return;
  if (finished) {
    const bool have = st.best.score < kDblMax;
    SsfmPairResult& o = A.results[a];
    for (int i = 0; i < 9; ++i) o.E[i] = st.best.G[i];
    for (int i = 0; i < 3; ++i) { o.r[i] = st.best.m.r[i]; o.t[i] = st.best.m.t[i]; }
    o.best_model_score = st.best.score;
    o.inlier_ratio = st.inlier_ratio;
    o.num_iterations = st.it;
    o.best_num_inliers = st.best_num_inliers;
    o.number_lo_iterations = st.num_lo;
    o.status = have ? SSFM_PAIR_OK : SSFM_PAIR_NO_MODEL;
    o.evals = st.evals;
    o.focal = st.best.m.f;
    A.navail[a] = 0;
  }
  A.lo_states[a] = st;
  atomicAdd(&A.counters[1], (unsigned long long)exact);
}

// Pairs that never enter the loop (fewer than 6 correspondences, or below the caller's min_num_points).
__global__ void k_sixpt_lo_trivial(Params P, const long long* __restrict__ offsets, int pair0, int npairs, const SixLoState* states,
                                   SsfmPairResult* results, unsigned char* flags) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= npairs) return;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  if (!(n < 6 || n < P.min_points)) return;
  SsfmPairResult& o = results[a];
  for (int i = 0; i < 9; ++i) o.E[i] = 0.0;
  for (int i = 0; i < 3; ++i) { o.r[i] = 0.0; o.t[i] = 0.0; }
  o.best_model_score = kDblMax;
  o.inlier_ratio = 0.0;
  o.num_iterations = 0;
  o.best_num_inliers = 0;
  o.number_lo_iterations = 0;
  o.status = n < 6 ? SSFM_PAIR_TOO_FEW_POINTS : SSFM_PAIR_SKIPPED;
  o.evals = 0;
  o.focal = 0.0;
  if (flags)
    for (int i = 0; i < n; ++i) flags[off + i] = 0;
  (void)states;
}

// Hook: the minimal solver on explicit samples (6 indices each) -- the same warp code as the batched kernel.
// grid ceil(ns / kSixSamplesPerBlock), kSixSolveThreads threads, kSixSolveSmem bytes.
__global__ void __launch_bounds__(kSixSolveThreads)
    k_sixpt_solve_samples(const double* __restrict__ rays, const int* __restrict__ samples, int ns,
                          double* __restrict__ models /* ns x 15 x 7 */, int* __restrict__ nmodels, double* __restrict__ scratch_M) {
  extern __shared__ double six_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const int s0 = blockIdx.x * kSixSamplesPerBlock + warp * sixc::kSixSamplesPerWarp;
  const int s = s0 + grp;
  const bool valid = s < ns;
  double* S = six_smem + (size_t)(warp * sixc::kSixSamplesPerWarp + grp) * sixc::kScratch;
  double c[6][6];
  if (valid && gl == 0)
    for (int i = 0; i < 6; ++i)
      for (int q = 0; q < 6; ++q) c[i][q] = rays[6 * (size_t)samples[6 * s + i] + q];
  const int nm = sixc::six_solve_group<true>(S, scratch_M + (size_t)s * sixc::kMSize, c, valid);
  if (!valid) return;
  if (gl == 0) nmodels[s] = nm;
  const SixPointModel* list = reinterpret_cast<const SixPointModel*>(S + sixc::kOffList);
  for (int k = gl; k < nm; k += 8) {
    double* d = models + ((size_t)s * kSixMaxModels + k) * 7;
    for (int q = 0; q < 3; ++q) { d[q] = list[k].t[q]; d[3 + q] = list[k].r[q]; }
    d[6] = list[k].f;
  }
}

// Hook: SixPointEstimator::LeastSquares, one thread per refit.  models: nprob x 7 (t, r, focal), in place.
__global__ void k_sixpt_least_squares(const double* __restrict__ rays, const int* __restrict__ idx,
                                      const int* __restrict__ sample_offsets, int nprob, double* __restrict__ models) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  SixPointModel m;
  double* d = models + 7 * (size_t)p;
  for (int q = 0; q < 3; ++q) { m.t[q] = d[q]; m.r[q] = d[3 + q]; }
  m.f = d[6];
  sixpt_least_squares(rays, idx + sample_offsets[p], sample_offsets[p + 1] - sample_offsets[p], m);
  for (int q = 0; q < 3; ++q) { d[q] = m.t[q]; d[3 + q] = m.r[q]; }
  d[6] = m.f;
}

}  // namespace ssfm
