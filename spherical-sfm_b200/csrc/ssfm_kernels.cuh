// ssfm_kernels.cuh -- sm_100a kernels of the batched relative-pose engine.
//
//   k_pack           AoS float64 RayPair memory -> two float4 SoA planes (u.xyz,0) / (v.xyz,0)
//   k_sample_solve   one thread per (pair, look-ahead iteration): Philox sample + 3-point solver (FP64)
//   k_score_rounds   FP32 Sampson/MSAC scoring of every look-ahead hypothesis against every
//                    correspondence of its pair; ray tiles staged into shared memory by 1-D TMA
//                    bulk copies (cp.async.bulk + mbarrier, double buffered); one thread owns the
//                    four roots of one iteration, so the per-iteration argmin needs no shuffles
//   k_score_models   same inner loop for arbitrary model lists x huge correspondence sets
//                    (config C5), correspondences split across CTAs, deterministic 2-stage reduce
//   k_chain          one warp per pair: FP64 certification + the sequential part of LO-MSAC
//   k_init_pairs     state initialisation
// Reference mapping: see include/ssfm.h and ssfm_chain.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssfm.h"
#include "ssfm_chain.cuh"

namespace ssfm {

// ------------------------------------------------------------------------------------------
// Warp execution context for the chain templates.
// ------------------------------------------------------------------------------------------
struct WarpCtx {
  int ln;
  __host__ __device__ int lane() const { return ln; }
  __host__ __device__ int width() const { return 32; }
  __host__ __device__ double sum(double x) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
#endif
    return x;
  }
  __host__ __device__ int sum_i(int x) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
#endif
    return x;
  }
  __host__ __device__ unsigned ballot(bool p) const {
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xffffffffu, p);
#else
    return p;
#endif
  }
  // min(init, min of x over lanes below this one)
  __host__ __device__ float prefix_min_excl(float x, float init) const {
#if defined(__CUDA_ARCH__)
    float v = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(0xffffffffu, v, d);
      if (ln >= d) v = fminf(v, o);
    }
    float ex = __shfl_up_sync(0xffffffffu, v, 1);
    if (ln == 0) ex = INFINITY;
    return fminf(ex, init);
#else
    return init;
#endif
  }
  __host__ __device__ float min_f(float x) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
#endif
    return x;
  }
  __host__ __device__ void sync() const {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
};

// ------------------------------------------------------------------------------------------
// k_pack
// ------------------------------------------------------------------------------------------
__global__ void k_pack(const double* __restrict__ rays, long long m, float4* __restrict__ u4, float4* __restrict__ v4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double2* src = reinterpret_cast<const double2*>(rays + 6 * i);  // 48-byte records, 16-byte aligned
  const double2 a = src[0], b = src[1], c = src[2];
  u4[i] = make_float4((float)a.x, (float)a.y, (float)b.x, 0.f);
  v4[i] = make_float4((float)b.y, (float)c.x, (float)c.y, 0.f);
}

// ------------------------------------------------------------------------------------------
// k_init_pairs
// ------------------------------------------------------------------------------------------
__global__ void k_init_pairs(Params P, const long long* __restrict__ offsets, int pair0, int npairs, PairState* states,
                             uint32_t* mt, int* active, int* navail, int first_cap) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= npairs) return;
  const int pair = pair0 + a;
  const int n = (int)(offsets[pair + 1] - offsets[pair]);
  PairState st;
  init_state(P, n, st);
  const uint32_t want = iterations_wanted(P, st);
  if (want == 0) st.done = 1;
  states[a] = st;
  active[a] = a;
  navail[a] = (int)(want < (uint32_t)first_cap ? want : (uint32_t)first_cap);
  if (P.driver == 0) mt19937_seed(mt + (size_t)a * 625, P.seed);
}

// ------------------------------------------------------------------------------------------
// k_sample_solve: grid.x = active pairs, grid.y = ceil(R / blockDim.x)
// models layout: [a][m*6+i][R]  (SoA over the look-ahead slot so both this kernel's stores and the
// scoring kernel's loads are coalesced)
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(64) k_sample_solve(Params P, const double* __restrict__ rays,
                                                     const long long* __restrict__ offsets, int pair0,
                                                     const int* __restrict__ active, const int* __restrict__ navail,
                                                     const PairState* __restrict__ states, int R,
                                                     double* __restrict__ models) {
  const int a = active[blockIdx.x];
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= navail[a]) return;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  const uint32_t it = states[a].it + (uint32_t)j;
  int idx[3];
  philox_sample<3>(P.seed, P.first_pair_id + (uint32_t)pair, it, 3, n, idx);
  double c[3][6];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const double2* src = reinterpret_cast<const double2*>(rays + 6 * (off + idx[s]));
    const double2 x0 = src[0], x1 = src[1], x2 = src[2];
    c[s][0] = x0.x; c[s][1] = x0.y; c[s][2] = x1.x; c[s][3] = x1.y; c[s][4] = x2.x; c[s][5] = x2.y;
  }
  double m[4][6];
  solve_minimal<KIND>(c[0], c[0] + 3, c[1], c[1] + 3, c[2], c[2] + 3, m);
  double* dst = models + (size_t)a * 24 * R + j;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[(size_t)(k * 6 + i) * R] = m[k][i];
}

// ------------------------------------------------------------------------------------------
// FP32 scoring core.
// ------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 128;  // one look-ahead iteration (4 roots) per thread
constexpr int kTile = 512;          // correspondences per shared-memory stage (2 x 8 KB)
constexpr int kStages = 2;

struct ScoreSmem {
  float4 u[kStages][kTile];
  float4 v[kStages][kTile];
  unsigned long long bar[kStages];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One correspondence against one structured model p (E = [p0 p1 p2; p1 -p0 p3; p4 p5 0]):
// squared Sampson distance, 22 FMA-pipe ops + 1 MUFU.RCP (src/spherical_estimator.cpp:67-78).
__device__ __forceinline__ float sampson_f32(const float (&p)[6], const float4 u, const float4 v) {
  const float Eu0 = fmaf(p[2], u.z, fmaf(p[1], u.y, p[0] * u.x));
  const float Eu1 = fmaf(p[3], u.z, fmaf(-p[0], u.y, p[1] * u.x));
  const float Eu2 = fmaf(p[5], u.y, p[4] * u.x);
  const float Et0 = fmaf(p[4], v.z, fmaf(p[1], v.y, p[0] * v.x));
  const float Et1 = fmaf(p[5], v.z, fmaf(-p[0], v.y, p[1] * v.x));
  const float d = fmaf(v.z, Eu2, fmaf(v.y, Eu1, v.x * Eu0));
  const float den = fmaf(Et1, Et1, fmaf(Et0, Et0, fmaf(Eu1, Eu1, Eu0 * Eu0)));
  return __fdividef(d * d, den);
}

// Streams correspondences [c0, c1) of one pair through shared memory and accumulates the MSAC
// cost (and optionally the inlier count) of the calling thread's four models.
template <bool COUNT>
__device__ __forceinline__ void score_stream(ScoreSmem& sm, const float4* __restrict__ u4, const float4* __restrict__ v4,
                                             long long c0, long long c1, const float (&p)[4][6], float thr, float (&acc)[4],
                                             int (&cnt)[4]) {
  const int ntiles = (int)((c1 - c0 + kTile - 1) / kTile);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int s = t % kStages;
    const long long b = c0 + (long long)t * kTile;
    const uint32_t n = (uint32_t)((c1 - b) < kTile ? (c1 - b) : kTile);
    mbar_expect_tx(&sm.bar[s], n * 32u);
    tma_load_1d(&sm.u[s][0], u4 + b, n * 16u, &sm.bar[s]);
    tma_load_1d(&sm.v[s][0], v4 + b, n * 16u, &sm.bar[s]);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&sm.bar[s], (uint32_t)((t / kStages) & 1));
    const long long b = c0 + (long long)t * kTile;
    const int n = (int)((c1 - b) < kTile ? (c1 - b) : kTile);
    const float4* su = sm.u[s];
    const float4* sv = sm.v[s];
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      const float4 u = su[i];
      const float4 v = sv[i];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float e = sampson_f32(p[m], u, v);
        acc[m] += fminf(e, thr);
        if (COUNT) cnt[m] += (e < thr) ? 1 : 0;
      }
    }
    __syncthreads();  // everyone is done with stage s before it is refilled
    if (threadIdx.x == 0 && t + kStages < ntiles) issue(t + kStages);
  }
}

// grid.x = active pairs, grid.y = ceil(R / kScoreThreads)
__global__ void __launch_bounds__(kScoreThreads) k_score_rounds(const float4* __restrict__ u4, const float4* __restrict__ v4,
                                                                const long long* __restrict__ offsets, int pair0,
                                                                const int* __restrict__ active,
                                                                const int* __restrict__ navail, int R,
                                                                const double* __restrict__ models, float thr,
                                                                float* __restrict__ s32) {
  __shared__ __align__(128) ScoreSmem sm;
  const int a = active[blockIdx.x];
  const int na = navail[a];
  const int j0 = blockIdx.y * kScoreThreads;
  if (j0 >= na) return;  // uniform for the CTA
  const int j = j0 + threadIdx.x;
  const bool live = j < na;
  const int pair = pair0 + a;
  float p[4][6];
  const double* src = models + (size_t)a * 24 * R + (live ? j : j0);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 6; ++i) p[m][i] = (float)src[(size_t)(m * 6 + i) * R];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt[4] = {0, 0, 0, 0};
  score_stream<false>(sm, u4, v4, offsets[pair], offsets[pair + 1], p, thr, acc, cnt);
  if (live) {
    float best = INFINITY;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float s = (p[m][0] == p[m][0]) ? acc[m] : INFINITY;  // absent / NaN model
      if (s < best) best = s;
    }
    s32[(size_t)a * R + j] = best;
  }
}

// Arbitrary models x one pair.  grid.x = ceil(M / (4 * kScoreThreads)), grid.y = correspondence chunks.
// part_score / part_cnt: [chunk][M]
__global__ void __launch_bounds__(kScoreThreads) k_score_models(const float4* __restrict__ u4, const float4* __restrict__ v4,
                                                                long long n, int chunk, const double* __restrict__ models6,
                                                                int M, float thr, float* __restrict__ part_score,
                                                                int* __restrict__ part_cnt) {
  __shared__ __align__(128) ScoreSmem sm;
  const int m0 = (blockIdx.x * kScoreThreads + threadIdx.x) * 4;
  float p[4][6];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 6; ++i) p[m][i] = (m0 + m < M) ? (float)models6[(size_t)(m0 + m) * 6 + i] : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt[4] = {0, 0, 0, 0};
  const long long c0 = (long long)blockIdx.y * chunk;
  const long long c1 = (c0 + chunk) < n ? (c0 + chunk) : n;
  score_stream<true>(sm, u4, v4, c0, c1, p, thr, acc, cnt);
#pragma unroll
  for (int m = 0; m < 4; ++m)
    if (m0 + m < M) {
      part_score[(size_t)blockIdx.y * M + m0 + m] = acc[m];
      part_cnt[(size_t)blockIdx.y * M + m0 + m] = cnt[m];
    }
}

__global__ void k_reduce_parts(const float* __restrict__ part_score, const int* __restrict__ part_cnt, int nchunks, int M,
                               float* __restrict__ scores, int* __restrict__ counts) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = 0.f;
  int c = 0;
  for (int k = 0; k < nchunks; ++k) {  // fixed order -> deterministic
    s += part_score[(size_t)k * M + m];
    c += part_cnt[(size_t)k * M + m];
  }
  scores[m] = s;
  counts[m] = c;
}

// ------------------------------------------------------------------------------------------
// k_chain: one warp per active pair.
// ------------------------------------------------------------------------------------------
constexpr int kChainWarps = 4;

__global__ void __launch_bounds__(kChainWarps * 32) k_chain(Params P, const double* __restrict__ rays,
                                                             const long long* __restrict__ offsets, int pair0,
                                                             const int* __restrict__ active, int nactive,
                                                             int* __restrict__ navail, PairState* states, int R,
                                                             const double* __restrict__ models, const float* __restrict__ s32,
                                                             int* list_a, int* list_b, uint32_t* mt, long long list_base,
                                                             unsigned char* flags, SsfmPairResult* results,
                                                             int* __restrict__ next_active, int* next_count, int next_cap,
                                                             unsigned long long* counters) {
  const int w = blockIdx.x * kChainWarps + (threadIdx.x >> 5);
  if (w >= nactive) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = active[w];
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  PairView pv{rays + 6 * off, n};
  Scratch sc{list_a + (off - list_base), list_b + (off - list_base), mt + (size_t)a * 625};
  PairState st = states[a];
  const int na = navail[a];
  const uint32_t it_before = st.it;
  process_round(cx, P, pv, sc, st, models + (size_t)a * 24 * R, R, s32 + (size_t)a * R, na);
  uint32_t want = iterations_wanted(P, st);
  if (!st.done && want == 0) st.done = 1;
  if (cx.lane() == 0) {
    // accounting: look-ahead hypotheses executed this round vs. the ones the reference loop used
    atomicAdd(&counters[0], (unsigned long long)na * 4ull * (unsigned long long)n);
    (void)it_before;
  }
  if (st.done) {
    double r[3], t[3];
    const int status = finalize_pair(cx, P, pv, sc, st, r, t, flags ? flags + (off - list_base) : (unsigned char*)0);
    if (cx.lane() == 0) {
      SsfmPairResult& o = results[a];
      for (int i = 0; i < 9; ++i) o.E[i] = st.E_best[i];
      for (int i = 0; i < 3; ++i) { o.r[i] = r[i]; o.t[i] = t[i]; }
      o.best_model_score = st.best_model_score;
      o.inlier_ratio = st.inlier_ratio;
      o.num_iterations = st.it;
      o.best_num_inliers = st.best_num_inliers;
      o.number_lo_iterations = st.num_lo;
      o.status = status;
      o.evals = (long long)st.it * 4ll * (long long)n;
      atomicAdd(&counters[1], (unsigned long long)st.evals_exact);
      navail[a] = 0;
      states[a] = st;
    }
  } else if (cx.lane() == 0) {
    states[a] = st;
    navail[a] = (int)(want < (uint32_t)next_cap ? want : (uint32_t)next_cap);
    const int pos = atomicAdd(next_count, 1);
    next_active[pos] = a;
  }
}

// Pairs that are finished before any round (n < 3): write their records.
__global__ void k_finish_trivial(Params P, const long long* __restrict__ offsets, int pair0, int npairs,
                                 const PairState* __restrict__ states, unsigned char* flags, long long list_base,
                                 SsfmPairResult* results, int* active_out, int* count_out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= npairs) return;
  const PairState st = states[a];
  if (!st.done) {
    const int pos = atomicAdd(count_out, 1);
    active_out[pos] = a;
    return;
  }
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  SsfmPairResult o;
  for (int i = 0; i < 9; ++i) o.E[i] = 0.0;
  for (int i = 0; i < 3; ++i) { o.r[i] = 0.0; o.t[i] = 0.0; }
  o.best_model_score = P.driver == 2 ? INFINITY : kDblMax;
  o.inlier_ratio = 0.0;
  o.num_iterations = 0;
  o.best_num_inliers = 0;
  o.number_lo_iterations = 0;
  o.status = n < 3 ? SSFM_PAIR_TOO_FEW_POINTS : SSFM_PAIR_NO_MODEL;
  o.evals = 0;
  results[a] = o;
  if (flags)
    for (int i = 0; i < n; ++i) flags[off - list_base + i] = 0;
}

// ------------------------------------------------------------------------------------------
// Hook kernels (parity tests drive the pieces one at a time).
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void k_solve_samples(const double* __restrict__ rays, const int* __restrict__ samples, int ns,
                                double* __restrict__ models, int* __restrict__ nmodels) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const double* c0 = rays + 6 * (size_t)samples[3 * s];
  const double* c1 = rays + 6 * (size_t)samples[3 * s + 1];
  const double* c2 = rays + 6 * (size_t)samples[3 * s + 2];
  double a[3][6];
  for (int i = 0; i < 6; ++i) { a[0][i] = c0[i]; a[1][i] = c1[i]; a[2][i] = c2[i]; }
  double m[4][6];
  const int nm = solve_minimal<KIND>(a[0], a[0] + 3, a[1], a[1] + 3, a[2], a[2] + 3, m);
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 6; ++i) models[(size_t)s * 24 + k * 6 + i] = m[k][i];
  nmodels[s] = nm;
}

__global__ void k_score_exact(const double* __restrict__ E9, int M, const double* __restrict__ rays, int n, double thr,
                              double* __restrict__ scores, int* __restrict__ counts) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= M) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  long long ev = 0;
  double E[9];
  for (int i = 0; i < 9; ++i) E[i] = E9[(size_t)w * 9 + i];
  const double s = msac_score_exact(cx, E, rays, n, thr, &ev);
  const int c = collect_inliers(cx, E, rays, n, thr, false, (int*)0, (unsigned char*)0, &ev);
  if (cx.lane() == 0) { scores[w] = s; counts[w] = c; }
}

__global__ void k_least_squares(const double* __restrict__ rays, const int* __restrict__ idx,
                                const int* __restrict__ sample_offsets, int nprob, int inward, double* E9) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nprob) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  double E[9];
  for (int i = 0; i < 9; ++i) E[i] = E9[(size_t)w * 9 + i];
  least_squares(cx, rays, idx + sample_offsets[w], sample_offsets[w + 1] - sample_offsets[w], inward != 0, E);
  if (cx.lane() == 0)
    for (int i = 0; i < 9; ++i) E9[(size_t)w * 9 + i] = E[i];
}

__global__ void k_non_minimal(const double* __restrict__ rays, int n, const int* __restrict__ idx,
                              const int* __restrict__ sample_offsets, int nprob, double* E9, int* ok) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nprob) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  PairView pv{rays, n};
  double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool good = non_minimal_solver(cx, pv, idx + sample_offsets[w], sample_offsets[w + 1] - sample_offsets[w], E);
  if (cx.lane() == 0) {
    for (int i = 0; i < 9; ++i) E9[(size_t)w * 9 + i] = E[i];
    ok[w] = good ? 1 : 0;
  }
}

__global__ void k_decompose(const double* __restrict__ E9, int num, int inward, double* r3, double* t3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  double E[9], r[3], t[3];
  for (int k = 0; k < 9; ++k) E[k] = E9[(size_t)i * 9 + k];
  decompose_spherical_E(E, inward != 0, r, t);
  for (int k = 0; k < 3; ++k) { r3[(size_t)i * 3 + k] = r[k]; t3[(size_t)i * 3 + k] = t[k]; }
}

__global__ void k_lo_shuffle(uint32_t seed, int ncalls, const int* __restrict__ sizes, const int* __restrict__ targets,
                             uint32_t* mt, int* work, int* out) {
  WarpCtx cx{(int)(threadIdx.x & 31)};
  if (cx.lane() == 0) mt19937_seed(mt, seed);
  cx.sync();
  int o = 0;
  for (int c = 0; c < ncalls; ++c) {
    for (int i = cx.lane(); i < sizes[c]; i += 32) work[i] = i;
    cx.sync();
    shuffle_and_resize(cx, mt, work, sizes[c]);
    for (int i = cx.lane(); i < targets[c]; i += 32) out[o + i] = work[i];
    cx.sync();
    o += targets[c];
  }
}

// FP32 peak probe: 8 independent FFMA chains per thread.
__global__ void k_fma_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace ssfm
