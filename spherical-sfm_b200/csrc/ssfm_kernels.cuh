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
#include "ssfm_triangulate.cuh"

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
  __host__ __device__ int slane() const { return ln; }
  __host__ __device__ int swidth() const { return 32; }
  template <int K>
  __host__ __device__ void sum_vec(double (&v)[K]) const {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = sum(v[k]);
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

// Look-ahead of the next round: the iterations the pair still wants, rounded up to a whole warp of
// scoring lanes (the extra lanes would idle otherwise; slots past max_iters are simply never consumed),
// capped by the round size (a multiple of 32).
__host__ __device__ inline int lookahead(uint32_t want, int cap) {
  if (want == 0) return 0;
  const uint32_t r = (want + 31u) & ~31u;
  return (int)(r < (uint32_t)cap ? r : (uint32_t)cap);
}

// ------------------------------------------------------------------------------------------
// k_pack
// ------------------------------------------------------------------------------------------
// uv4 (the unit-z plane) is always written; the general planes u4 / v4 only on request (k_pack_general), i.e. only for
// batches that turn out to contain rays with z != 1.
// xy64: the same four numbers in float64 (32 bytes per correspondence), what the chain's exact passes stream for a unit-z batch.
__global__ void k_pack(const double* __restrict__ rays, long long m, float4* __restrict__ uv4, double* __restrict__ xy64,
                       int* __restrict__ not_unit_z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double2* src = reinterpret_cast<const double2*>(rays + 6 * i);  // 48-byte records, 16-byte aligned
  const double2 a = src[0], b = src[1], c = src[2];
  uv4[i] = make_float4((float)a.x, (float)a.y, (float)b.y, (float)c.x);
  if (xy64) {
    double2* d = reinterpret_cast<double2*>(xy64 + 4 * i);
    d[0] = a;
    d[1] = make_double2(b.y, c.x);
  }
  // pipeline rays are K^-1 (x, y, 1) (examples/spherical_sfm_tools.cpp:364-373): z == 1 exactly
  if (b.x != 1.0 || c.y != 1.0) *not_unit_z = 1;
}
// SSFM_RAYS_F32 input: 6 floats per correspondence -> the float64 records the exact passes read (widening is exact) + the
// same planes as k_pack.
__global__ void k_pack_f32(const float* __restrict__ rays32, long long m, double* __restrict__ rays64, float4* __restrict__ uv4,
                           double* __restrict__ xy64, int* __restrict__ not_unit_z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float2* src = reinterpret_cast<const float2*>(rays32 + 6 * i);  // 24-byte records, 8-byte aligned
  const float2 a = src[0], b = src[1], c = src[2];
  double2* dst = reinterpret_cast<double2*>(rays64 + 6 * i);
  dst[0] = make_double2((double)a.x, (double)a.y);
  dst[1] = make_double2((double)b.x, (double)b.y);
  dst[2] = make_double2((double)c.x, (double)c.y);
  uv4[i] = make_float4(a.x, a.y, b.y, c.x);
  if (xy64) {
    double2* d = reinterpret_cast<double2*>(xy64 + 4 * i);
    d[0] = make_double2((double)a.x, (double)a.y);
    d[1] = make_double2((double)b.y, (double)c.x);
  }
  if (b.x != 1.0f || c.y != 1.0f) *not_unit_z = 1;
}
__global__ void k_pack_general(const double* __restrict__ rays, long long m, float4* __restrict__ u4, float4* __restrict__ v4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double2* src = reinterpret_cast<const double2*>(rays + 6 * i);
  const double2 a = src[0], b = src[1], c = src[2];
  u4[i] = make_float4((float)a.x, (float)a.y, (float)b.x, 0.f);
  v4[i] = make_float4((float)b.y, (float)c.x, (float)c.y, 0.f);
}

// ------------------------------------------------------------------------------------------
// k_build_rays: the ray construction of estimate_pairwise (examples/spherical_sfm_tools.cpp:357-376) on the
// device.  One thread per match: loc = Kinv * (x, y, 1) in float64 with Eigen's row.col accumulation order
// ((k0 x + k1 y) + k2), no FMA contraction, so the rays are bit-identical to the host's.
// ------------------------------------------------------------------------------------------
__global__ void k_build_rays(const float2* __restrict__ kp, const long long* __restrict__ kp_off,
                             const int* __restrict__ pair_images, const long long* __restrict__ match_off, int npairs,
                             const int2* __restrict__ matches, long long m0, long long m, const double* __restrict__ Kinv,
                             double* __restrict__ rays, float4* __restrict__ uv4, double* __restrict__ xy64,
                             int* __restrict__ not_unit_z, int* __restrict__ bad_index) {
  const long long i = m0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m0 + m) return;
  int lo = 0, hi = npairs;  // last pair whose first match is <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (match_off[mid] <= i) lo = mid; else hi = mid;
  }
  const int img0 = pair_images[2 * lo], img1 = pair_images[2 * lo + 1];
  const int2 mt = matches[i];
  const long long b0 = kp_off[img0], b1 = kp_off[img1];
  if (mt.x < 0 || mt.y < 0 || mt.x >= kp_off[img0 + 1] - b0 || mt.y >= kp_off[img1 + 1] - b1) {
    *bad_index = 1;  // reported as SSFM_ERR_INVALID by the host
    return;
  }
  const float2 p0 = kp[b0 + mt.x];
  const float2 p1 = kp[b1 + mt.y];
  double k[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) k[q] = Kinv[q];
  double out[6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    out[r] = add_rn(add_rn(mul_rn(k[3 * r], (double)p0.x), mul_rn(k[3 * r + 1], (double)p0.y)), k[3 * r + 2]);
    out[3 + r] = add_rn(add_rn(mul_rn(k[3 * r], (double)p1.x), mul_rn(k[3 * r + 1], (double)p1.y)), k[3 * r + 2]);
  }
  double2* dst = reinterpret_cast<double2*>(rays + 6 * i);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
  dst[2] = make_double2(out[4], out[5]);
  // the FP32 plane of the scoring kernel, in the same pass (what k_pack would produce from these rays)
  uv4[i] = make_float4((float)out[0], (float)out[1], (float)out[3], (float)out[4]);
  if (xy64) {
    double2* d = reinterpret_cast<double2*>(xy64 + 4 * i);
    d[0] = make_double2(out[0], out[1]);
    d[1] = make_double2(out[3], out[4]);
  }
  if (out[2] != 1.0 || out[5] != 1.0) *not_unit_z = 1;
}

// ------------------------------------------------------------------------------------------
// k_init_pairs
// ------------------------------------------------------------------------------------------
__global__ void k_init_pairs(Params P, const long long* __restrict__ offsets, int pair0, int npairs, PairState* states,
                             uint32_t* mt, int* active, int* ident, int* navail, int first_cap) {
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
  ident[a] = a;
  navail[a] = lookahead(want, first_cap);
  if (P.driver == 0) mt19937_seed(mt + (size_t)a * 625, P.seed);
}

// ------------------------------------------------------------------------------------------
// k_sample_solve: grid.x = active pairs, grid.y = ceil(R / blockDim.x)
// models layout: [a][m*6+i][R]  (SoA over the look-ahead slot so both this kernel's stores and the
// scoring kernel's loads are coalesced)
// ------------------------------------------------------------------------------------------
#ifndef SSFM_SOLVE_MINBLOCKS
#define SSFM_SOLVE_MINBLOCKS 4
#endif
#ifndef SSFM_SOLVE_THREADS
#define SSFM_SOLVE_THREADS 64
#endif
#ifndef SSFM_SOLVE_SYNC
#define SSFM_SOLVE_SYNC 0  // 1: block barriers between the solver's stages (action-matrix solver only)
#endif
constexpr int kSolveThreads = SSFM_SOLVE_THREADS;
struct BlockStageSync {
  __host__ __device__ void operator()() const {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#endif
  }
};
template <int KIND>
__global__ void __launch_bounds__(kSolveThreads, SSFM_SOLVE_MINBLOCKS) k_sample_solve(Params P, const double* __restrict__ rays,
                                                     const long long* __restrict__ offsets, int pair0,
                                                     const int* __restrict__ active, const int* __restrict__ navail,
                                                     const PairState* __restrict__ states, int R,
                                                     double* __restrict__ models) {
  const int a = active[blockIdx.x];
  const int na = navail[a];
  int j = blockIdx.y * blockDim.x + threadIdx.x;
  constexpr bool kSync = SSFM_SOLVE_SYNC != 0 && KIND == 0;
  if (kSync) {
    if ((int)(blockIdx.y * blockDim.x) >= na) return;  // the whole block is past the look-ahead (uniform)
  } else if (j >= na) {
    return;
  }
  const bool valid = j < na;
  if (!valid) j = na - 1;  // barrier variant: surplus threads solve the last slot again and store nothing
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  const uint32_t it = states[a].it + (uint32_t)j;
  int idx[3];
  if (P.driver == 2)  // msac.h:83 draws with random_sample (selection sampling on rand()), the RansacLib drivers with DrawSample
    knuth_sample(P.seed, P.first_pair_id + (uint32_t)pair, it, n, 3, idx);
  else
    philox_sample<3>(P.seed, P.first_pair_id + (uint32_t)pair, it, 3, n, idx);
  double c[3][6];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const double2* src = reinterpret_cast<const double2*>(rays + 6 * (off + idx[s]));
    const double2 x0 = src[0], x1 = src[1], x2 = src[2];
    c[s][0] = x0.x; c[s][1] = x0.y; c[s][2] = x1.x; c[s][3] = x1.y; c[s][4] = x2.x; c[s][5] = x2.y;
  }
  double m[4][6];
  if (kSync)
    solve_minimal<KIND, BlockStageSync>(c[0], c[0] + 3, c[1], c[1] + 3, c[2], c[2] + 3, m, P.skip_complex != 0);
  else
    solve_minimal<KIND>(c[0], c[0] + 3, c[1], c[1] + 3, c[2], c[2] + 3, m, P.skip_complex != 0);
  if (!valid) return;
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
constexpr int kTile = 512;          // correspondences per shared-memory stage
constexpr int kStages = 2;
#ifndef SSFM_PACKED_SCORING
#define SSFM_PACKED_SCORING 1
#endif
constexpr bool kPackedScoring = SSFM_PACKED_SCORING != 0;  // unit-z scoring loop on FFMA2 (0: the scalar FFMA loop)

// UNITZ = every ray of the batch has z == 1 exactly (the pipeline's K^-1 (x,y,1) rays and the
// reference's generator): one float4 (u0,u1,v0,v1) per correspondence and 14 FMA-pipe instructions per PAIR of
// evaluations in the packed loop (score_stream); otherwise two float4 (u.xyz, v.xyz) and 24 per evaluation.
#ifndef SSFM_SCORE_EXPANDED
#define SSFM_SCORE_EXPANDED 2  // unit-z packed loop: 0 = plain, 1 = v-side denominator expanded, 2 = d and the whole denominator
#endif                         // expanded over per-correspondence products computed once per tile (score_stream)
constexpr bool kScoreDerived = SSFM_SCORE_EXPANDED == 2 && kPackedScoring;
template <bool UNITZ>
struct ScoreSmem {
  float4 a[kStages][kTile];
  // general rays: the v plane.  Unit-z with kScoreDerived: the tile's derived plane (zx - wy, zy + wx, x^2+y^2+z^2+w^2, -)
  float4 b[(UNITZ && !kScoreDerived) ? 1 : kStages][(UNITZ && !kScoreDerived) ? 1 : kTile];
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
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One correspondence against one structured model p (E = [p0 p1 p2; p1 -p0 p3; p4 p5 0]):
// squared Sampson distance (src/spherical_estimator.cpp:67-78).
__device__ __forceinline__ float sampson_f32(const float (&p)[6], const float4 u, const float4 v) {
  const float Eu0 = fmaf(p[2], u.z, fmaf(p[1], u.y, p[0] * u.x));
  const float Eu1 = fmaf(p[3], u.z, fmaf(-p[0], u.y, p[1] * u.x));
  const float Eu2 = fmaf(p[5], u.y, p[4] * u.x);
  const float Et0 = fmaf(p[4], v.z, fmaf(p[1], v.y, p[0] * v.x));
  const float Et1 = fmaf(p[5], v.z, fmaf(-p[0], v.y, p[1] * v.x));
  const float d = fmaf(v.z, Eu2, fmaf(v.y, Eu1, v.x * Eu0));
  const float den = fmaf(Et1, Et1, fmaf(Et0, Et0, fmaf(Eu1, Eu1, Eu0 * Eu0)));
  return (d * d) * rcp_ftz(den);
}
// Same with u = (x, y, 1), v = (z, w, 1) packed as one float4.
__device__ __forceinline__ float sampson_f32_unitz(const float (&p)[6], const float4 c) {
  const float Eu0 = fmaf(p[1], c.y, fmaf(p[0], c.x, p[2]));
  const float Eu1 = fmaf(-p[0], c.y, fmaf(p[1], c.x, p[3]));
  const float Eu2 = fmaf(p[5], c.y, p[4] * c.x);
  const float Et0 = fmaf(p[1], c.w, fmaf(p[0], c.z, p[4]));
  const float Et1 = fmaf(-p[0], c.w, fmaf(p[1], c.z, p[5]));
  const float d = fmaf(c.w, Eu1, fmaf(c.z, Eu0, Eu2));
  const float den = fmaf(Et1, Et1, fmaf(Et0, Et0, fmaf(Eu1, Eu1, Eu0 * Eu0)));
  return (d * d) * rcp_ftz(den);
}

// ---- packed FP32 (sm_100 fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2) ----
// One instruction carries two independent FP32 operations, halving the issue slots of the scoring loop (the FMA pipe
// still retires 128 lanes-FMAs per clock per SM: the gain is that the loop stops being issue-bound).  The two halves
// of a register pair hold two MODELS (roots) of the calling thread; the correspondence is broadcast to both halves.
// Every packed operation is the same IEEE operation as its scalar twin, in the same order, so results are bit-identical.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// Two models against one unit-z correspondence: P[0..5] = (p_i of model a, p_i of model b), P[6] = -P[0];
// X, Y, Z, W = the correspondence's (u.x, u.y, v.x, v.y) broadcast to both halves.  Same expression tree as
// sampson_f32_unitz; returns d^2 and the denominator (the reciprocal is a scalar MUFU per half).
__device__ __forceinline__ void sampson2_unitz(const f32x2 (&P)[7], f32x2 X, f32x2 Y, f32x2 Z, f32x2 W, f32x2& d2, f32x2& den) {
  const f32x2 Eu0 = fma2(P[1], Y, fma2(P[0], X, P[2]));
  const f32x2 Eu1 = fma2(P[6], Y, fma2(P[1], X, P[3]));
  const f32x2 Eu2 = fma2(P[5], Y, mul2(P[4], X));
  const f32x2 Et0 = fma2(P[1], W, fma2(P[0], Z, P[4]));
  const f32x2 Et1 = fma2(P[6], W, fma2(P[1], Z, P[5]));
  const f32x2 d = fma2(W, Eu1, fma2(Z, Eu0, Eu2));
  den = fma2(Et1, Et1, fma2(Et0, Et0, fma2(Eu1, Eu1, mul2(Eu0, Eu0))));
  d2 = mul2(d, d);
}

// The same, with the v-side half of the denominator expanded: for E = [p0 p1 p2; p1 -p0 p3; p4 p5 0]
//   (E^T v)_0^2 + (E^T v)_1^2 = (p0^2 + p1^2)(z^2 + w^2) + 2 (p0 p4 + p1 p5) z + 2 (p1 p4 - p0 p5) w + (p4^2 + p5^2),
// four per-model constants Q (computed once per thread) and one per-correspondence z^2 + w^2 shared by the thread's four
// models: 3 FMA for that half instead of 6 (the two products themselves are not needed by anything else), 16 FMA-pipe
// instructions per evaluation pair instead of 19.  The u-side half stays a sum of squares because d needs E u anyway.
// This is the FP32 pre-filter (and config C5's FP32 score): the float64 certification is untouched.
__device__ __forceinline__ void sampson2_unitz_expanded(const f32x2 (&P)[7], const f32x2 (&Q)[4], f32x2 X, f32x2 Y, f32x2 Z, f32x2 W,
                                                        f32x2 T, f32x2& d2, f32x2& den) {
  const f32x2 Eu0 = fma2(P[1], Y, fma2(P[0], X, P[2]));
  const f32x2 Eu1 = fma2(P[6], Y, fma2(P[1], X, P[3]));
  const f32x2 Eu2 = fma2(P[5], Y, mul2(P[4], X));
  const f32x2 d = fma2(W, Eu1, fma2(Z, Eu0, Eu2));
  const f32x2 denT = fma2(Q[0], T, fma2(Q[1], Z, fma2(Q[2], W, Q[3])));
  den = fma2(Eu1, Eu1, fma2(Eu0, Eu0, denT));
  d2 = mul2(d, d);
}

// Everything bilinear expanded.  With u = (x, y, 1), v = (z, w, 1) and E as above
//   d   = v^T E u = p0 (zx - wy) + p1 (zy + wx) + p2 z + p3 w + p4 x + p5 y
//   den = (p0^2 + p1^2)(x^2 + y^2 + z^2 + w^2) + 2 (p0 p2 + p1 p3) x + 2 (p1 p2 - p0 p3) y + 2 (p0 p4 + p1 p5) z
//         + 2 (p1 p4 - p0 p5) w + (p2^2 + p3^2 + p4^2 + p5^2)
// The three per-correspondence products (zx - wy, zy + wx, x^2 + y^2 + z^2 + w^2) do not depend on the model: they are computed
// ONCE per tile by the block (four correspondences per thread) into a second shared-memory plane, so an evaluation pair is
// 10 FFMA2 + 3 FMUL2 + 1 FADD2 = 14 FMA-pipe instructions (19 for the plain form, 16 with only the v side expanded).
// Q = (A, Bu, Cu, Bt, Ct, D) per model pair.
__device__ __forceinline__ void sampson2_unitz_bilinear(const f32x2 (&P)[7], const f32x2 (&Q)[6], f32x2 X, f32x2 Y, f32x2 Z, f32x2 W,
                                                        f32x2 ZXWY, f32x2 ZYWX, f32x2 ST, f32x2& d2, f32x2& den) {
  const f32x2 d = fma2(P[0], ZXWY, fma2(P[1], ZYWX, fma2(P[2], Z, fma2(P[3], W, fma2(P[4], X, mul2(P[5], Y))))));
  den = fma2(Q[0], ST, fma2(Q[1], X, fma2(Q[2], Y, fma2(Q[3], Z, fma2(Q[4], W, Q[5])))));
  d2 = mul2(d, d);
}

// Streams correspondences [c0, c1) of one pair through shared memory and accumulates the MSAC
// cost (and optionally the inlier count) of the calling thread's four models.  Warps whose lanes
// are all idle (`active` false) only take part in the barriers.  Partial sums are folded per tile
// so the FP32 accumulation error stays ~(kTile + ntiles) ulp.
// min(e, thr) that PROPAGATES a NaN residual (FMNMX.NAN, same cost as fminf): the reference's std::min(e, thr) returns the NaN
// and such a model never wins; with fminf it would count as thr, get a finite FP32 score and could lower the running
// minimum of the pre-filter.  A NaN FP32 score is never a candidate and never enters the running minimum.
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <bool UNITZ, bool COUNT>
__device__ __forceinline__ void score_stream(ScoreSmem<UNITZ>& sm, const float4* __restrict__ pa, const float4* __restrict__ pb,
                                             long long c0, long long c1, const float (&p)[4][6], float thr, bool active,
                                             float (&acc)[4], int (&cnt)[4]) {
  const int ntiles = (int)((c1 - c0 + kTile - 1) / kTile);
  f32x2 P2[2][7];  // models (0,1) and (2,3) packed parameter by parameter; [6] = -p0
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 6; ++i) P2[h][i] = pack2(p[2 * h][i], p[2 * h + 1][i]);
    P2[h][6] = pack2(-p[2 * h][0], -p[2 * h + 1][0]);
  }
  f32x2 Q2[2][6];
  if (UNITZ && kPackedScoring && SSFM_SCORE_EXPANDED) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float q[2][6];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float* m = p[2 * h + k];
        const float A = m[0] * m[0] + m[1] * m[1];
        const float Bt = 2.f * (m[0] * m[4] + m[1] * m[5]), Ct = 2.f * (m[1] * m[4] - m[0] * m[5]);
        const float Dt = m[4] * m[4] + m[5] * m[5];
        if (SSFM_SCORE_EXPANDED == 2) {
          q[k][0] = A;
          q[k][1] = 2.f * (m[0] * m[2] + m[1] * m[3]);
          q[k][2] = 2.f * (m[1] * m[2] - m[0] * m[3]);
          q[k][3] = Bt;
          q[k][4] = Ct;
          q[k][5] = (m[2] * m[2] + m[3] * m[3]) + Dt;
        } else {
          q[k][0] = A; q[k][1] = Bt; q[k][2] = Ct; q[k][3] = Dt; q[k][4] = 0.f; q[k][5] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) Q2[h][i] = pack2(q[0][i], q[1][i]);
    }
  }
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
    mbar_expect_tx(&sm.bar[s], UNITZ ? n * 16u : n * 32u);
    tma_load_1d(&sm.a[s][0], pa + b, n * 16u, &sm.bar[s]);
    if (!UNITZ) tma_load_1d(&sm.b[UNITZ ? 0 : s][0], pb + b, n * 16u, &sm.bar[s]);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&sm.bar[s], (uint32_t)((t / kStages) & 1));
    if (UNITZ && kScoreDerived) {  // the tile's model-independent products, once per block (every warp takes part)
      const long long b0 = c0 + (long long)t * kTile;
      const int nt = (int)((c1 - b0) < kTile ? (c1 - b0) : kTile);
      for (int i = threadIdx.x; i < nt; i += kScoreThreads) {
        const float4 ca = sm.a[s][i];
        sm.b[s][i] = make_float4(fmaf(ca.z, ca.x, -ca.w * ca.y), fmaf(ca.z, ca.y, ca.w * ca.x),
                                 fmaf(ca.w, ca.w, fmaf(ca.z, ca.z, fmaf(ca.y, ca.y, ca.x * ca.x))), 0.f);
      }
      __syncthreads();
    }
    if (active) {
      const long long b = c0 + (long long)t * kTile;
      const int n = (int)((c1 - b) < kTile ? (c1 - b) : kTile);
      const float4* sa = sm.a[s];
      const float4* sb = sm.b[(UNITZ && !kScoreDerived) ? 0 : s];
      if (UNITZ && kPackedScoring) {
        f32x2 tacc2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll 8
        for (int i = 0; i < n; ++i) {
          const float4 ca = sa[i];
          const f32x2 X = pack2(ca.x, ca.x), Y = pack2(ca.y, ca.y), Z = pack2(ca.z, ca.z), W = pack2(ca.w, ca.w);
#if SSFM_SCORE_EXPANDED == 2
          const float4 cd = sb[i];
          const f32x2 ZXWY = pack2(cd.x, cd.x), ZYWX = pack2(cd.y, cd.y), ST = pack2(cd.z, cd.z);
#elif SSFM_SCORE_EXPANDED == 1
          const float tt = fmaf(ca.w, ca.w, ca.z * ca.z);
          const f32x2 T = pack2(tt, tt);
#endif
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            f32x2 d2, den;
#if SSFM_SCORE_EXPANDED == 2
            sampson2_unitz_bilinear(P2[h], Q2[h], X, Y, Z, W, ZXWY, ZYWX, ST, d2, den);
#elif SSFM_SCORE_EXPANDED == 1
            {
              const f32x2 Q4[4] = {Q2[h][0], Q2[h][1], Q2[h][2], Q2[h][3]};
              sampson2_unitz_expanded(P2[h], Q4, X, Y, Z, W, T, d2, den);
            }
#else
            sampson2_unitz(P2[h], X, Y, Z, W, d2, den);
#endif
            float dl, dh;
            unpack2(den, dl, dh);
            const f32x2 e2 = mul2(d2, pack2(rcp_ftz(dl), rcp_ftz(dh)));
            float el, eh;
            unpack2(e2, el, eh);
            tacc2[h] = add2(tacc2[h], pack2(min_nan(el, thr), min_nan(eh, thr)));
            if (COUNT) {
              cnt[2 * h] += (el < thr) ? 1 : 0;
              cnt[2 * h + 1] += (eh < thr) ? 1 : 0;
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float tl, th;
          unpack2(tacc2[h], tl, th);
          acc[2 * h] += tl;
          acc[2 * h + 1] += th;
        }
      } else {
        float tacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int i = 0; i < n; ++i) {
          const float4 ca = sa[i];
          float4 cb;
          if (!UNITZ) cb = sb[i];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const float e = UNITZ ? sampson_f32_unitz(p[m], ca) : sampson_f32(p[m], ca, cb);
            tacc[m] += min_nan(e, thr);
            if (COUNT) cnt[m] += (e < thr) ? 1 : 0;
          }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[m] += tacc[m];
      }
    }
    __syncthreads();  // everyone is done with stage s before it is refilled
    if (threadIdx.x == 0 && t + kStages < ntiles) issue(t + kStages);
  }
}

// grid.x = active pairs, grid.y = ceil(R / kScoreThreads)
// s32 [a][R]: min over the four roots; s32m [a][4][R]: the four costs.
template <bool UNITZ>
__global__ void __launch_bounds__(kScoreThreads) k_score_rounds(const float4* __restrict__ pa, const float4* __restrict__ pb,
                                                                const long long* __restrict__ offsets, int pair0,
                                                                const int* __restrict__ active,
                                                                const int* __restrict__ navail, int R,
                                                                const double* __restrict__ models, float thr,
                                                                float* __restrict__ s32, float* __restrict__ s32m) {
  __shared__ __align__(128) ScoreSmem<UNITZ> sm;
  const int a = active[blockIdx.x];
  const int na = navail[a];
  const int j0 = blockIdx.y * kScoreThreads;
  if (j0 >= na) return;  // uniform for the CTA
  const int j = j0 + threadIdx.x;
  const bool live = j < na;
  const bool warp_live = (j0 + (int)(threadIdx.x & ~31u)) < na;
  const int pair = pair0 + a;
  float p[4][6];
  const double* src = models + (size_t)a * 24 * R + (live ? j : j0);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 6; ++i) p[m][i] = (float)src[(size_t)(m * 6 + i) * R];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt[4] = {0, 0, 0, 0};
  score_stream<UNITZ, false>(sm, pa, pb, offsets[pair], offsets[pair + 1], p, thr, warp_live, acc, cnt);
  if (live) {
    float best = INFINITY;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float s = (p[m][0] == p[m][0]) ? acc[m] : INFINITY;  // absent / NaN model
      s32m[((size_t)a * 4 + m) * R + j] = s;
      if (s < best) best = s;
    }
    s32[(size_t)a * R + j] = best;
  }
}

// Arbitrary models x pairs (config C5).  grid.x = ceil(M / (4 * kScoreThreads)), grid.y = correspondence chunks,
// grid.z = pairs.  Pair z owns correspondences [offsets[z], offsets[z+1]) and the models models6[z][M][6].
// part_score / part_cnt: [pair][chunk][M]; chunks past the end of a (shorter) pair write zeros.
template <bool UNITZ>
__global__ void __launch_bounds__(kScoreThreads) k_score_models(const float4* __restrict__ pa, const float4* __restrict__ pb,
                                                                const long long* __restrict__ offsets, int chunk,
                                                                const double* __restrict__ models6, int M, float thr,
                                                                float* __restrict__ part_score, int* __restrict__ part_cnt) {
  __shared__ __align__(128) ScoreSmem<UNITZ> sm;
  const int pair = blockIdx.z;
  const long long base = offsets[pair], n = offsets[pair + 1] - base;
  const int m0 = (blockIdx.x * kScoreThreads + threadIdx.x) * 4;
  float p[4][6];
  const double* mp = models6 + (size_t)pair * M * 6;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 6; ++i) p[m][i] = (m0 + m < M) ? (float)mp[(size_t)(m0 + m) * 6 + i] : 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt[4] = {0, 0, 0, 0};
  const long long c0 = (long long)blockIdx.y * chunk;
  const long long c1 = (c0 + chunk) < n ? (c0 + chunk) : n;
  const bool warp_live = (int)((blockIdx.x * kScoreThreads + (threadIdx.x & ~31u)) * 4) < M;
  if (c0 < n) score_stream<UNITZ, true>(sm, pa, pb, base + c0, base + c1, p, thr, warp_live, acc, cnt);  // uniform per CTA
  const size_t row = ((size_t)pair * gridDim.y + blockIdx.y) * M;
#pragma unroll
  for (int m = 0; m < 4; ++m)
    if (m0 + m < M) {
      part_score[row + m0 + m] = acc[m];
      part_cnt[row + m0 + m] = cnt[m];
    }
}

// scores / counts: [pair][M]
__global__ void k_reduce_parts(const float* __restrict__ part_score, const int* __restrict__ part_cnt, int nchunks, int M,
                               float* __restrict__ scores, int* __restrict__ counts) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = blockIdx.y;
  if (m >= M) return;
  float s = 0.f;
  int c = 0;
  for (int k = 0; k < nchunks; ++k) {  // fixed order -> deterministic
    s += part_score[((size_t)pair * nchunks + k) * M + m];
    c += part_cnt[((size_t)pair * nchunks + k) * M + m];
  }
  scores[(size_t)pair * M + m] = s;
  counts[(size_t)pair * M + m] = c;
}

// ------------------------------------------------------------------------------------------
// k_chain: one warp per listed pair.
//   DEFER = false: everything inline (used when num_lo_steps > 0: RansacLib's full LO schedule).
//   DEFER = true : refits are parked as tasks (see ssfm_chain.cuh); the host alternates
//                  k_chain<true> / k_refit_small + k_refit_big until no pair of the round is parked.
// list addressing: mode 0 = list[w]; mode 1 = the parked list of the previous wave: small tasks were
// appended from the front (list[0..n_front)), big ones from the back (list[cap-1-k]).
// ------------------------------------------------------------------------------------------
constexpr int kChainWarps = 4;
constexpr int kSmallRefit = 32;  // refits with at most this many residuals get one thread each

struct ChainArgs {
  const double* rays;
  const double* xy64;  // compact float64 plane (u.x, u.y, v.x, v.y) of a unit-z batch, or NULL
  const long long* offsets;
  int pair0;
  const int* list;
  int nlist, mode, n_front, cap;
  int* navail;
  PairState* states;
  int R;
  const double* models;
  const float* s32;
  const float* s32m;
  int* list_a;
  int* list_b;
  uint32_t* mt;
  double* lm_E;
  long long list_base;
  unsigned char* flags;
  SsfmPairResult* results;
  int* next_active;
  int* next_count;
  int next_cap;
  int* parked;        // output task list (front: small, back: big)
  int* parked_small;  // counters
  int* parked_big;
  unsigned long long* counters;
};

#ifndef SSFM_CHAIN_MINBLOCKS
#define SSFM_CHAIN_MINBLOCKS 8  // 64 registers: this kernel is latency bound, occupancy pays (chain stage 124 -> 117 ms)
#endif
template <bool DEFER>
__global__ void __launch_bounds__(kChainWarps * 32, DEFER ? SSFM_CHAIN_MINBLOCKS : 4) k_chain(Params P, ChainArgs A) {
  const int w = blockIdx.x * kChainWarps + (threadIdx.x >> 5);
  if (w >= A.nlist) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = A.mode == 0 ? A.list[w] : (w < A.n_front ? A.list[w] : A.list[A.cap - 1 - (w - A.n_front)]);
  const int pair = A.pair0 + a;
  const long long off = A.offsets[pair];
  const int n = (int)(A.offsets[pair + 1] - off);
  PairView pv{A.rays + 6 * off, n, A.xy64 ? compact_stream(A.xy64 + 4 * off) : A.rays + 6 * off};
  Scratch sc{A.list_a + (off - A.list_base), A.list_b + (off - A.list_base), A.mt + (size_t)a * 625, A.counters,
             A.lm_E + (size_t)a * 9};
#if defined(SSFM_PROFILE_CHAIN)
  const long long t_start = clock64();
#endif
  PairState st = A.states[a];
  const int na = A.navail[a];
  const bool fresh = st.phase == PHASE_NONE;  // first visit of this pair in this round
  bool parked = false;
  if (!(DEFER && (st.phase == PHASE_LO_LATE || st.phase == PHASE_FINAL_LSQ))) {  // those resume inside finalize_pair
    process_round<WarpCtx, DEFER>(cx, P, pv, sc, st, A.models + (size_t)a * 24 * A.R, A.R, A.s32 + (size_t)a * A.R,
                                  A.s32m + (size_t)a * 4 * A.R, na);
    parked = DEFER && st.phase != PHASE_NONE;
  }
  if (fresh && cx.lane() == 0) atomicAdd(&A.counters[0], (unsigned long long)na * 4ull * (unsigned long long)n);
  uint32_t want = 0;
  if (!parked) {
    want = iterations_wanted(P, st);
    if (!st.done && want == 0) st.done = 1;
    if (st.done) {
      double r[3], t[3];
      const int status = finalize_pair<WarpCtx, DEFER>(cx, P, pv, sc, st, r, t, A.flags ? A.flags + (off - A.list_base) : (unsigned char*)0);
      if (status < 0) {
        parked = true;
      } else if (cx.lane() == 0) {
        SsfmPairResult& o = A.results[a];
        for (int i = 0; i < 9; ++i) o.E[i] = st.E_best[i];
        for (int i = 0; i < 3; ++i) { o.r[i] = r[i]; o.t[i] = t[i]; }
        o.best_model_score = st.best_model_score;
        o.inlier_ratio = st.inlier_ratio;
        o.num_iterations = st.it;
        o.best_num_inliers = st.best_num_inliers;
        o.number_lo_iterations = st.num_lo;
        o.status = status;
        o.evals = (long long)st.it * 4ll * (long long)n;
        o.focal = 0.0;
        atomicAdd(&A.counters[1], (unsigned long long)st.evals_exact);
        A.navail[a] = 0;
      }
    }
  }
  if (cx.lane() == 0) {
    A.states[a] = st;
    if (parked) {
      if (st.lm_n <= kSmallRefit) A.parked[atomicAdd(A.parked_small, 1)] = a;
      else A.parked[A.cap - 1 - atomicAdd(A.parked_big, 1)] = a;
    } else if (!st.done) {
      A.navail[a] = lookahead(want, A.next_cap);
      A.next_active[atomicAdd(A.next_count, 1)] = a;
    }
#if defined(SSFM_PROFILE_CHAIN)
    atomicAdd(&A.counters[PH_TOTAL], (unsigned long long)(clock64() - t_start));
#endif
  }
}

// Parked refits, SphericalEstimator::LeastSquares (src/spherical_estimator.cpp:110-157).
// small: one THREAD per refit (<= kSmallRefit residuals), 32 independent LMs per warp.  Iteration
// counts vary from ~6 to 200 between problems, so lanes pull tasks from a queue: a lane whose LM
// terminates fetches the next task while its neighbours keep iterating.
#ifndef SSFM_REFIT_MINBLOCKS
#define SSFM_REFIT_MINBLOCKS 4
#endif
// A refit that has not converged after kHandover trust-region iterations is a straggler (the median needs 11,
// 1 % need more than 40, the cap is 200 -- and the slowest refit of a wave sets the wave's duration): its LMState
// is written out and k_refit_long continues it with a whole warp (residuals across lanes).  The switch depends
// only on the refit's own iteration count, so results do not depend on what else is in the batch.
constexpr int kHandover = 24;  // default; SSFM_HANDOVER overrides for sweeps

__global__ void __launch_bounds__(64, SSFM_REFIT_MINBLOCKS) k_refit_small(Params P, const double* __restrict__ rays,
                                                    const long long* __restrict__ offsets, int pair0,
                                                    const int* __restrict__ parked, int ntasks, int* queue_head,
                                                    const PairState* __restrict__ states, const int* __restrict__ list_a,
                                                    long long list_base, double* lm_E, LMState* __restrict__ lm_states,
                                                    int* __restrict__ long_list, int* long_count, int handover_at,
                                                    double* __restrict__ stage /* kSmallRefit * 6 doubles per thread, or NULL */) {
  SerialCtx cx;
  // The lane's own copy of its refit's correspondences, contiguous (L1-resident across the ~30 passes of a refit) instead of
  // up to 32 scattered 48-byte records re-gathered through L2 on every pass.  Same values, same order: results unchanged.
  double* mine = stage ? stage + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (kSmallRefit * 6) : (double*)0;
  const bool inward = P.inward != 0;
  int task = ntasks, a = 0, n = 0;
  bool have = false;
  const double* ry = rays;
  const int* smp = list_a;
  LMState S;
  for (;;) {
    if (!have) {
      task = atomicAdd(queue_head, 1);
      if (task < ntasks) {
        a = parked[task];
        const long long off = offsets[pair0 + a];
        ry = rays + 6 * off;
        smp = list_a + (off - list_base);
        n = states[a].lm_n;
        if (mine && n <= kSmallRefit) {
          for (int i = 0; i < n; ++i) {
            double c[6];
            load6(ry + 6 * (size_t)smp[i], c);
            double2* d = reinterpret_cast<double2*>(mine + 6 * i);
            d[0] = make_double2(c[0], c[1]);
            d[1] = make_double2(c[2], c[3]);
            d[2] = make_double2(c[4], c[5]);
          }
          ry = mine;
          smp = (const int*)0;
        }
        double E[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] = lm_E[(size_t)a * 9 + i];
        lm_init(cx, ry, smp, n, inward, E, S);
        have = true;
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;
    if (have) {
      if (lm_step(cx, ry, smp, n, S)) {
        double E[9];
        lm_finish(S, inward, E);
#pragma unroll
        for (int i = 0; i < 9; ++i) lm_E[(size_t)a * 9 + i] = E[i];
        have = false;
      } else if (S.iteration >= handover_at && long_list != nullptr) {
        lm_states[a] = S;
        long_list[atomicAdd(long_count, 1)] = a;
        have = false;
      }
    }
  }
}

#ifndef SSFM_REFITBIG_MINBLOCKS
#define SSFM_REFITBIG_MINBLOCKS 1
#endif
// Stragglers handed over by k_refit_small: persistent warps pull them from the list.
__global__ void __launch_bounds__(128, SSFM_REFITBIG_MINBLOCKS) k_refit_long(Params P, const double* __restrict__ rays,
                                                    const long long* __restrict__ offsets, int pair0,
                                                    const int* __restrict__ long_list, const int* __restrict__ long_count,
                                                    int* queue_head, const PairState* __restrict__ states,
                                                    const int* __restrict__ list_a, long long list_base, double* lm_E,
                                                    const LMState* __restrict__ lm_states) {
  WarpCtx cx{(int)(threadIdx.x & 31)};
  __shared__ __align__(16) double long_stage[4][kSmallRefit * 6];
  const int ntasks = *long_count;
  const bool inward = P.inward != 0;
  for (;;) {
    int task = 0;
    if (cx.lane() == 0) task = atomicAdd(queue_head, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= ntasks) break;
    const int a = long_list[task];
    const long long off = offsets[pair0 + a];
    const double* ry = rays + 6 * off;
    const int* smp = list_a + (off - list_base);
    const int n = states[a].lm_n;
    LMState S = lm_states[a];
    if (n <= kSmallRefit) {  // always, by construction: gather the residuals' correspondences once (see k_refit_big)
      double* mine = long_stage[threadIdx.x >> 5];
      __syncwarp();
      for (int i = cx.lane(); i < n; i += 32) {
        double c[6];
        load6(ry + 6 * (size_t)smp[i], c);
#pragma unroll
        for (int q = 0; q < 6; ++q) mine[6 * i + q] = c[q];
      }
      __syncwarp();
      while (!lm_step(cx, mine, (const int*)0, n, S)) {
      }
    } else {
      while (!lm_step(cx, ry, smp, n, S)) {
      }
    }
    double E[9];
    lm_finish(S, inward, E);
    if (cx.lane() == 0)
#pragma unroll
      for (int i = 0; i < 9; ++i) lm_E[(size_t)a * 9 + i] = E[i];
  }
}
constexpr int kRefitStage = 512;  // correspondences staged per warp (24 KB); larger refits read through L2 as before
constexpr size_t kRefitBigSmem = (size_t)4 * kRefitStage * 6 * sizeof(double);
// big: one WARP per refit (the final least squares over all inliers).
// Also used for the SMALL refits of a wave that has too few of them to fill the machine with one
// thread each (from_front = 1): a warp per refit has ~3x lower latency, and those waves are pure tail.
__global__ void __launch_bounds__(128, SSFM_REFITBIG_MINBLOCKS) k_refit_big(Params P, const double* __restrict__ rays,
                                                   const long long* __restrict__ offsets, int pair0,
                                                   const int* __restrict__ parked, int cap, int ntasks, int from_front,
                                                   const PairState* __restrict__ states, const int* __restrict__ list_a,
                                                   long long list_base, double* lm_E) {
  extern __shared__ double refit_stage[];  // kRefitStage gathered correspondences per warp
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= ntasks) return;
  WarpCtx cx{(int)(threadIdx.x & 31)};
  const int a = from_front ? parked[w] : parked[cap - 1 - w];
  const long long off = offsets[pair0 + a];
  double E[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) E[i] = lm_E[(size_t)a * 9 + i];
  cx.sync();
  const int n = states[a].lm_n;
  const double* ry = rays + 6 * off;
  const int* smp = list_a + (off - list_base);
  if (n <= kRefitStage) {
    // Every trust-region iteration reads the refit's correspondences twice (Jacobian pass + candidate cost): gather them
    // once into shared memory instead of ~30 gathers through L2.  Same values, same order: the result is unchanged.
    double* mine = refit_stage + (size_t)(threadIdx.x >> 5) * kRefitStage * 6;
    for (int i = cx.lane(); i < n; i += 32) {
      double c[6];
      load6(ry + 6 * (size_t)smp[i], c);
#pragma unroll
      for (int q = 0; q < 6; ++q) mine[6 * i + q] = c[q];
    }
    __syncwarp();
    least_squares(cx, mine, (const int*)0, n, P.inward != 0, E);
  } else {
    least_squares(cx, ry, smp, n, P.inward != 0, E);
  }
  if (cx.lane() == 0)
#pragma unroll
    for (int i = 0; i < 9; ++i) lm_E[(size_t)a * 9 + i] = E[i];
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
  o.status = n < 3 ? SSFM_PAIR_TOO_FEW_POINTS : (n < P.min_points ? SSFM_PAIR_SKIPPED : SSFM_PAIR_NO_MODEL);
  o.evals = 0;
  o.focal = 0.0;
  results[a] = o;
  if (flags)
    for (int i = 0; i < n; ++i) flags[off - list_base + i] = 0;
}

// ------------------------------------------------------------------------------------------
// Retriangulate (ssfm_triangulate.cuh): camera table, then one thread per point.
// ------------------------------------------------------------------------------------------
__global__ void k_tri_cameras(const double* __restrict__ cam_tr, int nc, tri::Cam* __restrict__ cams) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  tri::Cam c;
  tri::make_camera(cam_tr + 6 * i, cam_tr + 6 * i + 3, c);  // Pose::Pose (src/sfm_types.cpp:14-19)
  cams[i] = c;
}

// list == nullptr: the points order[point0 .. point0+npoints) (first phase); otherwise the compacted list of points
// left unfinished by the previous phase.  A point that reaches `stop_at` iterations without finishing is appended
// to next_list, its loop state saved, and continued by the next launch among points of similar length.
#ifndef SSFM_TRI_MINBLOCKS
#define SSFM_TRI_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(64, SSFM_TRI_MINBLOCKS) k_retriangulate(Params P, const tri::Cam* __restrict__ cams,
                                                      const long long* __restrict__ offsets, const int* __restrict__ obs_cam,
                                                      const double* __restrict__ obs_xy, double focal, int point0, int npoints,
                                                      const int* __restrict__ order, const int* __restrict__ list,
                                                      int* __restrict__ next_list, int* __restrict__ next_count,
                                                      unsigned int stop_at, tri::LoState* __restrict__ states,
                                                      int* __restrict__ scratch, uint32_t* __restrict__ mt,
                                                      double* __restrict__ points, int* __restrict__ num_inliers,
                                                      int* __restrict__ status, unsigned int* __restrict__ iterations) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const bool exists = w < npoints;  // every lane of a warp takes part in lo_msac_run's votes: no early return
  const int a = exists ? (list ? list[w] : w) : 0;  // slot of this point within the pass (state / mt index)
  const int pt = order[point0 + a];                 // points sorted by track length: the lanes of a warp loop over similar n
  const long long off = offsets[pt];
  const int n = exists ? (int)(offsets[pt + 1] - off) : 0;
  const bool active = exists && n >= 3;  // src/sfm.cpp:173
  tri::View v{cams, obs_cam + off, obs_xy + 2 * off, active ? n : 0, focal};
  int* sc = scratch + 4 * off;
  tri::Lists L{sc, sc + n, sc + 2 * n, sc + 3 * n, mt + (size_t)a * 625};
  tri::LoState S;
  S.started = 0;
  if (active && list) S = states[a];
  const bool finished = tri::lo_msac_run(P, v, L, P.first_pair_id + (uint32_t)pt, S, stop_at, active);
  if (!exists) return;
  if (active && !finished) {
    states[a] = S;
    next_list[atomicAdd(next_count, 1)] = a;
    return;
  }
  int code = SSFM_PAIR_SKIPPED, ninl = 0;
  unsigned int iters = 0;
  double X[3] = {0.0, 0.0, 0.0};  // SetPoint(j, Zero) (src/sfm.cpp:172)
  if (active) {
    ninl = S.st.best_num_inliers;
    iters = S.st.num_iterations;
    if (ninl < 3) {  // :186
      code = SSFM_PAIR_NO_MODEL;
    } else {
      code = SSFM_PAIR_OK;
      X[0] = S.X[0]; X[1] = S.X[1]; X[2] = S.X[2];
    }
  }
  points[3 * (size_t)pt] = X[0]; points[3 * (size_t)pt + 1] = X[1]; points[3 * (size_t)pt + 2] = X[2];
  num_inliers[pt] = ninl;
  status[pt] = code;
  if (iterations) iterations[pt] = iters;
}

// ------------------------------------------------------------------------------------------
// Hook kernels (parity tests drive the pieces one at a time).
// ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void k_solve_samples(const double* __restrict__ rays, const int* __restrict__ samples, int ns, int skip_complex,
                                double* __restrict__ models, int* __restrict__ nmodels) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const double* c0 = rays + 6 * (size_t)samples[3 * s];
  const double* c1 = rays + 6 * (size_t)samples[3 * s + 1];
  const double* c2 = rays + 6 * (size_t)samples[3 * s + 2];
  double a[3][6];
  for (int i = 0; i < 6; ++i) { a[0][i] = c0[i]; a[1][i] = c1[i]; a[2][i] = c2[i]; }
  double m[4][6];
  const int nm = solve_minimal<KIND>(a[0], a[0] + 3, a[1], a[1] + 3, a[2], a[2] + 3, m, skip_complex != 0);
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
  int c = 0;
  const double s = msac_score_exact(cx, E, rays, n, thr, &c, &ev);
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
  PairView pv{rays, n, rays};
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

// transform_image_matches (examples/spherical_sfm_tools.cpp:1118-1131): E' = T E T, T = diag(s, s, 1), decompose.
__global__ void k_decompose_rescaled(const double* __restrict__ E9, int num, const double* __restrict__ scales, int nscales,
                                     int inward, double* __restrict__ r3) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)num * nscales) return;
  const int pair = (int)(i % num), sc = (int)(i / num);
  const double s = scales[sc];
  const double T[3] = {s, s, 1.0};
  double E[9], r[3], t[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) E[3 * a + b] = (T[a] * E9[(size_t)pair * 9 + 3 * a + b]) * T[b];
  decompose_spherical_E(E, inward != 0, r, t);
  for (int k = 0; k < 3; ++k) r3[(size_t)i * 3 + k] = r[k];
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
    shuffle_and_resize(cx, mt, work, sizes[c], targets[c]);
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

// The same probe on the packed instruction (fma.rn.f32x2 -> FFMA2): 8 independent chains of two floats per thread.
__global__ void k_fma2_peak(float* out, int iters, float a, float b) {
  const f32x2 A = pack2(a, a), B = pack2(b, b);
  f32x2 x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = pack2((float)threadIdx.x + k, (float)threadIdx.x - k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = fma2(x[k], A, B);
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float lo, hi;
    unpack2(x[k], lo, hi);
    acc += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace ssfm
