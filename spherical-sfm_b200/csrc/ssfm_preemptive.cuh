// ssfm_preemptive.cuh -- pre-emptive RANSAC (sphericalsfm::PreemptiveRANSAC::compute,
// include/sphericalsfm/preemptive_ransac.h:46-139) for a batch of pairs.
//
//   k_preempt_hypotheses  one thread per (pair, hypothesis): selection sample of m+1 = 4 correspondences
//                         (random_sample, :8-28, Knuth 3.4.2S with a Philox-backed rand()), 3-point solver on
//                         the first three, the fourth picks the solution (:75-90)
//   k_preempt_select      one CTA per pair: block-wise inlier counting of the surviving hypotheses, halving
//                         of the survivors every B blocks (:95-120), final inlier mask of the winner (:122-137),
//                         pose
//
// All decisions are made in the reference's float64 arithmetic (sampson_exact); the counts are integers, so
// the result does not depend on thread scheduling.
#pragma once
#include "ssfm_kernels.cuh"

namespace ssfm {

// hyps layout: [pair-in-pass][6][M] (SoA over the hypothesis so stores coalesce); p[0] = NaN marks a
// hypothesis without a model.  has[pair][M]: 0 = the solver returned no solution (:72).
template <int KIND>
__global__ void __launch_bounds__(64, SSFM_SOLVE_MINBLOCKS)
    k_preempt_hypotheses(Params P, const double* __restrict__ rays, const long long* __restrict__ offsets, int pair0,
                         int M, double* __restrict__ hyps, unsigned char* __restrict__ has) {
  const int a = blockIdx.x;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int n = (int)(offsets[pair + 1] - off);
  if (n < 4 || n < P.min_points) return;
  int idx[4];
  knuth_sample(P.seed, P.first_pair_id + (uint32_t)pair, (uint32_t)i, n, 4, idx);
  double c[4][6];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double2* src = reinterpret_cast<const double2*>(rays + 6 * (off + idx[s]));
    const double2 x0 = src[0], x1 = src[1], x2 = src[2];
    c[s][0] = x0.x; c[s][1] = x0.y; c[s][2] = x1.x; c[s][3] = x1.y; c[s][4] = x2.x; c[s][5] = x2.y;
  }
  double m[4][6];
  solve_minimal<KIND>(c[0], c[0] + 3, c[1], c[1] + 3, c[2], c[2] + 3, m);
  // The action-matrix and polynomial solvers always return four matrices (some may be NaN); the Sturm variant
  // returns the real roots it kept, in order, the rest absent (NaN).
  int nsolns = 0, first = -1, best = -1;
  double best_score = INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool present = KIND != 2 || m[k][0] == m[k][0];
    if (!present) continue;
    ++nsolns;
    if (first < 0) first = k;
    double E[9];
    E_from_p(m[k], E);
    const double score = sampson_exact(E, c[3], c[3] + 3);
    if (score < best_score) {
      best_score = score;
      best = k;
    }
  }
  // nsolns == 1: upstream skips the disambiguation (:75) -- the single solution is the hypothesis.
  // nsolns > 1 and every score NaN: index 0 stays chosen (:78).
  const int pick = nsolns == 0 ? -1 : ((nsolns == 1 || best < 0) ? first : best);
  double* dst = hyps + (size_t)a * 6 * M + i;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double v = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k == pick) v = m[k][q];
    dst[(size_t)q * M] = v;
  }
  has[(size_t)a * M + i] = pick >= 0 ? 1 : 0;
}

// Descending bitonic sort of `len` (power of two) signed keys in shared memory.
__device__ inline void bitonic_sort_desc(long long* keys, int len) {
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < len; t += blockDim.x) {
        const int x = t ^ j;
        if (x > t) {
          const long long a = keys[t], b = keys[x];
          const bool desc = (t & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[t] = b;
            keys[x] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kPreemptThreads = 256;

// Dynamic shared memory: Mpad keys (count << 32 | index, so descending key order is std::greater on
// pair<int,size_t>, :113) + Mpad per-slot counters.
__global__ void __launch_bounds__(kPreemptThreads)
    k_preempt_select(Params P, const double* __restrict__ rays, const long long* __restrict__ offsets, int pair0, int M,
                     int Mpad, int B, const double* __restrict__ hyps, const unsigned char* __restrict__ has,
                     unsigned char* __restrict__ flags, SsfmPairResult* __restrict__ results,
                     unsigned long long* __restrict__ counters) {
  extern __shared__ long long smem_ll[];
  long long* keys = smem_ll;
  int* cnt = reinterpret_cast<int*>(keys + Mpad);
  __shared__ double red_s[kPreemptThreads / 32];
  __shared__ int red_c[kPreemptThreads / 32];
  __shared__ long long s_top;
  const int a = blockIdx.x;
  const int pair = pair0 + a;
  const long long off = offsets[pair];
  const int N = (int)(offsets[pair + 1] - off);
  const double* pr = rays + 6 * off;
  const int tid = threadIdx.x;
  SsfmPairResult o;
  if (N < 4 || N < P.min_points) {
    if (flags)
      for (int i = tid; i < N; i += blockDim.x) flags[off + i] = 0;
    if (tid == 0) {
      for (int i = 0; i < 9; ++i) o.E[i] = 0.0;
      for (int i = 0; i < 3; ++i) { o.r[i] = 0.0; o.t[i] = 0.0; }
      o.best_model_score = kDblMax;
      o.inlier_ratio = 0.0;
      o.num_iterations = 0;
      o.best_num_inliers = 0;
      o.number_lo_iterations = 0;
      o.status = N < 4 ? SSFM_PAIR_TOO_FEW_POINTS : SSFM_PAIR_SKIPPED;
      o.evals = 0;
      o.focal = 0.0;
      results[a] = o;
    }
    return;
  }
  const double* hp = hyps + (size_t)a * 6 * M;
  const unsigned char* hh = has + (size_t)a * M;
  for (int j = tid; j < Mpad; j += blockDim.x) {
    keys[j] = j < M ? (long long)j : -1LL;  // count 0; pads sort last
    cnt[j] = 0;
  }
  __syncthreads();
  long long evals = 0;
  int f = M, it = 0;
  // Group g = blocks i = gB+1 .. (g+1)B of the reference loop (:96): all are counted with the same f survivors
  // (f changes only when floor(i/B) does, :110), so their B*B observations are evaluated in one sweep.
  for (int g = 0;; ++g) {
    long long nblocks = (long long)N - 1 - (long long)g * B;  // the loop runs i = 1 .. N-1 (N >= 4 here)
    if (nblocks > B) nblocks = B;
    if (f <= 1) nblocks = 1;  // after the first block of a group with f <= 1 the reference stops (:116)
    long long it2 = (long long)it + nblocks * B;
    if (it2 > N) it2 = N;
    const int np = (int)(it2 - it);
    const long long items = (long long)f * np;
    for (long long wi = tid; wi < items; wi += blockDim.x) {
      const int j = (int)(wi / np);
      const int t = it + (int)(wi - (long long)j * np);
      const int h = (int)(keys[j] & 0xffffffffLL);
      if (!hh[h]) continue;
      double p[6], E[9];
#pragma unroll
      for (int q = 0; q < 6; ++q) p[q] = hp[(size_t)q * M + h];
      E_from_p(p, E);
      const double e = sampson_exact(E, pr + 6 * (size_t)t, pr + 6 * (size_t)t + 3);
      ++evals;
      if (e <= P.thr2) atomicAdd(&cnt[j], 1);
    }
    __syncthreads();
    for (int j = tid; j < f; j += blockDim.x) {
      keys[j] += (long long)cnt[j] << 32;
      cnt[j] = 0;
    }
    __syncthreads();
    it = (int)it2;
    const bool data_end = it == N;
    const long long i_last = (long long)g * B + nblocks;  // the last block index processed (if the data did not end first)
    const bool loop_end = i_last >= (long long)N - 1;     // i < N (:96)
    // f after that block (:110): floor(i/B) is g inside the group and g+1 at its last block
    const int f_new = (int)floor((double)M * pow(2.0, -floor((double)(i_last / B))));
    const bool finishing = f_new <= 1 || data_end || loop_end;
    if (f_new < f || finishing) {
      // partial_sort (:113): the first f_new positions hold the best by (count, index), descending
      int len = 1;
      while (len < f) len <<= 1;
      for (int j = f + tid; j < len; j += blockDim.x) keys[j] = -1LL;
      __syncthreads();
      bitonic_sort_desc(keys, len);
      if (f_new < f) f = f_new;
    }
    if (finishing) break;
  }
  __syncthreads();
  if (tid == 0) s_top = keys[0];
  __syncthreads();
  const int top = (int)(s_top & 0xffffffffLL);
  double E[9];
  {
    double p[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) p[q] = hp[(size_t)q * M + top];
    E_from_p(p, E);
  }
  const bool have = hh[top] != 0;
  // evaluate everything under the winner (:122-137); cost = the legacy MSAC cost (msac.h:56-64)
  double s = 0.0;
  int c = 0;
  if (have) {
    for (int i = tid; i < N; i += blockDim.x) {
      const double e = sampson_exact(E, pr + 6 * (size_t)i, pr + 6 * (size_t)i + 3);
      const bool in = e <= P.thr2;
      s += in ? e : P.thr2;
      c += in ? 1 : 0;
      if (flags) flags[off + i] = in ? 1 : 0;
    }
    evals += (N + blockDim.x - 1 - tid) / blockDim.x;
  } else if (flags) {
    for (int i = tid; i < N; i += blockDim.x) flags[off + i] = 0;
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, w);
    c += __shfl_xor_sync(0xffffffffu, c, w);
    evals += __shfl_xor_sync(0xffffffffu, evals, w);
  }
  __shared__ long long red_e[kPreemptThreads / 32];
  if ((tid & 31) == 0) { red_s[tid >> 5] = s; red_c[tid >> 5] = c; red_e[tid >> 5] = evals; }
  __syncthreads();
  if (tid == 0) {
    double S = 0.0;
    int Cn = 0;
    long long Ev = 0;
    for (int w = 0; w < kPreemptThreads / 32; ++w) { S += red_s[w]; Cn += red_c[w]; Ev += red_e[w]; }
    for (int i = 0; i < 9; ++i) o.E[i] = have ? E[i] : 0.0;
    for (int i = 0; i < 3; ++i) { o.r[i] = 0.0; o.t[i] = 0.0; }
    o.num_iterations = (uint32_t)M;
    o.number_lo_iterations = 0;
    o.evals = Ev;
    o.focal = 0.0;
    if (have) {
      o.best_model_score = S;
      o.best_num_inliers = Cn;
      o.inlier_ratio = (double)Cn / (double)N;
      o.status = SSFM_PAIR_OK;
      decompose_spherical_E(E, P.inward != 0, o.r, o.t);
    } else {
      o.best_model_score = kDblMax;
      o.best_num_inliers = 0;
      o.inlier_ratio = 0.0;
      o.status = SSFM_PAIR_NO_MODEL;
    }
    results[a] = o;
    atomicAdd(&counters[1], (unsigned long long)Ev);
  }
}

}  // namespace ssfm
