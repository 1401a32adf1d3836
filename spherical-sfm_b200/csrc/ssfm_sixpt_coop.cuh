// ssfm_sixpt_coop.cuh -- the six-point shared-focal minimal solver (ssfm_sixpt.cuh) restructured so that a GROUP of lanes
// works on one sample: the batched kernel runs 4 samples per warp (8 lanes each), with every sample's 16 x 16 companion /
// Hessenberg matrix in shared memory.  Why: one thread per sample needs 2 KB of indexed storage for the eigen-iteration, so
// only ~100 samples fit one SM -- as 100 threads that is 3 warps of latency-bound FP64 (measured: slower than keeping the
// matrices in local memory); as 100 groups of 8 lanes it is 25 warps.
//
// Same algorithm and the same floating-point formulas as the per-thread solver; what changes is the order of independent
// operations (row / column updates are spread over the lanes, the elimination to Hessenberg form is applied as
// "all columns, then all rows" instead of interleaved -- the elementary transformations of one step commute).
//
//   six_setup       lane 0: normalisation, 6 x 9 null space (sixpt::nullspace_6x9), products of the basis entries;
//                   all lanes: the ten cubics, one equation per lane  -> M[3][10][10] in a global scratch slot
//   six_companion   rows of Mk Q (Householder from F22, sixpt::companion16), LU with partial pivoting spread over the lanes
//   six_balance / six_hessenberg / six_hqr   the eigenvalues; scalars are computed redundantly by every lane of the group
//                   (same inputs, same instructions: identical), so control flow is group-uniform and needs no broadcast
//   six_polish      ONE LANE PER CANDIDATE eigenvalue (all candidates of the warp's samples, packed): least-squares start,
//                   Gauss-Newton on the ten equations, residual test
//   six_decompose   one lane per accepted solution: K F K -> (R, t) in front of both cameras
//
// The code is written against a small "group" interface so tests/hostshim can run it on the CPU with a one-lane group.
#pragma once
#include "ssfm_sixpt.cuh"

namespace ssfm {
namespace sixc {

// One lane does everything (host build, and the serial reference of the same code path).
struct SerialGroup {
  static constexpr int kSize = 1;
  SSFM_HD int lane() const { return 0; }
  SSFM_HD void sync() const {}
  SSFM_HD double sum(double v) const { return v; }
  SSFM_HD bool any(bool b) const { return b; }
};
#ifdef __CUDACC__
// Eight consecutive lanes of a warp.  Groups of one warp may sit in different branches: every collective names its own lanes.
struct LaneGroup8 {
  static constexpr int kSize = 8;
  int l;
  unsigned mask;
  __device__ __forceinline__ int lane() const { return l; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ double sum(double v) const {
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
  }
  __device__ __forceinline__ bool any(bool b) const { return (__ballot_sync(mask, b) & mask) != 0u; }
};
#endif

// 1 / x and 1 / sqrt(x) for the scalar part of the eigen-iteration.  Device: the hardware's approximation (2^-23) refined by
// two Newton steps (a few DFMA instead of the ~12-instruction IEEE division sequence with its special-case branches); the
// operands here are pivots and norms that are neither zero, infinite nor denormal (checked by the callers).  The result
// is within an ulp or two of the exact quotient, which is all a backward-stable iteration needs.  Host: exact.
SSFM_HD double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}
SSFM_HD double fast_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// Per-sample scratch (doubles).  T has a row stride of 17 so that a column walk (stride 17 doubles = 34 banks) is
// conflict-free; kScratch = 4 (mod 16) puts the four samples of a warp on different banks for the broadcast reads; 436
// doubles = 3.4 KB let two blocks of 32 samples share one SM.
constexpr int kTS = 17;
constexpr int kOffT = 0;                 // 16 x 17; after the eigenvalues: least-squares tableau (0..99), model list (140..244)
constexpr int kOffX = 16 * kTS;          // 100: L (10 x 10) while the companion matrix is built; then wr, wi, candidates
constexpr int kOffFb = kOffX + 100;      // 27
constexpr int kOffPts = kOffFb + 27;     // 36: the normalised sample (x1[6][3], x2[6][3])
constexpr int kScratch = 436;
constexpr int kOffMisc = kOffT + 14 * kTS + 16;  // two doubles in the unused 17th column of T's rows 14 and 15: scale | flag
constexpr int kMiscStride = kTS;
static_assert(kOffPts + 36 <= kScratch && kScratch % 16 == 4, "scratch layout");
constexpr int kLS = 10;                  // row stride of L
constexpr int kOffWr = kOffX, kOffWi = kOffX + 16, kOffCand = kOffX + 32;  // candidates: 16 x (x, y, w, state)
constexpr int kOffList = kOffT + 140;    // SixPointModel[15]
constexpr int kMSize = 300;              // M[3][10][10] in the global scratch slot of a sample

#define SIX_T(i, j) S[kOffT + (i) * kTS + (j)]
#define SIX_L(i, j) S[kOffX + (i) * kLS + (j)]

// symmetric 3 x 3 index -> 0..5
SSFM_HD int sym3(int i, int j) { return i <= j ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : sym3(j, i); }

// ----------------------------------------------------------------------------------------------------------------
// six_setup: returns false (group-uniform) if the sample is degenerate
// ----------------------------------------------------------------------------------------------------------------
template <class G>
SSFM_HD bool six_setup(const G& g, const double (*c)[6], double* S, double* Mg) {
  using namespace sixpt;
  if (g.lane() == 0) {
    double x1[6][3], x2[6][3];
    double s = 0.0;
    for (int i = 0; i < 6; ++i) s += c[i][0] * c[i][0] + c[i][1] * c[i][1] + c[i][3] * c[i][3] + c[i][4] * c[i][4];
    s = sqrt(s / 12.0);
    if (!(s > 0.0) || !(s < 1e300)) s = 1.0;
    const double is = 1.0 / s;
    for (int i = 0; i < 6; ++i) {
      x1[i][0] = c[i][0] * is; x1[i][1] = c[i][1] * is; x1[i][2] = c[i][2];
      x2[i][0] = c[i][3] * is; x2[i][1] = c[i][4] * is; x2[i][2] = c[i][5];
    }
    double Fb[3][9];
    const bool ok = nullspace_6x9_ws(x1, x2, Fb, reinterpret_cast<double(*)[9]>(S + kOffT));
    S[kOffMisc] = s;
    S[kOffMisc + kMiscStride] = ok ? 1.0 : 0.0;
    for (int i = 0; i < 6; ++i)
      for (int d = 0; d < 3; ++d) { S[kOffPts + 3 * i + d] = x1[i][d]; S[kOffPts + 18 + 3 * i + d] = x2[i][d]; }
    if (ok) {
      for (int k = 0; k < 3; ++k)
        for (int e = 0; e < 9; ++e) S[kOffFb + 9 * k + e] = Fb[k][e];
      // G0 = F diag(1,1,0) F^T, G1 = F diag(0,0,1) F^T as quadratics in (x, y); their traces with diag(1,1,w)
      double F[9][3];
      for (int e = 0; e < 9; ++e) { F[e][0] = Fb[0][e]; F[e][1] = Fb[1][e]; F[e][2] = Fb[2][e]; }
      double* G0 = S + kOffX;        // 6 x 6
      double* G1 = S + kOffX + 36;   // 6 x 6
      double* tr = S + kOffX + 72;   // 3 x 6 (90 of the 100 doubles of X)
      for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
          double a[6], b[6];
          mul11(F[3 * i + 0], F[3 * j + 0], a);
          mul11(F[3 * i + 1], F[3 * j + 1], b);
          for (int q = 0; q < 6; ++q) G0[6 * sym3(i, j) + q] = a[q] + b[q];
          mul11(F[3 * i + 2], F[3 * j + 2], a);
          for (int q = 0; q < 6; ++q) G1[6 * sym3(i, j) + q] = a[q];
        }
      for (int q = 0; q < 6; ++q) {
        tr[q] = G0[6 * sym3(0, 0) + q] + G0[6 * sym3(1, 1) + q];
        tr[6 + q] = G1[6 * sym3(0, 0) + q] + G1[6 * sym3(1, 1) + q] + G0[6 * sym3(2, 2) + q];
        tr[12 + q] = G1[6 * sym3(2, 2) + q];
      }
    }
  }
  g.sync();
  if (S[kOffMisc + kMiscStride] == 0.0) return false;
  // the ten cubics, one per lane: equation 0 = det F, 1 + 3 i + j = (2 H F - tr(H) F)_ij with H = F Q F^T Q, Q = diag(1,1,w)
  for (int e = g.lane(); e < 10; e += G::kSize) {
    double o0[10], o1[10], o2[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) o0[q] = o1[q] = o2[q] = 0.0;
    double F[9][3];
    for (int k = 0; k < 9; ++k) { F[k][0] = S[kOffFb + k]; F[k][1] = S[kOffFb + 9 + k]; F[k][2] = S[kOffFb + 18 + k]; }
    if (e == 0) {
      double m0[6], m1[6], cc[6];
      mul11(F[4], F[8], m0); mul11(F[5], F[7], m1);
      for (int q = 0; q < 6; ++q) cc[q] = m0[q] - m1[q];
      fma21(1.0, cc, F[0], o0);
      mul11(F[3], F[8], m0); mul11(F[5], F[6], m1);
      for (int q = 0; q < 6; ++q) cc[q] = m0[q] - m1[q];
      fma21(-1.0, cc, F[1], o0);
      mul11(F[3], F[7], m0); mul11(F[4], F[6], m1);
      for (int q = 0; q < 6; ++q) cc[q] = m0[q] - m1[q];
      fma21(1.0, cc, F[2], o0);
    } else {
      const int i = (e - 1) / 3, j = (e - 1) % 3;
      const double* G0 = S + kOffX;
      const double* G1 = S + kOffX + 36;
      const double* tr = S + kOffX + 72;
      fma21(2.0, G0 + 6 * sym3(i, 0), F[0 + j], o0);
      fma21(2.0, G0 + 6 * sym3(i, 1), F[3 + j], o0);
      fma21(-1.0, tr, F[3 * i + j], o0);
      fma21(2.0, G1 + 6 * sym3(i, 0), F[0 + j], o1);
      fma21(2.0, G1 + 6 * sym3(i, 1), F[3 + j], o1);
      fma21(2.0, G0 + 6 * sym3(i, 2), F[6 + j], o1);
      fma21(-1.0, tr + 6, F[3 * i + j], o1);
      fma21(2.0, G1 + 6 * sym3(i, 2), F[6 + j], o2);
      fma21(-1.0, tr + 12, F[3 * i + j], o2);
    }
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      Mg[(0 * 10 + e) * 10 + q] = o0[q];
      Mg[(1 * 10 + e) * 10 + q] = o1[q];
      Mg[(2 * 10 + e) * 10 + q] = o2[q];
    }
  }
  g.sync();
  return true;
}

// ----------------------------------------------------------------------------------------------------------------
// six_companion: the deflated 16 x 16 companion matrix (sixpt::companion16 explains the deflation) into T
// ----------------------------------------------------------------------------------------------------------------
template <class G>
SSFM_HD bool six_companion(const G& g, double* S, const double* Mg) {
  double* V = S + kOffT;  // Householder vectors (6 x 10) and betas (6): in T's first rows, which are written last
  if (g.lane() == 0) {
    const double fa = S[kOffFb + 8], fb = S[kOffFb + 17], fc = S[kOffFb + 26];
    double Bt[10][6];
    for (int i = 0; i < 10; ++i)
      for (int p = 0; p < 6; ++p) Bt[i][p] = 0.0;
    Bt[0][0] = fa; Bt[1][0] = fb; Bt[4][0] = fc;
    Bt[1][1] = fa; Bt[2][1] = fb; Bt[5][1] = fc;
    Bt[2][2] = fa; Bt[3][2] = fb; Bt[6][2] = fc;
    Bt[4][3] = fa; Bt[5][3] = fb; Bt[7][3] = fc;
    Bt[5][4] = fa; Bt[6][4] = fb; Bt[8][4] = fc;
    Bt[7][5] = fa; Bt[8][5] = fb; Bt[9][5] = fc;
    bool ok = true;
    for (int k = 0; k < 6; ++k) {
      double nrm = 0.0;
      for (int i = k; i < 10; ++i) nrm += Bt[i][k] * Bt[i][k];
      nrm = sqrt(nrm);
      if (!(nrm > 0.0)) { ok = false; break; }
      const double alpha = Bt[k][k] > 0 ? -nrm : nrm;
      for (int i = 0; i < k; ++i) V[10 * k + i] = 0.0;
      const double v0 = Bt[k][k] - alpha;
      V[10 * k + k] = v0;
      double vn = v0 * v0;
      for (int i = k + 1; i < 10; ++i) { V[10 * k + i] = Bt[i][k]; vn += Bt[i][k] * Bt[i][k]; }
      const double beta = vn > 0.0 ? 2.0 / vn : 0.0;
      V[60 + k] = beta;
      for (int j = k + 1; j < 6; ++j) {
        double d = 0.0;
        for (int i = k; i < 10; ++i) d += V[10 * k + i] * Bt[i][j];
        d *= beta;
        for (int i = k; i < 10; ++i) Bt[i][j] -= d * V[10 * k + i];
      }
    }
    S[kOffMisc + kMiscStride] = ok ? 1.0 : 0.0;
  }
  g.sync();
  if (S[kOffMisc + kMiscStride] == 0.0) return false;
  // rows of Mk Q = ((row H0) H1) ... H5: 30 rows over the lanes
  for (int r = g.lane(); r < 30; r += G::kSize) {
    const int which = r / 10, e = r - 10 * which;
    double row[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) row[q] = Mg[(which * 10 + e) * 10 + q];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double d = 0.0;
#pragma unroll
      for (int i = k; i < 10; ++i) d += row[i] * V[10 * k + i];
      d *= V[60 + k];
#pragma unroll
      for (int i = k; i < 10; ++i) row[i] -= d * V[10 * k + i];
    }
    if (which == 0) {
#pragma unroll
      for (int q = 0; q < 10; ++q) SIX_L(e, q) = row[q];
    } else if (which == 1) {
#pragma unroll
      for (int q = 0; q < 10; ++q) SIX_T(6 + e, 6 + q) = -row[q];
    } else {
#pragma unroll
      for (int q = 0; q < 6; ++q) SIX_T(6 + e, q) = -row[q];  // columns 6..9 of M2 Q vanish (rounding only)
    }
  }
  g.sync();
  // Gaussian elimination with partial pivoting on [A0 | X], X = rows 6..15 of T
  for (int k = 0; k < 10; ++k) {
    int piv = k;
    double best = fabs(SIX_L(k, k));
    for (int i = k + 1; i < 10; ++i) {
      const double v = fabs(SIX_L(i, k));
      if (v > best) { best = v; piv = i; }
    }
    if (!(best > 1e-300)) return false;  // every lane read the same column: uniform
    g.sync();
    if (piv != k) {
      for (int idx = g.lane(); idx < 26; idx += G::kSize) {
        if (idx < 10) { const double tt = SIX_L(k, idx); SIX_L(k, idx) = SIX_L(piv, idx); SIX_L(piv, idx) = tt; }
        else { const int j = idx - 10; const double tt = SIX_T(6 + k, j); SIX_T(6 + k, j) = SIX_T(6 + piv, j); SIX_T(6 + piv, j) = tt; }
      }
      g.sync();
    }
    const double inv = 1.0 / SIX_L(k, k);
    const int ncol = (9 - k) + 16, total = (9 - k) * ncol;
    for (int idx = g.lane(); idx < total; idx += G::kSize) {
      const int i = k + 1 + idx / ncol, cidx = idx - (i - k - 1) * ncol;
      const double f = SIX_L(i, k) * inv;
      if (cidx < 9 - k) { const int j = k + 1 + cidx; SIX_L(i, j) -= f * SIX_L(k, j); }
      else { const int j = cidx - (9 - k); SIX_T(6 + i, j) -= f * SIX_T(6 + k, j); }
    }
    g.sync();
  }
  for (int j = g.lane(); j < 16; j += G::kSize)
    for (int i = 9; i >= 0; --i) {
      double v = SIX_T(6 + i, j);
      for (int q = i + 1; q < 10; ++q) v -= SIX_L(i, q) * SIX_T(6 + q, j);
      SIX_T(6 + i, j) = v / SIX_L(i, i);
    }
  g.sync();
  bool bad = false;
  for (int idx = g.lane(); idx < 256; idx += G::kSize) {
    const int i = idx >> 4, j = idx & 15;
    if (i < 6) SIX_T(i, j) = (j == 6 + i) ? 1.0 : 0.0;
    else if (!(fabs(SIX_T(i, j)) < 1e300)) bad = true;
  }
  const bool any_bad = g.any(bad);
  g.sync();
  return !any_bad;
}

// Eigenvalue-preserving diagonal scaling (powers of two), sixpt::balance with the row / column sums spread over the lanes.
template <class G>
SSFM_HD void six_balance(const G& g, double* S) {
  for (int pass = 0; pass < 20; ++pass) {
    bool done = true;
    for (int i = 0; i < 16; ++i) {
      double r = 0.0, c = 0.0;
      for (int j = g.lane(); j < 16; j += G::kSize)
        if (j != i) { c += fabs(SIX_T(j, i)); r += fabs(SIX_T(i, j)); }
      r = g.sum(r);
      c = g.sum(c);
      if (c != 0.0 && r != 0.0) {
        double gg = r * 0.5, f = 1.0;
        const double s = c + r;
        while (c < gg) { f *= 2.0; c *= 4.0; }
        gg = r * 2.0;
        while (c > gg) { f *= 0.5; c *= 0.25; }
        if ((c + r) / f < 0.95 * s) {
          done = false;
          gg = 1.0 / f;
          g.sync();
          for (int j = g.lane(); j < 16; j += G::kSize) SIX_T(i, j) *= gg;
          g.sync();
          for (int j = g.lane(); j < 16; j += G::kSize) SIX_T(j, i) *= f;
          g.sync();
        }
      }
    }
    if (done) break;
  }
}

// Reduction to upper Hessenberg form by stabilised elementary similarity transformations (sixpt::to_hessenberg).  Step m:
// A <- L^-1 A L with L = I + y e_m^T, y_r = a(r, m-1) / a(m, m-1): first column m += sum_r y_r column r, then row r -= y_r row m.
template <class G>
SSFM_HD void six_hessenberg(const G& g, double* S) {
  for (int m = 1; m < 15; ++m) {
    double x = 0.0;
    int piv = m;
    for (int j = m; j < 16; ++j) {
      const double v = SIX_T(j, m - 1);
      if (fabs(v) > fabs(x)) { x = v; piv = j; }
    }
    g.sync();
    if (piv != m) {
      for (int j = m - 1 + g.lane(); j < 16; j += G::kSize) { const double t = SIX_T(piv, j); SIX_T(piv, j) = SIX_T(m, j); SIX_T(m, j) = t; }
      g.sync();
      for (int j = g.lane(); j < 16; j += G::kSize) { const double t = SIX_T(j, piv); SIX_T(j, piv) = SIX_T(j, m); SIX_T(j, m) = t; }
      g.sync();
    }
    if (x != 0.0) {
      const double ix = 1.0 / x;
      for (int i = g.lane(); i < 16; i += G::kSize) {  // column m (reads columns m+1.. and column m-1, writes column m)
        double acc = SIX_T(i, m);
        for (int r = m + 1; r < 16; ++r) acc += (SIX_T(r, m - 1) * ix) * SIX_T(i, r);
        SIX_T(i, m) = acc;
      }
      g.sync();
      const int ncol = 16 - m, total = (15 - m) * ncol;
      for (int idx = g.lane(); idx < total; idx += G::kSize) {  // rows m+1..15, columns m..15
        const int r = m + 1 + idx / ncol, j = m + idx % ncol;
        SIX_T(r, j) -= (SIX_T(r, m - 1) * ix) * SIX_T(m, j);
      }
      g.sync();
      for (int r = m + 1 + g.lane(); r < 16; r += G::kSize) SIX_T(r, m - 1) = 0.0;
      g.sync();
    }
  }
}

// Eigenvalues of the upper Hessenberg T by the Francis double-shift QR iteration (sixpt::hessenberg_eigenvalues): the
// three-row and three-column updates of every bulge-chasing step are spread over the lanes; everything else is scalar work
// that each lane repeats.  wr / wi go to the scratch.  Returns false if an eigenvalue failed to converge.
template <class G>
SSFM_HD bool six_hqr(const G& g, double* S) {
  using sixpt::sign_of;
  double* wr = S + kOffWr;
  double* wi = S + kOffWi;
  double anorm = 0.0;
  for (int i = 0; i < 16; ++i)
    for (int j = (i > 0 ? i - 1 : 0); j < 16; ++j) anorm += fabs(SIX_T(i, j));
  int nn = 15;
  double t = 0.0;
  double p = 0.0, q = 0.0, r = 0.0;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        double s = fabs(SIX_T(l - 1, l - 1)) + fabs(SIX_T(l, l));
        if (s == 0.0) s = anorm;
        if (fabs(SIX_T(l, l - 1)) + s == s) break;
      }
      g.sync();
      if (l >= 1 && g.lane() == 0) SIX_T(l, l - 1) = 0.0;
      double x = SIX_T(nn, nn);
      if (l == nn) {  // one real root
        if (g.lane() == 0) { wr[nn] = x + t; wi[nn] = 0.0; }
        --nn;
      } else {
        double y = SIX_T(nn - 1, nn - 1);
        double w = SIX_T(nn, nn - 1) * SIX_T(nn - 1, nn);
        if (l == nn - 1) {  // a 2x2 block: two roots
          p = 0.5 * (y - x);
          q = p * p + w;
          double z = sqrt(fabs(q));
          x += t;
          if (g.lane() == 0) {
            if (q >= 0.0) {
              z = p + sign_of(z, p);
              wr[nn - 1] = wr[nn] = x + z;
              if (z != 0.0) wr[nn] = x - w / z;
              wi[nn - 1] = wi[nn] = 0.0;
            } else {
              wr[nn - 1] = wr[nn] = x + p;
              wi[nn - 1] = z;
              wi[nn] = -z;
            }
          }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20 || its == 30 || its == 40) {  // exceptional shift
            t += x;
            const double s = fabs(SIX_T(nn, nn - 1)) + fabs(SIX_T(nn - 1, nn - 2));
            g.sync();
            for (int i = g.lane(); i <= nn; i += G::kSize) SIX_T(i, i) -= x;
            g.sync();
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          double z;
          for (m = nn - 2; m >= l; --m) {  // look for two consecutive small sub-diagonal elements
            z = SIX_T(m, m);
            r = x - z;
            double s = y - z;
            p = (r * s - w) * fast_rcp(SIX_T(m + 1, m)) + SIX_T(m, m + 1);
            q = SIX_T(m + 1, m + 1) - z - r - s;
            r = SIX_T(m + 2, m + 1);
            if (m == l) break;
            // the test is homogeneous in (p, q, r): their common scale is taken out once, after the loop
            const double u = fabs(SIX_T(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(SIX_T(m - 1, m - 1)) + fabs(z) + fabs(SIX_T(m + 1, m + 1)));
            if (u + v == v) break;
          }
          {
            const double s = fast_rcp(fabs(p) + fabs(q) + fabs(r));
            p *= s; q *= s; r *= s;
          }
          g.sync();
          for (int i = m + 2 + g.lane(); i <= nn; i += G::kSize) {
            SIX_T(i, i - 2) = 0.0;
            if (i != m + 2) SIX_T(i, i - 3) = 0.0;
          }
          g.sync();
          for (int k = m; k <= nn - 1; ++k) {  // double QR step on rows l..nn, columns m..nn
            if (k != m) {
              p = SIX_T(k, k - 1);
              q = SIX_T(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = SIX_T(k + 2, k - 1);
              x = fabs(p) + fabs(q) + fabs(r);
              if (x != 0.0) { const double ix = fast_rcp(x); p *= ix; q *= ix; r *= ix; }
            }
            const double n2 = p * p + q * q + r * r;
            if (n2 != 0.0) {
              const double is = sign_of(fast_rsqrt(n2), p);  // 1 / s, s = sign(p) |(p, q, r)|
              const double s = n2 * is;
              const double sub = (k == m) ? ((l != m) ? -SIX_T(k, k - 1) : 0.0) : -s * x;
              const bool write_sub = (k != m) || (l != m);
              p += s;
              x = p * is;
              y = q * is;
              z = r * is;
              const double ip = fast_rcp(p);
              q *= ip;
              r *= ip;
              const bool three = k != nn - 1;
              g.sync();  // every lane has read column k-1
              if (write_sub && g.lane() == 0) SIX_T(k, k - 1) = sub;
              for (int j = k + g.lane(); j <= nn; j += G::kSize) {
                const double a0 = SIX_T(k, j), a1 = SIX_T(k + 1, j);
                double pp = a0 + q * a1;
                if (three) {
                  const double a2 = SIX_T(k + 2, j);
                  pp += r * a2;
                  SIX_T(k + 2, j) = a2 - pp * z;
                }
                SIX_T(k + 1, j) = a1 - pp * y;
                SIX_T(k, j) = a0 - pp * x;
              }
              g.sync();
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l + g.lane(); i <= mmin; i += G::kSize) {
                const double a0 = SIX_T(i, k), a1 = SIX_T(i, k + 1);
                double pp = x * a0 + y * a1;
                if (three) {
                  const double a2 = SIX_T(i, k + 2);
                  pp += z * a2;
                  SIX_T(i, k + 2) = a2 - pp * r;
                }
                SIX_T(i, k + 1) = a1 - pp * q;
                SIX_T(i, k) = a0 - pp;
              }
              g.sync();
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  g.sync();
  return true;
}

// ----------------------------------------------------------------------------------------------------------------
// Candidates, polish, decomposition: per-lane code (one candidate / one solution per call)
// ----------------------------------------------------------------------------------------------------------------
// Is eigenvalue k (mu = f^2 in the normalised units) nearly real and positive?  Then *w = the start value of 1 / f^2.
SSFM_HD bool six_candidate(double mu, double wik, double* w) {
  if (!(mu > 0.0) || !(mu < 1e300) || fabs(wik) > 1e-4 * mu) return false;  // nearly real: the polish + residual test decide
  // A nearly-real conjugate pair may be two close real roots: start the polish on either side.
  *w = 1.0 / (mu + wik);
  if (!(*w > 0.0)) *w = 1.0 / mu;
  return true;
}

// min || A x - b || for a 10 x COLS tableau in the group's scratch (rows 10 doubles apart: A[e][q] at tab[10 * e + q],
// b[e] at tab[10 * e + COLS]), by Householder QR with the columns spread over the lanes (sixpt::least_squares10).  On return
// the tableau holds R and Q^T b; the caller back-substitutes what it needs.  Returns false (group-uniform) on a zero column.
template <int COLS, class G>
SSFM_HD bool six_ls_reduce(const G& g, double* tab) {
  for (int k = 0; k < COLS; ++k) {
    double nrm = 0.0;
    for (int i = k; i < 10; ++i) nrm += tab[10 * i + k] * tab[10 * i + k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) return false;
    const double akk = tab[10 * k + k];
    const double alpha = akk > 0 ? -nrm : nrm;
    const double v0 = akk - alpha;
    double vn = v0 * v0;
    for (int i = k + 1; i < 10; ++i) vn += tab[10 * i + k] * tab[10 * i + k];
    if (vn > 0.0) {
      const double beta = 2.0 / vn;
      for (int j = k + 1 + g.lane(); j <= COLS; j += G::kSize) {  // column COLS is the right-hand side
        double d = v0 * tab[10 * k + j];
        for (int i = k + 1; i < 10; ++i) d += tab[10 * i + k] * tab[10 * i + j];
        d *= beta;
        tab[10 * k + j] -= d * v0;
        for (int i = k + 1; i < 10; ++i) tab[10 * i + j] -= d * tab[10 * i + k];
      }
    }
    g.sync();
    if (g.lane() == 0) tab[10 * k + k] = alpha;
  }
  g.sync();
  return true;
}

// (x, y) at a given w: least squares for the nine non-constant monomials of the ten equations, by the group of the sample.
// M: the sample's M[3][10][10]; tab: 100 doubles of the group's scratch.  Group-uniform.
template <class G>
SSFM_HD bool six_start(const G& g, const double* M, double* tab, double w, double* xy) {
  g.sync();
  for (int idx = g.lane(); idx < 100; idx += G::kSize) {
    const int e = idx / 10, q = idx - 10 * e;
    const double v = M[e * 10 + q] + w * (M[100 + e * 10 + q] + w * M[200 + e * 10 + q]);
    tab[idx] = q < 9 ? v : -v;
  }
  g.sync();
  if (!six_ls_reduce<9>(g, tab)) return false;
  xy[1] = tab[10 * 8 + 9] / tab[10 * 8 + 8];
  xy[0] = (tab[10 * 7 + 9] - tab[10 * 7 + 8] * xy[1]) / tab[10 * 7 + 7];
  return true;
}

// Gauss-Newton on the ten equations in (x, y, w) + residual test: ONE LANE per candidate, everything in registers.  The
// 10 x 3 least-squares problem of a step is reduced row by row with Givens rotations as the rows are produced (same
// stability as the Householder form, no tableau to store).  sol3 = (x, y, w) start in, solution out.
SSFM_HD_NOINLINE bool six_newton(const double* M, double* sol3) {
  using namespace sixpt;
  double x = sol3[0], y = sol3[1], w = sol3[2];
  bool ok = true, converged = false;
  for (int itn = 0; itn < 12 && ok && !converged; ++itn) {
    double m[10], dx[10], dy[10];
    monomials(x, y, m, dx, dy);
    double r00 = 0, r01 = 0, r02 = 0, r11 = 0, r12 = 0, r22 = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int e = 0; e < 10; ++e) {
      double r0 = 0, a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const double m1 = M[100 + e * 10 + q], m2 = M[200 + e * 10 + q];
        const double mw = M[e * 10 + q] + w * (m1 + w * m2);
        r0 += mw * m[q];
        a0 += mw * dx[q];
        a1 += mw * dy[q];
        a2 += (m1 + 2.0 * w * m2) * m[q];
      }
      double rb = -r0;
      // rotate the row (a0, a1, a2 | rb) into the triangle
      double h = r00 * r00 + a0 * a0;
      if (h > 0.0) {
        const double ih = fast_rsqrt(h), c = r00 * ih, sn = a0 * ih;
        r00 = h * ih;
        double t = c * r01 + sn * a1; a1 = c * a1 - sn * r01; r01 = t;
        t = c * r02 + sn * a2; a2 = c * a2 - sn * r02; r02 = t;
        t = c * b0 + sn * rb; rb = c * rb - sn * b0; b0 = t;
      }
      h = r11 * r11 + a1 * a1;
      if (h > 0.0) {
        const double ih = fast_rsqrt(h), c = r11 * ih, sn = a1 * ih;
        r11 = h * ih;
        double t = c * r12 + sn * a2; a2 = c * a2 - sn * r12; r12 = t;
        t = c * b1 + sn * rb; rb = c * rb - sn * b1; b1 = t;
      }
      h = r22 * r22 + a2 * a2;
      if (h > 0.0) {
        const double ih = fast_rsqrt(h), c = r22 * ih, sn = a2 * ih;
        r22 = h * ih;
        b2 = c * b2 + sn * rb;
      }
    }
    if (!(r00 > 0.0) || !(r11 > 0.0) || !(r22 > 0.0)) { ok = false; break; }
    double step[3];
    step[2] = b2 / r22;
    step[1] = (b1 - r12 * step[2]) / r11;
    step[0] = (b0 - r01 * step[1] - r02 * step[2]) / r00;
    x += step[0]; y += step[1]; w += step[2];
    if (!(fabs(x) < 1e300) || !(fabs(y) < 1e300) || !(fabs(w) < 1e300)) ok = false;
    // accepted only once the iteration has settled (a start that wanders is a spurious eigenvalue)
    converged = fabs(step[0]) <= 1e-12 * (1.0 + fabs(x)) && fabs(step[1]) <= 1e-12 * (1.0 + fabs(y)) &&
                fabs(step[2]) <= 1e-12 * fabs(w);
  }
  ok = ok && converged;
  if (!ok || !(w > 0.0)) return false;
  {  // residual test against the scale of each equation
    double m[10], dx[10], dy[10];
    monomials(x, y, m, dx, dy);
    for (int e = 0; e < 10 && ok; ++e) {
      double r0 = 0, sc = 0;
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const double mw = M[e * 10 + q] + w * (M[100 + e * 10 + q] + w * M[200 + e * 10 + q]);
        r0 += mw * m[q];
        sc += fabs(mw) * fabs(m[q]);
      }
      if (fabs(r0) > 1e-8 * sc) ok = false;
    }
  }
  if (!ok) return false;
  sol3[0] = x; sol3[1] = y; sol3[2] = w;
  return true;
}

SSFM_HD bool six_same_solution(const double* a, const double* b) {  // b = the later one
  return fabs(b[0] - a[0]) + fabs(b[1] - a[1]) < 1e-6 * (1.0 + fabs(b[0]) + fabs(b[1])) && fabs(b[2] - a[2]) < 1e-6 * b[2];
}

// One solution (x, y, w) -> the (R, t) in front of both cameras (at most 4), in the order the per-thread solver emits them.
SSFM_HD_NOINLINE int six_decompose(const double* S, const double* sol3, SixPointModel* out4) {
  using namespace sixpt;
  const double s = S[kOffMisc];
  const double* Fb = S + kOffFb;
  const double* x1 = S + kOffPts;
  const double* x2 = S + kOffPts + 18;
  const double f = 1.0 / sqrt(sol3[2]);
  double E[9];
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < 3; ++cc) {
      const double Fv = sol3[0] * Fb[3 * r + cc] + sol3[1] * Fb[9 + 3 * r + cc] + Fb[18 + 3 * r + cc];
      E[3 * r + cc] = Fv * (r < 2 ? f : 1.0) * (cc < 2 ? f : 1.0);
    }
  double U[9], sv[3], V[9];
  svd3(E, U, sv, V);
  if (mat3_det(U) < 0) { U[2] = -U[2]; U[5] = -U[5]; U[8] = -U[8]; }
  if (mat3_det(V) < 0) { V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8]; }
  double b1[6][3], b2[6][3];  // unit bearings of the sample in both cameras
  for (int i = 0; i < 6; ++i) {
    double a[3] = {x1[3 * i] / f, x1[3 * i + 1] / f, x1[3 * i + 2]}, b[3] = {x2[3 * i] / f, x2[3 * i + 1] / f, x2[3 * i + 2]};
    const double na = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = 1.0 / sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    for (int d = 0; d < 3; ++d) { b1[i][d] = a[d] * na; b2[i][d] = b[d] * nb; }
  }
  const double Wm[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  double Vt[9];
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < 3; ++cc) Vt[3 * r + cc] = V[3 * cc + r];
  int n = 0;
  for (int which = 0; which < 2; ++which) {
    double tmp[9], R[9];
    mat3_mul(U, which == 0 ? Wm : Wt, tmp);
    mat3_mul(tmp, Vt, R);
    for (int sg = 0; sg < 2; ++sg) {
      const double t[3] = {sg == 0 ? U[2] : -U[2], sg == 0 ? U[5] : -U[5], sg == 0 ? U[8] : -U[8]};
      bool front = true;
      for (int i = 0; i < 6 && front; ++i) front = in_front(R, t, b1[i], b2[i]);
      if (!front) continue;
      SixPointModel& mdl = out4[n++];
      for (int d = 0; d < 3; ++d) mdl.t[d] = t[d];
      so3ln(R, mdl.r);
      mdl.f = f * s;
    }
  }
  return n;
}

// Insert into a list sorted by focal (stable: equal focals keep their arrival order); returns the new count (<= 15).
SSFM_HD int six_insert_sorted(SixPointModel* out, int n_out, const SixPointModel& mdl) {
  if (n_out >= kSixMaxModels) return n_out;
  int pos = n_out;
  while (pos > 0 && out[pos - 1].f > mdl.f) { out[pos] = out[pos - 1]; --pos; }
  out[pos] = mdl;
  return n_out + 1;
}

// The whole solver with a one-lane group: what tests/hostshim runs on the CPU, and the reference order of operations of
// the warp kernel (k_sixpt_sample_solve walks the same stages with LaneGroup8 and one lane per candidate).
SSFM_HD_NOINLINE int solve_sixpt_focal_staged(const double (*c)[6], SixPointModel* out) {
  SerialGroup g;
  double S[kScratch], M[kMSize];
  if (!six_setup(g, c, S, M)) return 0;
  if (!six_companion(g, S, M)) return 0;
  six_balance(g, S);
  six_hessenberg(g, S);
  if (!six_hqr(g, S)) return 0;
  double sols[kSixMaxModels][3];
  int n_sol = 0, n_out = 0;
  for (int k = 0; k < 16; ++k) {
    double w, sol[3];
    if (!six_candidate(S[kOffWr + k], S[kOffWi + k], &w)) continue;
    if (!six_start(g, M, S + kOffT, w, sol)) continue;
    sol[2] = w;
    if (!six_newton(M, sol)) continue;
    bool dup = false;
    for (int q = 0; q < n_sol; ++q) dup = dup || six_same_solution(sols[q], sol);
    if (dup || n_sol >= kSixMaxModels) continue;
    for (int d = 0; d < 3; ++d) sols[n_sol][d] = sol[d];
    ++n_sol;
  }
  for (int q = 0; q < n_sol; ++q) {
    SixPointModel four[4];
    const int n4 = six_decompose(S, sols[q], four);
    for (int i = 0; i < n4; ++i) n_out = six_insert_sorted(out, n_out, four[i]);
  }
  return n_out;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------------------------
// The group driver: 8 lanes solve one sample from start to end (4 samples per warp, independent of each other).
// ----------------------------------------------------------------------------------------------------------------
constexpr int kSixSamplesPerWarp = 4;

// S: the group's scratch (kScratch doubles, shared memory); Mg: the sample's global M slot; c: the six correspondences
// (read by lane 0 of the group only).  Returns the number of models (group-uniform); they are left sorted by focal at
// S + kOffT (SixPointModel records).  Every warp-level collective inside names the group's own lanes.
// kBlockSync: every warp of the block walks the stages together (__syncthreads between them), so that all warps of an SM
// execute the same few KB of code at a time: the solver is ~200 KB of instructions, and with every block in a stage of
// its own the kernel spent more than half of its cycles waiting for instruction fetches (ncu: stall_no_instruction 7.8
// per issue).  A group without a sample (valid == false) only takes part in the block barriers.
template <bool kBlockSync>
__device__ __forceinline__ int six_solve_group(double* S, double* Mg, const double (*c)[6], bool valid) {
  const int lane = threadIdx.x & 31, grp = lane >> 3;
  const LaneGroup8 g{lane & 7, 0xFFu << (8 * grp)};
  bool live = valid;
  if (live) live = six_setup(g, c, S, Mg) && six_companion(g, S, Mg);
  if (live) {
    six_balance(g, S);
    six_hessenberg(g, S);
  }
  if (kBlockSync) __syncthreads();
  if (live) live = six_hqr(g, S);
  if (kBlockSync) __syncthreads();
  if (live) {
    // candidates in eigenvalue order: least-squares start by the whole group, one after the other ...
    for (int k = 0; k < 16; ++k) {
      double w, xy[2];
      bool keep = six_candidate(S[kOffWr + k], S[kOffWi + k], &w);
      if (keep) keep = six_start(g, Mg, S + kOffT, w, xy);
      g.sync();
      if (g.l == 0) {
        double* dst = S + kOffCand + 4 * k;
        dst[3] = keep ? 1.0 : 0.0;
        if (keep) { dst[0] = xy[0]; dst[1] = xy[1]; dst[2] = w; }
      }
    }
    g.sync();
  }
  if (kBlockSync) __syncthreads();
  if (live) {
    // ... Gauss-Newton with one lane per candidate ...
    for (int r = 0; r < 2; ++r) {
      double* cand = S + kOffCand + 4 * (g.l + 8 * r);
      if (cand[3] == 1.0) {
        double sol[3] = {cand[0], cand[1], cand[2]};
        if (six_newton(Mg, sol)) { cand[0] = sol[0]; cand[1] = sol[1]; cand[2] = sol[2]; }
        else cand[3] = 0.0;
      }
    }
    g.sync();
  }
  if (kBlockSync) __syncthreads();
  if (!live) return 0;
  if (g.l == 0) {  // ... and a solution equal to an earlier accepted one is dropped
    int n_sol = 0;
    for (int k = 0; k < 16; ++k) {
      if (S[kOffCand + 4 * k + 3] != 1.0) continue;
      bool dup = false;
      for (int q = 0; q < k; ++q)
        if (S[kOffCand + 4 * q + 3] == 2.0) dup = dup || six_same_solution(S + kOffCand + 4 * q, S + kOffCand + 4 * k);
      if (dup || n_sol >= kSixMaxModels) continue;
      S[kOffCand + 4 * k + 3] = 2.0;
      ++n_sol;
    }
    S[kOffMisc + kMiscStride] = 0.0;  // number of models in the list
  }
  g.sync();
  // decomposition: one lane per accepted solution; the models enter the focal-sorted list in eigenvalue order
  SixPointModel* list = reinterpret_cast<SixPointModel*>(S + kOffList);
  for (int r = 0; r < 2; ++r) {
    const int k = g.l + 8 * r;
    SixPointModel four[4];
    int n4 = 0;
    if (S[kOffCand + 4 * k + 3] == 2.0) n4 = six_decompose(S, S + kOffCand + 4 * k, four);
    for (int kk = 0; kk < 8; ++kk) {
      if (g.l == kk && n4 > 0) {
        int n_out = (int)S[kOffMisc + kMiscStride];
        for (int i = 0; i < n4; ++i) n_out = six_insert_sorted(list, n_out, four[i]);
        S[kOffMisc + kMiscStride] = (double)n_out;
      }
      g.sync();
    }
  }
  return (int)S[kOffMisc + kMiscStride];
}
#endif  // __CUDACC__

}  // namespace sixc
}  // namespace ssfm
