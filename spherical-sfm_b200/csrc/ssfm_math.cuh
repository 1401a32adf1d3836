// ssfm_math.cuh -- per-thread float64 building blocks of the B200 relative-pose engine.
//
// Everything here is __host__ __device__ and free of warp intrinsics, so the same source is
// compiled by nvcc into the kernels (ssfm_kernels.cu) and by g++ into tests/hostshim (a TEST-ONLY
// build used to unit-test the device arithmetic on machines without a GPU; it is never linked
// into libssfm_b200.so and is not a CPU fallback).
//
// Reference citations are relative to the reference root.  The numerics are deliberately NOT the
// reference's (nor the oracle's): null space by pivoted Householder, but roots by a real
// quadratic factorisation of the characteristic quartic + Bairstow/Newton polish and eigenvectors
// by 2x2 solves, so that parity tests compare two independent implementations.  What is returned
// is the reference's: four models per sample, the polynomial variant's Re(y) models included.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SSFM_HD __host__ __device__ __forceinline__
#define SSFM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SSFM_HD inline
#define SSFM_HD_NOINLINE inline
#endif

namespace ssfm {

constexpr double kDblMax = 1.7976931348623157e308;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 sampler: replaces UniformSampling (include/RansacLib/sampling.h:47-135).
// Sample of iteration `iter` of pair `pair` = pure function of (seed, pair, iter).
// ---------------------------------------------------------------------------------------------
SSFM_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// k distinct indices in [0, n): draw-with-rejection like DrawSample (sampling.h:81-97).
template <int KMAX>
SSFM_HD void philox_sample(uint32_t seed, uint32_t pair, uint32_t iter, int k, int n, int* idx) {
  uint32_t w[4];
  uint32_t block = 0;
  int used = 4;
  for (int i = 0; i < k && i < KMAX; ++i) {
    bool dup = true;
    while (dup) {
      if (used == 4) {
        philox4x32_10(iter, block++, 0u, 0u, seed, pair, w);
        used = 0;
      }
      const uint32_t x = used == 0 ? w[0] : (used == 1 ? w[1] : (used == 2 ? w[2] : w[3]));
      ++used;
      const int cand = (int)(((uint64_t)x * (uint64_t)(uint32_t)n) >> 32);
      dup = false;
      for (int j = 0; j < i; ++j) dup = dup || (idx[j] == cand);
      idx[i] = cand;
    }
  }
}

// rand() of the legacy drivers as a pure function of (seed, pair, hypothesis, draw): word (draw % 4) of
// Philox(counter = (hyp, draw / 4, 1, 0), key = (seed, pair)) >> 1, RAND_MAX folded onto RAND_MAX - 1 so that
// u = rand() / RAND_MAX < 1 (with u == 1 upstream's selection sampling can run past the end of the list).
struct PhiloxRand31 {
  uint32_t seed, pair, hyp, draw;
  uint32_t w[4];
  SSFM_HD int next() {
    if ((draw & 3u) == 0u) philox4x32_10(hyp, draw >> 2, 1u, 0u, seed, pair, w);
    const uint32_t k = draw & 3u;
    uint32_t r = (k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]))) >> 1;
    ++draw;
    if (r == 0x7fffffffu) r = 0x7ffffffeu;
    return (int)r;
  }
};

// random_sample (preemptive_ransac.h:8-28): n of N records, in increasing order.
SSFM_HD void knuth_sample(uint32_t seed, uint32_t pair, uint32_t hyp, int N, int n, int* out) {
  PhiloxRand31 g;
  g.seed = seed; g.pair = pair; g.hyp = hyp; g.draw = 0;
  int t = 0, m = 0;
  while (m < n) {
    const double u = (double)g.next() / 2147483647.0;
    if ((double)(N - t) * u >= (double)(n - m)) {
      t++;
    } else {
      out[m] = t;
      t++;
      m++;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Exactly-rounded float64 helpers: the certification stage must reproduce the reference's
// (FMA-free x86) arithmetic bit for bit, so contraction is switched off explicitly.
// ---------------------------------------------------------------------------------------------
SSFM_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
SSFM_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
SSFM_HD double dot3_rn(double a0, double b0, double a1, double b1, double a2, double b2) {
  return add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2));
}

// EvaluateModelOnPoint: squared Sampson distance (src/spherical_estimator.cpp:67-78), float64,
// same operation order as the reference (Eigen row.col products, left to right).
SSFM_HD double sampson_exact(const double* E, const double* u, const double* v) {
  const double Eu0 = dot3_rn(E[0], u[0], E[1], u[1], E[2], u[2]);
  const double Eu1 = dot3_rn(E[3], u[0], E[4], u[1], E[5], u[2]);
  const double Eu2 = dot3_rn(E[6], u[0], E[7], u[1], E[8], u[2]);
  const double Etv0 = dot3_rn(E[0], v[0], E[3], v[1], E[6], v[2]);
  const double Etv1 = dot3_rn(E[1], v[0], E[4], v[1], E[7], v[2]);
  const double d = dot3_rn(v[0], Eu0, v[1], Eu1, v[2], Eu2);
  const double den = add_rn(add_rn(mul_rn(Eu0, Eu0), mul_rn(Eu1, Eu1)), add_rn(mul_rn(Etv0, Etv0), mul_rn(Etv1, Etv1)));
  return mul_rn(d, d) / den;
}

SSFM_HD void E_from_p(const double* p, double* E) {  // src/spherical_solvers.cpp:299-303
  E[0] = p[0]; E[1] = p[1];  E[2] = p[2];
  E[3] = p[1]; E[4] = -p[0]; E[5] = p[3];
  E[6] = p[4]; E[7] = p[5];  E[8] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// 3x3 helpers, so(3), spherical essential matrix (src/so3.cpp:6-70, src/spherical_utils.cpp:9-66)
// ---------------------------------------------------------------------------------------------
SSFM_HD void mat3_mul(const double* a, const double* b, double* r) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
SSFM_HD double mat3_det(const double* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
SSFM_HD void so3exp(const double* r, double* R) {  // so3.cpp:16-23
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  if (theta < 1e-10) return;
  const double k0 = r[0] / theta, k1 = r[1] / theta, k2 = r[2] / theta;
  const double K[9] = {0, -k2, k1, k2, 0, -k0, -k1, k0, 0};
  double KK[9];
  mat3_mul(K, K, KK);
  const double s = sin(theta), c = 1.0 - cos(theta);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] += s * K[i] + c * KK[i];
}
SSFM_HD void so3ln(const double* R, double* res) {  // so3.cpp:25-69
  const double cos_angle = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  res[0] = (R[7] - R[5]) / 2;
  res[1] = (R[2] - R[6]) / 2;
  res[2] = (R[3] - R[1]) / 2;
  const double sin_abs = sqrt(res[0] * res[0] + res[1] * res[1] + res[2] * res[2]);
  const double kSqrt1_2 = 0.70710678118654752440;
  if (cos_angle > kSqrt1_2) {
    if (sin_abs > 0) {
      const double f = asin(sin_abs) / sin_abs;
      res[0] *= f; res[1] *= f; res[2] *= f;
    }
  } else if (cos_angle > -kSqrt1_2) {
    const double f = acos(cos_angle) / sin_abs;
    res[0] *= f; res[1] *= f; res[2] *= f;
  } else {
    const double angle = 3.14159265358979323846 - asin(sin_abs);
    const double d0 = R[0] - cos_angle, d1 = R[4] - cos_angle, d2 = R[8] - cos_angle;
    double r2[3];
    if (fabs(d0) > fabs(d1) && fabs(d0) > fabs(d2)) {
      r2[0] = d0; r2[1] = (R[3] + R[1]) / 2; r2[2] = (R[2] + R[6]) / 2;
    } else if (fabs(d1) > fabs(d2)) {
      r2[0] = (R[3] + R[1]) / 2; r2[1] = d1; r2[2] = (R[7] + R[5]) / 2;
    } else {
      r2[0] = (R[2] + R[6]) / 2; r2[1] = (R[7] + R[5]) / 2; r2[2] = d2;
    }
    if (r2[0] * res[0] + r2[1] * res[1] + r2[2] * res[2] < 0) { r2[0] = -r2[0]; r2[1] = -r2[1]; r2[2] = -r2[2]; }
    const double n = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    res[0] = angle * r2[0] / n; res[1] = angle * r2[1] / n; res[2] = angle * r2[2] / n;
  }
}
SSFM_HD void make_spherical_E(const double* R, bool inward, double* E) {  // spherical_utils.cpp:9-14
  double t0 = R[2], t1 = R[5], t2 = R[8] - 1;
  if (inward) { t0 = -t0; t1 = -t1; t2 = -t2; }
  const double S[9] = {0, -t2, t1, t2, 0, -t0, -t1, t0, 0};
  mat3_mul(S, R, E);
}

// 3x3 SVD via the symmetric eigen-decomposition of A^T A by cyclic Jacobi rotations, with
// U = A V / s and completion of the null direction by a cross product.  (Role of Eigen::JacobiSVD,
// spherical_utils.cpp:18.)  Singular values sorted descending.
SSFM_HD_NOINLINE void svd3(const double* A, double* U, double* s, double* V) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) a[i][j] = A[3 * i + j];
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0;
      const int q = pq == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        alpha += a[i][p] * a[i][p];
        beta += a[i][q] * a[i][q];
        gamma += a[i][p] * a[i][q];
      }
      if (gamma != 0.0) {
        const double ab = alpha * beta;
        off = fmax(off, fabs(gamma) / sqrt(ab > 1e-300 ? ab : 1e-300));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double ap = a[i][p], aq = a[i][q];
          a[i][p] = c * ap - sn * aq;
          a[i][q] = sn * ap + c * aq;
          const double vp = v[i][p], vq = v[i][q];
          v[i][p] = c * vp - sn * vq;
          v[i][q] = sn * vp + c * vq;
        }
      }
    }
    if (off < 1e-15) break;
  }
  double sv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) sv[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
  // sorting network on (sv, columns)
#define SSFM_SWAPCOL(x, y)                                             \
  if (sv[x] < sv[y]) {                                                 \
    double tmp = sv[x]; sv[x] = sv[y]; sv[y] = tmp;                    \
    for (int i = 0; i < 3; ++i) {                                      \
      tmp = a[i][x]; a[i][x] = a[i][y]; a[i][y] = tmp;                 \
      tmp = v[i][x]; v[i][x] = v[i][y]; v[i][y] = tmp;                 \
    }                                                                  \
  }
  SSFM_SWAPCOL(0, 1)
  SSFM_SWAPCOL(1, 2)
  SSFM_SWAPCOL(0, 1)
#undef SSFM_SWAPCOL
  double u[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s[k] = sv[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      V[3 * i + k] = v[i][k];
      u[i][k] = sv[k] > 0 ? a[i][k] / sv[k] : 0.0;
    }
  }
  if (s[2] <= 1e-14 * s[0]) {
    if (s[1] <= 1e-14 * s[0]) {
      double e[3] = {0, 0, 0};
      int mn = 0;
      if (fabs(u[1][0]) < fabs(u[mn][0])) mn = 1;
      if (fabs(u[2][0]) < fabs(u[mn][0])) mn = 2;
      e[0] = mn == 0; e[1] = mn == 1; e[2] = mn == 2;
      const double w0 = u[1][0] * e[2] - u[2][0] * e[1], w1 = u[2][0] * e[0] - u[0][0] * e[2],
                   w2 = u[0][0] * e[1] - u[1][0] * e[0];
      const double n = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
      u[0][1] = w0 / n; u[1][1] = w1 / n; u[2][1] = w2 / n;
    }
    u[0][2] = u[1][0] * u[2][1] - u[2][0] * u[1][1];
    u[1][2] = u[2][0] * u[0][1] - u[0][0] * u[2][1];
    u[2][2] = u[0][0] * u[1][1] - u[1][0] * u[0][1];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) U[3 * i + k] = u[i][k];
}

// decompose_spherical_essential_matrix (src/spherical_utils.cpp:16-66)
SSFM_HD_NOINLINE void decompose_spherical_E(const double* E, bool inward, double* r, double* t) {
  double U[9], V[9], s[3];
  svd3(E, U, s, V);
  if (mat3_det(U) < 0)
    for (int i = 0; i < 9; ++i) U[i] = -U[i];
  if (mat3_det(V) < 0)
    for (int i = 0; i < 9; ++i) V[i] = -V[i];
  const double D[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  const double DT[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double VT[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) VT[3 * i + j] = V[3 * j + i];
  double UD[9], R1[9], R2[9];
  mat3_mul(U, D, UD);
  mat3_mul(UD, VT, R1);
  mat3_mul(U, DT, UD);
  mat3_mul(UD, VT, R2);
  double t1[3] = {R1[2], R1[5], R1[8] - 1}, t2[3] = {R2[2], R2[5], R2[8] - 1};
  if (inward)
    for (int i = 0; i < 3; ++i) { t1[i] = -t1[i]; t2[i] = -t2[i]; }
  const double n1 = sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
  const double n2 = sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
  const double tu[3] = {U[2], U[5], U[8]};
  const double score1 = fabs((t1[0] * tu[0] + t1[1] * tu[1] + t1[2] * tu[2]) / n1);
  const double score2 = fabs((t2[0] * tu[0] + t2[1] * tu[1] + t2[2] * tu[2]) / n2);
  double r1[3], r2[3];
  so3ln(R1, r1);
  so3ln(R2, r2);
  if (score1 > score2) {
    for (int i = 0; i < 3; ++i) { r[i] = r1[i]; t[i] = t1[i]; }
  } else {
    for (int i = 0; i < 3; ++i) { r[i] = r2[i]; t[i] = t2[i]; }
  }
}

// ---------------------------------------------------------------------------------------------
// The 3-point solvers.
// ---------------------------------------------------------------------------------------------
struct Cplx {
  double re, im;
};
SSFM_HD Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
SSFM_HD Cplx cadd(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
SSFM_HD Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
SSFM_HD Cplx cscale(double s, Cplx a) { return {s * a.re, s * a.im}; }
SSFM_HD double cabs2(Cplx a) { return a.re * a.re + a.im * a.im; }
SSFM_HD Cplx cdiv(Cplx a, Cplx b) {
  const double d = 1.0 / cabs2(b);
  return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}

// Epipolar row of one correspondence (src/spherical_solvers.cpp:119).
SSFM_HD void epipolar_row(const double* u, const double* v, double* a) {
  a[0] = u[0] * v[0] - u[1] * v[1];
  a[1] = u[0] * v[1] + u[1] * v[0];
  a[2] = u[2] * v[0];
  a[3] = u[2] * v[1];
  a[4] = u[0] * v[2];
  a[5] = u[1] * v[2];
}

// B = last three columns of Q of the column-pivoted Householder QR of A^T (6 x N)
// (src/spherical_solvers.cpp:124-125; Eigen's makeHouseholder sign convention, pivot = largest
// remaining column norm).  N is a compile-time constant (3 for minimal samples).
template <int N>
SSFM_HD void nullspace_colpiv(double (&m)[6][N], double (&B)[6][3]) {
  constexpr int STEPS = N < 6 ? N : 6;
  double vs[STEPS][6];
  double taus[STEPS];
#pragma unroll
  for (int k = 0; k < STEPS; ++k) {
    int piv = k;
    double best = -1.0;
#pragma unroll
    for (int c = k; c < N; ++c) {
      double s = 0.0;
#pragma unroll
      for (int r = k; r < 6; ++r) s += m[r][c] * m[r][c];
      if (s > best) { best = s; piv = c; }
    }
#pragma unroll
    for (int c = k + 1; c < N; ++c)
      if (c == piv) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { const double tmp = m[r][k]; m[r][k] = m[r][c]; m[r][c] = tmp; }
      }
    const double c0 = m[k][k];
    double tail2 = 0.0;
#pragma unroll
    for (int r = k + 1; r < 6; ++r) tail2 += m[r][k] * m[r][k];
    double tau, beta;
#pragma unroll
    for (int r = 0; r < 6; ++r) vs[k][r] = 0.0;
    vs[k][k] = 1.0;
    if (tail2 <= 2.2250738585072014e-308) {
      tau = 0.0;
      beta = c0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      const double inv = 1.0 / (c0 - beta);
#pragma unroll
      for (int r = k + 1; r < 6; ++r) vs[k][r] = m[r][k] * inv;
      tau = (beta - c0) / beta;
    }
#pragma unroll
    for (int c = k; c < N; ++c) {
      double d = 0.0;
#pragma unroll
      for (int r = k; r < 6; ++r) d += vs[k][r] * m[r][c];
      d *= tau;
#pragma unroll
      for (int r = k; r < 6; ++r) m[r][c] -= d * vs[k][r];
    }
    taus[k] = tau;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double q[6] = {0, 0, 0, 0, 0, 0};
    q[3 + j] = 1.0;
#pragma unroll
    for (int k = STEPS - 1; k >= 0; --k) {
      double d = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) d += vs[k][r] * q[r];
      d *= taus[k];
#pragma unroll
      for (int r = 0; r < 6; ++r) q[r] -= d * vs[k][r];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) B[r][j] = q[r];
  }
}

// Polynomial arithmetic on forms in (x, y, z).
// quadratic order: xx xy xz yy yz zz ; cubic order: xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz
SSFM_HD void qmul(const double* l, const double* m, double* q) {
  q[0] = l[0] * m[0];
  q[1] = l[0] * m[1] + l[1] * m[0];
  q[2] = l[0] * m[2] + l[2] * m[0];
  q[3] = l[1] * m[1];
  q[4] = l[1] * m[2] + l[2] * m[1];
  q[5] = l[2] * m[2];
}
// r (+)= l * q
template <bool ACC>
SSFM_HD void cmul_lq(const double* l, const double* q, double* r) {
  const double lx = l[0], ly = l[1], lz = l[2];
  const double t0 = lx * q[0];
  const double t1 = lx * q[1] + ly * q[0];
  const double t2 = lx * q[2] + lz * q[0];
  const double t3 = lx * q[3] + ly * q[1];
  const double t4 = lx * q[4] + ly * q[2] + lz * q[1];
  const double t5 = lx * q[5] + lz * q[2];
  const double t6 = ly * q[3];
  const double t7 = ly * q[4] + lz * q[3];
  const double t8 = ly * q[5] + lz * q[4];
  const double t9 = lz * q[5];
  if (ACC) {
    r[0] += t0; r[1] += t1; r[2] += t2; r[3] += t3; r[4] += t4; r[5] += t5; r[6] += t6; r[7] += t7; r[8] += t8; r[9] += t9;
  } else {
    r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7; r[8] = t8; r[9] = t9;
  }
}

// Column order of the 6x10 system per solver kind, as indices into the canonical cubic order
// (src/spherical_solvers.cpp:271-279, :559-621; spherical_fast_estimator.cpp:207-215).
template <int KIND>
SSFM_HD int col_of(int c) {
  if (KIND == 0) { const int o[10] = {0, 1, 3, 6, 2, 4, 7, 5, 8, 9}; return o[c]; }
  if (KIND == 1) { const int o[10] = {0, 1, 3, 2, 4, 5, 6, 7, 8, 9}; return o[c]; }
  const int o[10] = {0, 1, 3, 6, 7, 8, 2, 4, 5, 9};
  return o[c];
}

// The six cubic constraints [T10, T20, T00, T21, T12, T22] of T = 2EE^T E - tr(EE^T)E on
// p = x B0 + y B1 + z B2 (what the hand-expanded block at src/spherical_solvers.cpp:127-277 encodes).
template <int KIND>
SSFM_HD void build_constraints(const double (&B)[6][3], double (&C)[6][10]) {
  double q22[6], q33[6], q44[6], q55[6], q23[6], q45[6], q24[6], q35[6], q25[6], q34[6];
  qmul(B[2], B[2], q22); qmul(B[3], B[3], q33); qmul(B[4], B[4], q44); qmul(B[5], B[5], q55);
  qmul(B[2], B[3], q23); qmul(B[4], B[5], q45);
  qmul(B[2], B[4], q24); qmul(B[3], B[5], q35); qmul(B[2], B[5], q25); qmul(B[3], B[4], q34);
  double S1[6], S2[6], S3[6], S4[6], S5[6], S6[6], S7[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    S1[i] = -q22[i] - q33[i] + q44[i] + q55[i];
    S2[i] = q22[i] - q33[i] + q44[i] - q55[i];
    S3[i] = 2.0 * (q23[i] + q45[i]);
    S4[i] = 2.0 * (q23[i] - q45[i]);
    S5[i] = -q22[i] + q33[i] + q44[i] - q55[i];
    S6[i] = 2.0 * (q24[i] - q35[i]);
    S7[i] = 2.0 * (q25[i] + q34[i]);
  }
  double rows[6][10];
  cmul_lq<false>(B[0], S4, rows[0]); cmul_lq<true>(B[1], S5, rows[0]);  // T10
  cmul_lq<false>(B[4], S1, rows[1]);                                     // T20
  cmul_lq<false>(B[0], S2, rows[2]); cmul_lq<true>(B[1], S3, rows[2]);  // T00
  cmul_lq<false>(B[5], S1, rows[3]);                                     // T21
  cmul_lq<false>(B[3], S1, rows[4]);                                     // -T12
  cmul_lq<false>(B[0], S6, rows[5]); cmul_lq<true>(B[1], S7, rows[5]);  // T22
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 10; ++c) C[r][c] = (r == 4 ? -1.0 : 1.0) * rows[r][col_of<KIND>(c)];
}

// G = C[:,0:6]^-1 C[:,6:10] (PartialPivLU, src/spherical_solvers.cpp:279).  Row swaps are
// predicated moves so C stays in registers.  Only rows FIRST_ROW..5 of G are produced.
template <int FIRST_ROW>
SSFM_HD bool eliminate_G(double (&a)[6][10], double (&G)[6][4]) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(a[k][k]);
#pragma unroll
    for (int r = k + 1; r < 6; ++r) {
      const double v = fabs(a[r][k]);
      if (v > best) { best = v; piv = r; }
    }
#pragma unroll
    for (int r = k + 1; r < 6; ++r)
      if (r == piv) {
#pragma unroll
        for (int c = k; c < 10; ++c) { const double tmp = a[k][c]; a[k][c] = a[r][c]; a[r][c] = tmp; }
      }
    const double inv = 1.0 / a[k][k];
#pragma unroll
    for (int r = k + 1; r < 6; ++r) {
      const double f = a[r][k] * inv;
#pragma unroll
      for (int c = k + 1; c < 10; ++c) a[r][c] -= f * a[k][c];
    }
  }
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 5; r >= FIRST_ROW; --r) {
      double s = a[r][6 + j];
#pragma unroll
      for (int c = r + 1; c < 6; ++c) s -= a[r][c] * G[c][j];
      G[r][j] = s / a[r][r];
      ok = ok && isfinite(G[r][j]);
    }
  return ok;
}

// ---- quartic roots ---------------------------------------------------------------------------
// A monic quartic always splits into two REAL quadratics x^2 + al x + be.  They are found from the
// largest real root of the resolvent cubic and polished by Bairstow steps on the original quartic;
// each quadratic then yields a real pair or a conjugate pair.  (This plays the role of the 4x4
// eigenvalue problem at src/spherical_solvers.cpp:287 and of SolveQuartic at :15-69.)
SSFM_HD double cubic_largest_real_root(double a2, double a1, double a0) {  // m^3 + a2 m^2 + a1 m + a0
  const double sh = a2 / 3.0;
  const double P = a1 - a2 * a2 / 3.0;
  const double Q = 2.0 * a2 * a2 * a2 / 27.0 - a2 * a1 / 3.0 + a0;
  const double disc = 0.25 * Q * Q + P * P * P / 27.0;
  double t;
  if (disc >= 0.0) {
    const double sq = sqrt(disc);
    const double A = -(Q >= 0 ? 1.0 : -1.0) * cbrt(0.5 * fabs(Q) + sq);
    t = A != 0.0 ? A - P / (3.0 * A) : 0.0;
  } else {
    const double rr = sqrt(-P / 3.0);
    double arg = 3.0 * Q / (2.0 * P * rr);
    arg = arg > 1.0 ? 1.0 : (arg < -1.0 ? -1.0 : arg);
    t = 2.0 * rr * cos(acos(arg) / 3.0);
  }
  double m = t - sh;
  for (int it = 0; it < 3; ++it) {  // Newton polish
    const double f = ((m + a2) * m + a1) * m + a0;
    const double df = (3.0 * m + 2.0 * a2) * m + a1;
    if (df != 0.0 && isfinite(f / df)) m -= f / df;
  }
  return m;
}

// One Bairstow refinement of the factor x^2 + al x + be of x^4 + a x^3 + b x^2 + c x + d.
SSFM_HD void bairstow_step(double a, double b, double c, double d, double& al, double& be) {
  // divide: quartic = (x^2 + al x + be)(x^2 + q1 x + q0) + (r1 x + r0)
  const double q1 = a - al;
  const double q0 = b - be - al * q1;
  const double r1 = c - al * q0 - be * q1;
  const double r0 = d - be * q0;
  // divide the quotient again for the partial derivatives
  const double s1 = q1 - al;          // quotient (x^2+q1x+q0) / (x^2+al x+be) -> 1, remainder s1 x + s0
  const double s0 = q0 - be;
  // dr1/dal = -(q0 ... ) ; standard Bairstow Jacobian:
  //  d r1/d al = al*s1 - s0 , d r1/d be = -s1 , d r0/d al = be*s1 , d r0/d be = -s0
  const double j11 = al * s1 - s0, j12 = -s1, j21 = be * s1, j22 = -s0;
  const double det = j11 * j22 - j12 * j21;
  if (det == 0.0) return;
  const double dal = (-r1 * j22 + r0 * j12) / det;
  const double dbe = (-r0 * j11 + r1 * j21) / det;
  if (isfinite(dal) && isfinite(dbe)) { al += dal; be += dbe; }
}

// Roots of x^2 + al x + be: returns (re0, im0), (re1, im1); im = 0 for a real pair.
SSFM_HD void quad_roots(double al, double be, Cplx& r0, Cplx& r1) {
  const double disc = al * al - 4.0 * be;
  if (disc >= 0.0) {
    const double sq = sqrt(disc);
    const double qq = -0.5 * (al + (al >= 0 ? sq : -sq));
    r0 = {qq, 0.0};
    r1 = {qq != 0.0 ? be / qq : 0.0, 0.0};
  } else {
    const double sq = sqrt(-disc);
    r0 = {-0.5 * al, 0.5 * sq};
    r1 = {-0.5 * al, -0.5 * sq};
  }
}

// All four roots of c4 x^4 + c3 x^3 + c2 x^2 + c1 x + c0.
SSFM_HD void quartic_roots(double c4, double c3, double c2, double c1, double c0, Cplx* roots) {
  const double a = c3 / c4, b = c2 / c4, c = c1 / c4, d = c0 / c4;
  const double p = b - 0.375 * a * a;
  const double q = c - 0.5 * a * b + 0.125 * a * a * a;
  const double r = d - 0.25 * a * c + 0.0625 * a * a * b - 3.0 / 256.0 * a * a * a * a;
  // resolvent: m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, largest real root (>= 0)
  const double m = cubic_largest_real_root(p, 0.25 * p * p - r, -0.125 * q * q);
  double al1, be1, al2, be2;  // depressed factors y^2 + al y + be
  const double scale = fabs(p) + sqrt(fabs(r)) + cbrt(q * q);
  if (m > 1e-14 * scale && m > 0.0) {
    const double s = sqrt(2.0 * m);
    const double h = q / (2.0 * s);
    al1 = s;  be1 = 0.5 * p + m - h;
    al2 = -s; be2 = 0.5 * p + m + h;
  } else {
    // (numerically) biquadratic
    const double disc = p * p - 4.0 * r;
    if (disc >= 0.0) {
      const double sq = sqrt(disc);
      al1 = 0.0; be1 = 0.5 * (p + sq);
      al2 = 0.0; be2 = 0.5 * (p - sq);
    } else {
      const double sr = sqrt(r);
      const double s = sqrt(fmax(2.0 * sr - p, 0.0));
      al1 = s;  be1 = sr;
      al2 = -s; be2 = sr;
    }
  }
  // back to x = y - a/4:  y^2 + al y + be  ->  x^2 + (al + a/2) x + (be + al a/4 + a^2/16)
  double A1 = al1 + 0.5 * a, B1 = be1 + 0.25 * al1 * a + 0.0625 * a * a;
  double A2 = al2 + 0.5 * a, B2 = be2 + 0.25 * al2 * a + 0.0625 * a * a;
  for (int it = 0; it < 3; ++it) {
    bairstow_step(a, b, c, d, A1, B1);
    bairstow_step(a, b, c, d, A2, B2);
  }
  quad_roots(A1, B1, roots[0], roots[1]);
  quad_roots(A2, B2, roots[2], roots[3]);
  // Newton polish of the real roots on the quartic itself
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (roots[k].im == 0.0) {
      double x = roots[k].re;
      for (int it = 0; it < 2; ++it) {
        const double f = (((x + a) * x + b) * x + c) * x + d;
        const double df = ((4.0 * x + 3.0 * a) * x + 2.0 * b) * x + c;
        const double dx = f / df;
        if (df != 0.0 && isfinite(dx)) x -= dx;
      }
      roots[k].re = x;
    }
  }
}

// Canonical real model of a (possibly complex) projective solution pc = B (bx, by, bz):
// unit-Frobenius E along the major axis of { Re(e^{i th} pc) }.  For real solutions this is just
// normalisation (Esoln /= Esoln.norm(), src/spherical_solvers.cpp:305); for conjugate pairs both
// members give the same model, mirroring the reference's Re(eigenvector) (:296).
SSFM_HD void model_from_b(const double (&B)[6][3], Cplx bx, Cplx by, Cplx bz, double* p) {
  double pr[6], pi[6];
  double aa = 0, bb = 0, ab = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    pr[i] = B[i][0] * bx.re + B[i][1] * by.re + B[i][2] * bz.re;
    pi[i] = B[i][0] * bx.im + B[i][1] * by.im + B[i][2] * bz.im;
    aa += pr[i] * pr[i];
    bb += pi[i] * pi[i];
    ab += pr[i] * pi[i];
  }
  double c = 1.0, s = 0.0;
  if (bb > 0.0) {
    const double th = 0.5 * atan2(-2.0 * ab, aa - bb);
    c = cos(th);
    s = sin(th);
  }
  double nrm = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    p[i] = pr[i] * c - pi[i] * s;
    nrm += p[i] * p[i];
  }
  const double inv = 1.0 / sqrt(nrm + p[0] * p[0] + p[1] * p[1]);
#pragma unroll
  for (int i = 0; i < 6; ++i) p[i] *= inv;
}

// KIND 0 (action matrix): eigen-pairs of M = [-G2; -G4; -G5; e1^T] on the basis [y^2, x, y, 1]
// (src/spherical_solvers.cpp:281-308): eigenvalue x from the characteristic quartic, then (y^2, y)
// from the best-conditioned 2x2 subsystem of rows 0..2 of (M - x I) v = 0 with v = (v0, x, v2, 1).
SSFM_HD void roots_action_matrix(const double (&G)[6][4], Cplx* xs, Cplx* ys) {
  double M[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { M[0][j] = -G[2][j]; M[1][j] = -G[4][j]; M[2][j] = -G[5][j]; }
  // det(M - L I) = 1 * cof(3,1) + (-L) * cof(3,3);  cof(3,1) = +det(rows 0..2, cols {0,2,3}),
  // cof(3,3) = +det(rows 0..2, cols {0,1,2}).  Build both as polynomials in L.
  // minor33(L) = det [[m00-L, m01, m02],[m10, m11-L, m12],[m20, m21, m22-L]]  (cubic)
  const double m00 = M[0][0], m01 = M[0][1], m02 = M[0][2], m03 = M[0][3];
  const double m10 = M[1][0], m11 = M[1][1], m12 = M[1][2], m13 = M[1][3];
  const double m20 = M[2][0], m21 = M[2][1], m22 = M[2][2], m23 = M[2][3];
  const double tr = m00 + m11 + m22;
  const double pm = (m00 * m11 - m01 * m10) + (m00 * m22 - m02 * m20) + (m11 * m22 - m12 * m21);
  const double dt = m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
  // minor33(L) = -L^3 + tr L^2 - pm L + dt
  // minor31(L) = det [[m00-L, m02, m03],[m10, m12, m13],[m20, m22-L, m23]]
  //            = (m00-L)(m12 m23 - m13 (m22-L)) - m02 (m10 m23 - m13 m20) + m03 (m10 (m22-L) - m12 m20)
  // expand in L:  (m00 - L)(k0 + m13 L) - m02 k1 + m03 (k2 - m10 L),  k0 = m12 m23 - m13 m22
  const double k0 = m12 * m23 - m13 * m22;
  const double k1 = m10 * m23 - m13 * m20;
  const double k2 = m10 * m22 - m12 * m20;
  const double n2 = -m13;
  const double n1 = m00 * m13 - k0 - m03 * m10;
  const double n0 = m00 * k0 - m02 * k1 + m03 * k2;
  // det(M - L I) = minor31 - L * minor33 = L^4 - tr L^3 + (pm + n2) L^2 + (n1 - dt) L + n0
  Cplx rt[4];
  quartic_roots(1.0, -tr, pm + n2, n1 - dt, n0, rt);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const Cplx x = rt[k];
    // row r: a_r v0 + b_r v2 = rhs_r
    const Cplx a0 = {m00 - x.re, -x.im}, b0 = {m02, 0.0};
    const Cplx a1 = {m10, 0.0}, b1 = {m12, 0.0};
    const Cplx a2 = {m20, 0.0}, b2 = {m22 - x.re, -x.im};
    const Cplx xx = cmul(x, x);
    const Cplx h0 = {-(m01 * x.re + m03), -(m01 * x.im)};
    const Cplx h1 = {-(m11 * x.re + m13) + xx.re, -(m11 * x.im) + xx.im};
    const Cplx h2 = {-(m21 * x.re + m23), -(m21 * x.im)};
    const Cplx d01 = csub(cmul(a0, b1), cmul(a1, b0));
    const Cplx d02 = csub(cmul(a0, b2), cmul(a2, b0));
    const Cplx d12 = csub(cmul(a1, b2), cmul(a2, b1));
    const double n01 = cabs2(d01), n02 = cabs2(d02), n12 = cabs2(d12);
    Cplx y;
    if (n01 >= n02 && n01 >= n12) {
      y = cdiv(csub(cmul(a0, h1), cmul(a1, h0)), d01);
    } else if (n02 >= n12) {
      y = cdiv(csub(cmul(a0, h2), cmul(a2, h0)), d02);
    } else {
      y = cdiv(csub(cmul(a1, h2), cmul(a2, h1)), d12);
    }
    // Eigenpair refinement on M itself.  Roots of the characteristic quartic inherit the conditioning of polynomial
    // roots w.r.t. coefficients (measured at scale: models off by up to 2e-4 on ~1 sample in 10^6, enough to change a
    // RANSAC trajectory), whereas the reference's QR iteration on M is backward stable.  Two Gauss-Newton steps on the
    // three non-trivial rows of (M - x I) (y^2, x, y, 1)^T = 0 in the unknowns (x, y) restore that accuracy:
    //   r0 = m00 y^2 + m01 x + m02 y + m03 - x y^2,  r1 = m10 y^2 + m11 x + m12 y + m13 - x^2,  r2 = m20 y^2 + m21 x + m22 y + m23 - x y
    Cplx px = x, py = y;
    if (x.im == 0.0) {
      // real eigenvalue: real arithmetic; a second step only when the first one moved the root by more than 1e-9
      double rx = x.re, ry = y.re;
      for (int it = 0; it < 2; ++it) {
        const double y2 = ry * ry, xy = rx * ry;
        const double r0 = m00 * y2 + m01 * rx + m02 * ry + m03 - rx * y2;
        const double r1 = m10 * y2 + m11 * rx + m12 * ry + m13 - rx * rx;
        const double r2 = m20 * y2 + m21 * rx + m22 * ry + m23 - xy;
        const double jx0 = m01 - y2, jx1 = m11 - 2.0 * rx, jx2 = m21 - ry;
        const double jy0 = 2.0 * m00 * ry + m02 - 2.0 * xy, jy1 = 2.0 * m10 * ry + m12, jy2 = 2.0 * m20 * ry + m22 - rx;
        const double a11 = jx0 * jx0 + jx1 * jx1 + jx2 * jx2, a22 = jy0 * jy0 + jy1 * jy1 + jy2 * jy2;
        const double a12 = jx0 * jy0 + jx1 * jy1 + jx2 * jy2;
        const double b1 = jx0 * r0 + jx1 * r1 + jx2 * r2, b2 = jy0 * r0 + jy1 * r1 + jy2 * r2;
        const double det = a11 * a22 - a12 * a12;
        if (!(det > 1e-30 * a11 * a22) || !isfinite(det)) break;  // (numerically) multiple root: leave it
        const double inv = 1.0 / det;
        const double dx = (a22 * b1 - a12 * b2) * inv, dy = (a11 * b2 - a12 * b1) * inv;
        const double step2 = dx * dx + dy * dy, size2 = rx * rx + ry * ry + 1.0;
        if (!isfinite(step2) || step2 > 1e-4 * size2) break;  // a polish, never a jump to another root
        rx -= dx;
        ry -= dy;
        if (step2 <= 1e-18 * size2) break;
      }
      px = {rx, 0.0};
      py = {ry, 0.0};
    } else {
      for (int it = 0; it < 2; ++it) {
        const Cplx y2 = cmul(py, py);
        const Cplx xy = cmul(px, py);
        Cplx r[3], jx[3], jy[3];
        r[0] = csub(cadd(cadd(cscale(m00, y2), cscale(m01, px)), cadd(cscale(m02, py), Cplx{m03, 0.0})), cmul(px, y2));
        r[1] = csub(cadd(cadd(cscale(m10, y2), cscale(m11, px)), cadd(cscale(m12, py), Cplx{m13, 0.0})), cmul(px, px));
        r[2] = csub(cadd(cadd(cscale(m20, y2), cscale(m21, px)), cadd(cscale(m22, py), Cplx{m23, 0.0})), xy);
        jx[0] = csub(Cplx{m01, 0.0}, y2);
        jx[1] = csub(Cplx{m11, 0.0}, cscale(2.0, px));
        jx[2] = csub(Cplx{m21, 0.0}, py);
        jy[0] = csub(cadd(cscale(2.0 * m00, py), Cplx{m02, 0.0}), cscale(2.0, xy));
        jy[1] = cadd(cscale(2.0 * m10, py), Cplx{m12, 0.0});
        jy[2] = csub(cadd(cscale(2.0 * m20, py), Cplx{m22, 0.0}), px);
        // normal equations (J^H J) d = J^H r, 2x2 Hermitian
        double a11 = 0.0, a22 = 0.0;
        Cplx a12 = {0.0, 0.0}, b1 = {0.0, 0.0}, b2 = {0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const Cplx cjx = {jx[q].re, -jx[q].im}, cjy = {jy[q].re, -jy[q].im};
          a11 += cabs2(jx[q]);
          a22 += cabs2(jy[q]);
          a12 = cadd(a12, cmul(cjx, jy[q]));
          b1 = cadd(b1, cmul(cjx, r[q]));
          b2 = cadd(b2, cmul(cjy, r[q]));
        }
        const double det = a11 * a22 - cabs2(a12);
        if (!(det > 1e-30 * a11 * a22) || !isfinite(det)) break;
        const Cplx ca12 = {a12.re, -a12.im};
        const double inv = 1.0 / det;
        const Cplx dx = cscale(inv, csub(cscale(a22, b1), cmul(a12, b2)));
        const Cplx dy = cscale(inv, csub(cscale(a11, b2), cmul(ca12, b1)));
        const double step2 = cabs2(dx) + cabs2(dy), size2 = cabs2(px) + cabs2(py) + 1.0;
        if (!isfinite(step2) || step2 > 1e-4 * size2) break;
        px = csub(px, dx);
        py = csub(py, dy);
        if (step2 <= 1e-18 * size2) break;
      }
    }
    xs[k] = px;
    ys[k] = py;
  }
}

// Solve one minimal sample.  models: 4 x 6.  Returns the number of models (KIND 0/1: always 4,
// NaN-filled when the elimination breaks down; KIND 2: real roots with |y| <= 10 only).
// skip_complex (KIND 0 only): models that come from a complex eigenvalue of the action matrix are left out (NaN)
// -- upstream's own commented-out filter (src/spherical_solvers.cpp:294).  Otherwise they are the canonical
// representative (model_from_b): the reference's Re(eigenvector) has a phase fixed by rounding noise in Eigen's QR
// sweeps, which no second implementation can reproduce.
// `stage()` is called between the solver's stages: a no-op by default; the batched kernel passes a block barrier so that all
// warps of a CTA run the same stage at the same time and share its instructions in the instruction cache.
struct NoStageSync {
  SSFM_HD void operator()() const {}
};
template <int KIND, class StageSync = NoStageSync>
SSFM_HD_NOINLINE int solve_minimal(const double* u0, const double* v0, const double* u1, const double* v1, const double* u2,
                          const double* v2, double (&models)[4][6], bool skip_complex = false, StageSync stage = StageSync()) {
  double m[6][3];
  {
    double a[6];
    epipolar_row(u0, v0, a);
#pragma unroll
    for (int r = 0; r < 6; ++r) m[r][0] = a[r];
    epipolar_row(u1, v1, a);
#pragma unroll
    for (int r = 0; r < 6; ++r) m[r][1] = a[r];
    epipolar_row(u2, v2, a);
#pragma unroll
    for (int r = 0; r < 6; ++r) m[r][2] = a[r];
  }
  double B[6][3];
  nullspace_colpiv<3>(m, B);
  stage();
  double C[6][10], G[6][4];
  build_constraints<KIND>(B, C);
  stage();
  const double nanv = nan("");
  if (KIND == 0) {
    const bool ok = eliminate_G<2>(C, G);
    stage();
    Cplx xs[4], ys[4];
    if (ok) roots_action_matrix(G, xs, ys);
    stage();
    if (!ok) {
      for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 6; ++i) models[k][i] = nanv;
      return 4;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      model_from_b(B, xs[k], ys[k], Cplx{1.0, 0.0}, models[k]);
      if (skip_complex && xs[k].im != 0.0)
        for (int i = 0; i < 6; ++i) models[k][i] = nanv;
    }
    return 4;
  } else if (KIND == 1) {
    const bool ok = eliminate_G<4>(C, G);
    if (!ok) {
      for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 6; ++i) models[k][i] = nanv;
      return 4;
    }
    // quartic in y (src/spherical_solvers.cpp:623-627), x from row 5 (:633-640)
    Cplx yr[4];
    quartic_roots(-G[5][0], G[4][0] - G[5][1], G[4][1] - G[5][2], G[4][2] - G[5][3], G[4][3], yr);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // SolveQuarticReals without a tolerance keeps the REAL PART of every Ferrari root (:73-83, call at :631);
      // x is row 5 evaluated at that real y.  Deterministic upstream, so it is matched (not canonicalised).
      const double y = yr[k].re;
      const double y2 = y * y, y3 = y2 * y;
      const double x = -G[5][0] * y3 - G[5][1] * y2 - G[5][2] * y - G[5][3];
      model_from_b(B, Cplx{x, 0.0}, Cplx{y, 0.0}, Cplx{1.0, 0.0}, models[k]);
    }
    return 4;
  } else {
    const bool ok = eliminate_G<3>(C, G);
    for (int k = 0; k < 4; ++k)
      for (int i = 0; i < 6; ++i) models[k][i] = nanv;
    if (!ok) return 0;
    // det N(y) (src/spherical_fast_estimator.cpp:219), real roots in [-10,10] (:223), x by Cramer (:239)
    const double c4 = G[4][0] * G[5][1] - G[4][1] * G[5][0];
    const double c3 = G[3][1] * G[5][0] - G[3][0] * G[5][1] + G[4][0] * G[5][2] - G[4][2] * G[5][0];
    const double c2 = G[3][2] * G[5][0] - G[3][1] * G[4][0] + G[3][0] * (G[4][1] - G[5][2]);
    const double c1 = G[3][0] * (G[4][2] + G[4][1] * G[5][3] - G[4][3] * G[5][1]) +
                      G[3][3] * (G[4][0] * G[5][1] - G[4][1] * G[5][0]) -
                      G[3][1] * (G[4][0] * G[5][3] - G[4][3] * G[5][0]) - G[3][2] * G[4][0];
    const double c0 = G[3][3] * (G[4][0] * G[5][2] - G[4][2] * G[5][0]) -
                      G[3][2] * (G[4][0] * G[5][3] - G[4][3] * G[5][0]) +
                      G[3][0] * (G[4][2] * G[5][3] - G[4][3] * G[5][2]);
    Cplx yr[4];
    quartic_roots(c4, c3, c2, c1, c0, yr);
    // ascending order of the real roots, like a Sturm bracketing sweep
    double ys[4];
    int nr = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (yr[k].im == 0.0 && yr[k].re >= -10.0 && yr[k].re <= 10.0) {
        int pos = nr;
        for (int j = nr - 1; j >= 0; --j)
          if (ys[j] > yr[k].re) { ys[j + 1] = ys[j]; pos = j; }
        ys[pos] = yr[k].re;
        ++nr;
      }
    int nm = 0;
    for (int k = 0; k < nr; ++k) {
      const double y = ys[k];
      const double N00 = G[3][0], N01 = G[3][2] + G[3][1] * y, N02 = G[3][3] + y * y * y;
      const double N10 = G[4][0], N11 = G[4][2] + G[4][1] * y, N12 = G[4][3] + y * y;
      const double x = (N02 * N10 - N00 * N12) / (N00 * N11 - N01 * N10);
      if (x != x) continue;
      model_from_b(B, Cplx{x, 0.0}, Cplx{y, 0.0}, Cplx{1.0, 0.0}, models[nm]);
      ++nm;
    }
    return nm;
  }
}

// ---------------------------------------------------------------------------------------------
// Refit arithmetic: forward-mode jets over the 6 free parameters (r1, t1) of the reference's
// autodiff'd SampsonError functor (src/spherical_estimator.cpp:23-65, ceres::AngleAxisToRotationMatrix).
// ---------------------------------------------------------------------------------------------
struct Jet6 {
  double a;
  double v[6];
};
SSFM_HD Jet6 jconst(double x) { Jet6 r; r.a = x; for (int i = 0; i < 6; ++i) r.v[i] = 0.0; return r; }
SSFM_HD Jet6 jvar(double x, int k) { Jet6 r = jconst(x); r.v[k] = 1.0; return r; }
SSFM_HD Jet6 operator+(const Jet6& x, const Jet6& y) { Jet6 r; r.a = x.a + y.a; for (int i = 0; i < 6; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
SSFM_HD Jet6 operator-(const Jet6& x, const Jet6& y) { Jet6 r; r.a = x.a - y.a; for (int i = 0; i < 6; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
SSFM_HD Jet6 operator-(const Jet6& x) { Jet6 r; r.a = -x.a; for (int i = 0; i < 6; ++i) r.v[i] = -x.v[i]; return r; }
SSFM_HD Jet6 operator*(const Jet6& x, const Jet6& y) { Jet6 r; r.a = x.a * y.a; for (int i = 0; i < 6; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
SSFM_HD Jet6 operator*(const Jet6& x, double s) { Jet6 r; r.a = x.a * s; for (int i = 0; i < 6; ++i) r.v[i] = x.v[i] * s; return r; }
SSFM_HD Jet6 operator/(const Jet6& x, const Jet6& y) { Jet6 r; const double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < 6; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
SSFM_HD Jet6 jsqrt(const Jet6& x) { Jet6 r; r.a = sqrt(x.a); const double f = 0.5 / r.a; for (int i = 0; i < 6; ++i) r.v[i] = x.v[i] * f; return r; }
SSFM_HD double jsqrt(double x) { return sqrt(x); }
SSFM_HD void jsincos(const Jet6& x, Jet6& s, Jet6& c) {
  const double sv = sin(x.a), cv = cos(x.a);
  s.a = sv; c.a = cv;
  for (int i = 0; i < 6; ++i) { s.v[i] = cv * x.v[i]; c.v[i] = -sv * x.v[i]; }
}
SSFM_HD void jsincos(double x, double& s, double& c) { s = sin(x); c = cos(x); }
SSFM_HD double jval(double x) { return x; }
SSFM_HD double jval(const Jet6& x) { return x.a; }
SSFM_HD double jlift(double x, double*) { return x; }
SSFM_HD Jet6 jlift(double x, Jet6*) { return jconst(x); }

// E = [t]x R with R = exp(r1) (ceres::AngleAxisToRotationMatrix), t = -R t0 + t1, t0 = (0,0,t0z): the
// model part of the reference's SampsonError functor (src/spherical_estimator.cpp:34-53).  It does not
// depend on the correspondence, so the refit evaluates it ONCE per LM iteration (as jets: E[i].a is
// the matrix, E[i].v[k] its derivative w.r.t. parameter k) instead of once per residual.
template <typename T>
SSFM_HD void spherical_E_of_params(const T* r1, const T* t1, double t0z, T* E) {
  T R[9];
  const T theta2 = r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2];
  const T one = jlift(1.0, (T*)0);
  if (jval(theta2) > 2.220446049250313e-16) {
    const T theta = jsqrt(theta2);
    const T wx = r1[0] / theta, wy = r1[1] / theta, wz = r1[2] / theta;
    T st, ct;
    jsincos(theta, st, ct);
    const T omc = one - ct;
    R[0] = ct + wx * wx * omc;
    R[3] = wz * st + wx * wy * omc;
    R[6] = -(wy * st) + wx * wz * omc;
    R[1] = wx * wy * omc - wz * st;
    R[4] = ct + wy * wy * omc;
    R[7] = wx * st + wy * wz * omc;
    R[2] = wy * st + wx * wz * omc;
    R[5] = -(wx * st) + wy * wz * omc;
    R[8] = ct + wz * wz * omc;
  } else {
    R[0] = one; R[3] = r1[2]; R[6] = -r1[1];
    R[1] = -r1[2]; R[4] = one; R[7] = r1[0];
    R[2] = r1[1]; R[5] = -r1[0]; R[8] = one;
  }
  T t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = R[3 * i + 2] * (-t0z) + t1[i];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    E[0 + j] = t[1] * R[6 + j] - t[2] * R[3 + j];
    E[3 + j] = t[2] * R[0 + j] - t[0] * R[6 + j];
    E[6 + j] = t[0] * R[3 + j] - t[1] * R[0 + j];
  }
}

// residual r = d^2 / den of one correspondence (value only)
SSFM_HD double sampson_value(const double* E, const double* u, const double* v) {
  const double Eu0 = E[0] * u[0] + E[1] * u[1] + E[2] * u[2];
  const double Eu1 = E[3] * u[0] + E[4] * u[1] + E[5] * u[2];
  const double Eu2 = E[6] * u[0] + E[7] * u[1] + E[8] * u[2];
  const double Et0 = E[0] * v[0] + E[3] * v[1] + E[6] * v[2];
  const double Et1 = E[1] * v[0] + E[4] * v[1] + E[7] * v[2];
  const double d = Eu0 * v[0] + Eu1 * v[1] + Eu2 * v[2];
  return (d * d) / (Eu0 * Eu0 + Eu1 * Eu1 + Et0 * Et0 + Et1 * Et1);
}

// residual and its gradient w.r.t. the 6 parameters, given E as jets: de/dE (9 closed-form entries)
// contracted with dE/dx (chain rule of the expression tree the reference autodiffs,
// src/spherical_estimator.cpp:55-61).
SSFM_HD void sampson_value_grad(const Jet6* E, const double* u, const double* v, double& r, double* g) {
  const double Eu0 = E[0].a * u[0] + E[1].a * u[1] + E[2].a * u[2];
  const double Eu1 = E[3].a * u[0] + E[4].a * u[1] + E[5].a * u[2];
  const double Eu2 = E[6].a * u[0] + E[7].a * u[1] + E[8].a * u[2];
  const double Et0 = E[0].a * v[0] + E[3].a * v[1] + E[6].a * v[2];
  const double Et1 = E[1].a * v[0] + E[4].a * v[1] + E[7].a * v[2];
  const double d = Eu0 * v[0] + Eu1 * v[1] + Eu2 * v[2];
  const double den = Eu0 * Eu0 + Eu1 * Eu1 + Et0 * Et0 + Et1 * Et1;
  const double inv = 1.0 / den;
  r = (d * d) * inv;
  // de/dE_ij = (2 d v_i u_j - r * dden/dE_ij) / den,  dden/dE_ij = 2 (Eu_i u_j [i<2] + Et_j v_i [j<2])
  const double a2 = 2.0 * d * inv, b2 = 2.0 * r * inv;
  double w[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double dd = 0.0;
      if (i < 2) dd += (i == 0 ? Eu0 : Eu1) * u[j];
      if (j < 2) dd += (j == 0 ? Et0 : Et1) * v[i];
      w[3 * i + j] = a2 * v[i] * u[j] - b2 * dd;
    }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) s += w[q] * E[q].v[k];
    g[k] = s;
  }
}

// 6x6 SPD solve (DENSE_NORMAL_CHOLESKY, src/spherical_estimator.cpp:148).  H is the lower
// triangle packed row-wise (21 entries).  Returns false if not positive definite.
SSFM_HD bool cholesky_solve6(const double* Hp, const double* b, double* x) {
  double L[21], inv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = Hp[j * (j + 1) / 2 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
    if (!(d > 0.0)) return false;
    const double ljj = sqrt(d);
    L[j * (j + 1) / 2 + j] = ljj;
    inv[j] = 1.0 / ljj;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = Hp[i * (i + 1) / 2 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      L[i * (i + 1) / 2 + j] = s * inv[j];
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * (i + 1) / 2 + k] * y[k];
    y[i] = s * inv[i];
  }
  bool ok = true;
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k * (k + 1) / 2 + i] * x[k];
    x[i] = s * inv[i];
    ok = ok && isfinite(x[i]);
  }
  return ok;
}

// NumRequiredIterations (include/RansacLib/utils.h:110-140)
SSFM_HD uint32_t required_iterations(double w, double eta, int k, uint32_t lo, uint32_t hi) {
  if (w <= 0.0) return hi;
  if (w >= 1.0) return lo;
  const double miss = 1.0 - pow(w, (double)k);
  if (miss >= 0.99999999999999) return hi;
  const double n = ceil(log(eta) / log(miss) + 0.5);
  uint32_t it = n >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)n;
  it = it < hi ? it : hi;
  return lo > it ? lo : it;
}

// ---------------------------------------------------------------------------------------------
// The LO generator: std::mt19937 + libstdc++'s std::uniform_int_distribution<int>
// (include/RansacLib/ransac.h:143-144, utils.h:34-45), restated so the device consumes exactly
// the reference's draw sequence.  State lives in global memory: 624 words + position.
// ---------------------------------------------------------------------------------------------
SSFM_HD void mt19937_seed(uint32_t* mt, uint32_t seed) {
  mt[0] = seed;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  mt[624] = 624;  // position: forces a twist before the first draw
}
SSFM_HD void mt19937_twist(uint32_t* mt) {
  for (int i = 0; i < 624; ++i) {
    const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
    mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  mt[624] = 0;
}
SSFM_HD uint32_t mt19937_next(uint32_t* mt) {
  if (mt[624] >= 624) mt19937_twist(mt);
  uint32_t y = mt[mt[624]++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// uniform_int_distribution<int>(lo, hi)(mt19937): libstdc++ (GCC >= 11) uses Lemire's
// nearly-divisionless method when the generator range is 2^32.
SSFM_HD int uniform_int_libstdcxx(uint32_t* mt, int lo, int hi) {
  const uint32_t urange = (uint32_t)hi - (uint32_t)lo;
  if (urange == 0xFFFFFFFFu) return (int)(mt19937_next(mt) + (uint32_t)lo);
  const uint32_t range = urange + 1u;
  uint64_t product = (uint64_t)mt19937_next(mt) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)mt19937_next(mt) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (int)((uint32_t)(product >> 32) + (uint32_t)lo);
}

}  // namespace ssfm
