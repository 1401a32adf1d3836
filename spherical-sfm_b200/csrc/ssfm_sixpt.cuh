// ssfm_sixpt.cuh -- six-point shared-focal relative pose (SixPointEstimator::MinimalSolver,
// examples/six_point_estimator.cpp:93-119, which calls poselib::relpose_6pt_shared_focal, :104) as
// per-thread float64 code.
//
// PoseLib is not part of the reference tree (cloned at an unpinned HEAD, docker/Dockerfile:58-62), so what is
// implemented is the published problem it solves: with F = x F0 + y F1 + F2 spanning the null space of the six
// epipolar constraints, w = 1/f^2, Q = diag(1,1,w), all real (x, y, w > 0) with
//      det F = 0,     2 F Q F^T Q F - trace(F Q F^T Q) F = 0       (10 cubics in x,y, quadratic in w)
// i.e. the quadratic eigenvalue problem (M0 + w M1 + w^2 M2) m(x,y) = 0 of Kukelova, Bujnak, Pajdla (BMVC
// 2008).  Here: companion linearisation in mu = 1/w = f^2, balancing + elimination to Hessenberg form + Francis
// double-shift QR for the eigenvalues (per thread, 20x20 in local memory), a Householder least-squares for
// (x,y) at each (nearly) real positive eigenvalue, Gauss-Newton steps on the original ten equations, then the
// essential matrix K F K is decomposed and the (R, t) with every sample point in front of both cameras are
// kept (PoseLib's cheirality rule; |t| = 1).  Solutions are ordered by focal length.
#pragma once
#include "ssfm_math.cuh"

namespace ssfm {

constexpr int kSixMaxModels = 15;  // "<= 15 {t, r, f}" (the problem has 15 complex solutions)

struct SixPointModel {
  double t[3];  // unit translation
  double r[3];  // so3ln(R)
  double f;     // focal length in the units of the input rays' x,y
};

namespace sixpt {

// polynomials in (x,y): deg1 = [x,y,1]; deg2 = [x2,xy,y2,x,y,1]; deg3 = [x3,x2y,xy2,y3,x2,xy,y2,x,y,1]
SSFM_HD void mul11(const double* a, const double* b, double* o) {
  o[0] = a[0] * b[0];
  o[1] = a[0] * b[1] + a[1] * b[0];
  o[2] = a[1] * b[1];
  o[3] = a[0] * b[2] + a[2] * b[0];
  o[4] = a[1] * b[2] + a[2] * b[1];
  o[5] = a[2] * b[2];
}
// o += c * p * a
SSFM_HD void fma21(double c, const double* p, const double* a, double* o) {
  o[0] += c * (p[0] * a[0]);
  o[1] += c * (p[0] * a[1] + p[1] * a[0]);
  o[2] += c * (p[1] * a[1] + p[2] * a[0]);
  o[3] += c * (p[2] * a[1]);
  o[4] += c * (p[0] * a[2] + p[3] * a[0]);
  o[5] += c * (p[1] * a[2] + p[3] * a[1] + p[4] * a[0]);
  o[6] += c * (p[2] * a[2] + p[4] * a[1]);
  o[7] += c * (p[3] * a[2] + p[5] * a[0]);
  o[8] += c * (p[4] * a[2] + p[5] * a[1]);
  o[9] += c * (p[5] * a[2]);
}

// Null space of the 6x9 epipolar system by Gauss-Jordan with complete pivoting, then modified
// Gram-Schmidt.  Fb[k] (k<3) is a 3x3 row-major basis matrix.  Returns false if rank deficient.
// A: 54 doubles of workspace (the batched kernel passes shared memory: with local memory this was 11 % of its time).
SSFM_HD_NOINLINE bool nullspace_6x9_ws(const double (*x1)[3], const double (*x2)[3], double Fb[3][9], double (*A)[9]) {
  int colperm[9];
  for (int c = 0; c < 9; ++c) colperm[c] = c;
  for (int i = 0; i < 6; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) A[i][3 * r + c] = x2[i][r] * x1[i][c];
  for (int k = 0; k < 6; ++k) {
    int pr = k, pc = k;
    double best = -1.0;
    for (int i = k; i < 6; ++i)
      for (int j = k; j < 9; ++j) {
        const double v = fabs(A[i][j]);
        if (v > best) { best = v; pr = i; pc = j; }
      }
    if (!(best > 1e-14)) return false;
    if (pr != k)
      for (int j = 0; j < 9; ++j) { const double tmp = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = tmp; }
    if (pc != k) {
      for (int i = 0; i < 6; ++i) { const double tmp = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = tmp; }
      const int tc = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = tc;
    }
    const double inv = 1.0 / A[k][k];
    for (int j = k; j < 9; ++j) A[k][j] *= inv;
    for (int i = 0; i < 6; ++i) {
      if (i == k) continue;
      const double f = A[i][k];
      if (f != 0.0)
        for (int j = k; j < 9; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int q = 0; q < 3; ++q) {
    double v[9];
    for (int c = 0; c < 9; ++c) v[c] = 0.0;
    v[colperm[6 + q]] = 1.0;
    for (int i = 0; i < 6; ++i) v[colperm[i]] = -A[i][6 + q];
    for (int p = 0; p < q; ++p) {
      double d = 0.0;
      for (int c = 0; c < 9; ++c) d += v[c] * Fb[p][c];
      for (int c = 0; c < 9; ++c) v[c] -= d * Fb[p][c];
    }
    double nn = 0.0;
    for (int c = 0; c < 9; ++c) nn += v[c] * v[c];
    nn = 1.0 / sqrt(nn);
    for (int c = 0; c < 9; ++c) Fb[q][c] = v[c] * nn;
  }
  return true;
}

SSFM_HD bool nullspace_6x9(const double (*x1)[3], const double (*x2)[3], double Fb[3][9]) {
  double A[6][9];
  return nullspace_6x9_ws(x1, x2, Fb, A);
}

// M[k][e][mono]: coefficient of w^k in equation e.  Equation 0 = det F, 1..9 = the trace constraint.
SSFM_HD_NOINLINE void constraint_matrices(const double Fb[3][9], double M[3][10][10]) {
  double F[9][3];  // entry (i,j) as deg1 poly [x,y,1]
  for (int e = 0; e < 9; ++e) { F[e][0] = Fb[0][e]; F[e][1] = Fb[1][e]; F[e][2] = Fb[2][e]; }
  double G0[3][3][6], G1[3][3][6];
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      double a[6], b[6];
      mul11(F[3 * i + 0], F[3 * j + 0], a);
      mul11(F[3 * i + 1], F[3 * j + 1], b);
      for (int q = 0; q < 6; ++q) G0[i][j][q] = G0[j][i][q] = a[q] + b[q];
      mul11(F[3 * i + 2], F[3 * j + 2], a);
      for (int q = 0; q < 6; ++q) G1[i][j][q] = G1[j][i][q] = a[q];
    }
  double tr0[6], tr1[6], tr2[6];
  for (int q = 0; q < 6; ++q) {
    tr0[q] = G0[0][0][q] + G0[1][1][q];
    tr1[q] = G1[0][0][q] + G1[1][1][q] + G0[2][2][q];
    tr2[q] = G1[2][2][q];
  }
  for (int k = 0; k < 3; ++k)
    for (int e = 0; e < 10; ++e)
      for (int q = 0; q < 10; ++q) M[k][e][q] = 0.0;
  {  // det F
    double m0[6], m1[6], c[6];
    mul11(F[4], F[8], m0); mul11(F[5], F[7], m1);
    for (int q = 0; q < 6; ++q) c[q] = m0[q] - m1[q];
    fma21(1.0, c, F[0], M[0][0]);
    mul11(F[3], F[8], m0); mul11(F[5], F[6], m1);
    for (int q = 0; q < 6; ++q) c[q] = m0[q] - m1[q];
    fma21(-1.0, c, F[1], M[0][0]);
    mul11(F[3], F[7], m0); mul11(F[4], F[6], m1);
    for (int q = 0; q < 6; ++q) c[q] = m0[q] - m1[q];
    fma21(1.0, c, F[2], M[0][0]);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int e = 1 + 3 * i + j;
      // (2 H F - tr(H) F)_{ij},  H = (G0 + w G1) diag(1,1,w)
      fma21(2.0, G0[i][0], F[0 + j], M[0][e]);
      fma21(2.0, G0[i][1], F[3 + j], M[0][e]);
      fma21(-1.0, tr0, F[3 * i + j], M[0][e]);
      fma21(2.0, G1[i][0], F[0 + j], M[1][e]);
      fma21(2.0, G1[i][1], F[3 + j], M[1][e]);
      fma21(2.0, G0[i][2], F[6 + j], M[1][e]);
      fma21(-1.0, tr1, F[3 * i + j], M[1][e]);
      fma21(2.0, G1[i][2], F[6 + j], M[2][e]);
      fma21(-1.0, tr2, F[3 * i + j], M[2][e]);
    }
}

constexpr int kN = 16;  // the deflated companion matrix (see companion16)

// Where a solver instance keeps its kN x kN matrix.  LocalMat: a plain array (host build, hooks).  StridedMat: element
// (i, j) of thread t at base[(i * kN + j) * stride + t] -- shared memory of the batched kernel: every lane of a warp that
// touches the same (i, j) hits a different bank, and the 2 KB per instance never go through local memory.
struct LocalMat {
  double a[kN][kN];
  SSFM_HD double& operator()(int i, int j) { return a[i][j]; }
};
struct StridedMat {
  double* base;
  int stride;
  SSFM_HD double& operator()(int i, int j) { return base[(size_t)(i * kN + j) * stride]; }
};

// Eigenvalue-preserving diagonal scaling (powers of two).
template <class Mat>
SSFM_HD_NOINLINE void balance(Mat& a) {
  for (int pass = 0; pass < 20; ++pass) {
    bool done = true;
    for (int i = 0; i < kN; ++i) {
      double r = 0.0, c = 0.0;
      for (int j = 0; j < kN; ++j)
        if (j != i) { c += fabs(a(j, i)); r += fabs(a(i, j)); }
      if (c != 0.0 && r != 0.0) {
        double g = r * 0.5, f = 1.0;
        const double s = c + r;
        while (c < g) { f *= 2.0; c *= 4.0; }
        g = r * 2.0;
        while (c > g) { f *= 0.5; c *= 0.25; }
        if ((c + r) / f < 0.95 * s) {
          done = false;
          g = 1.0 / f;
          for (int j = 0; j < kN; ++j) a(i, j) *= g;
          for (int j = 0; j < kN; ++j) a(j, i) *= f;
        }
      }
    }
    if (done) break;
  }
}

// Reduction to upper Hessenberg form by stabilised elementary similarity transformations.
template <class Mat>
SSFM_HD_NOINLINE void to_hessenberg(Mat& a) {
  for (int m = 1; m < kN - 1; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j < kN; ++j) {
      const double v = a(j, m - 1);
      if (fabs(v) > fabs(x)) { x = v; i = j; }
    }
    if (i != m) {
      for (int j = m - 1; j < kN; ++j) { const double t = a(i, j); a(i, j) = a(m, j); a(m, j) = t; }
      for (int j = 0; j < kN; ++j) { const double t = a(j, i); a(j, i) = a(j, m); a(j, m) = t; }
    }
    if (x != 0.0) {
      for (int r = m + 1; r < kN; ++r) {
        double y = a(r, m - 1);
        if (y != 0.0) {
          y /= x;
          a(r, m - 1) = 0.0;
          for (int j = m; j < kN; ++j) a(r, j) -= y * a(m, j);
          for (int j = 0; j < kN; ++j) a(j, m) += y * a(j, r);
        }
      }
    }
  }
}

SSFM_HD double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// Eigenvalues of an upper Hessenberg matrix by the Francis double-shift QR iteration (destroys a).
// Returns false if an eigenvalue failed to converge.
template <class Mat>
SSFM_HD_NOINLINE bool hessenberg_eigenvalues(Mat& a, double* wr, double* wi) {
  double anorm = 0.0;
  for (int i = 0; i < kN; ++i)
    for (int j = (i > 0 ? i - 1 : 0); j < kN; ++j) anorm += fabs(a(i, j));
  int nn = kN - 1;
  double t = 0.0;
  double p = 0.0, q = 0.0, r = 0.0;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        double s = fabs(a(l - 1, l - 1)) + fabs(a(l, l));
        if (s == 0.0) s = anorm;
        if (fabs(a(l, l - 1)) + s == s) { a(l, l - 1) = 0.0; break; }
      }
      double x = a(nn, nn);
      if (l == nn) {  // one real root
        wr[nn] = x + t;
        wi[nn--] = 0.0;
      } else {
        double y = a(nn - 1, nn - 1);
        double w = a(nn, nn - 1) * a(nn - 1, nn);
        if (l == nn - 1) {  // a 2x2 block: two roots
          p = 0.5 * (y - x);
          q = p * p + w;
          double z = sqrt(fabs(q));
          x += t;
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
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20 || its == 30 || its == 40) {  // exceptional shift
            t += x;
            for (int i = 0; i <= nn; ++i) a(i, i) -= x;
            const double s = fabs(a(nn, nn - 1)) + fabs(a(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          double z;
          for (m = nn - 2; m >= l; --m) {  // look for two consecutive small sub-diagonal elements
            z = a(m, m);
            r = x - z;
            double s = y - z;
            p = (r * s - w) / a(m + 1, m) + a(m, m + 1);
            q = a(m + 1, m + 1) - z - r - s;
            r = a(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s; q /= s; r /= s;
            if (m == l) break;
            const double u = fabs(a(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(a(m - 1, m - 1)) + fabs(z) + fabs(a(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            a(i, i - 2) = 0.0;
            if (i != m + 2) a(i, i - 3) = 0.0;
          }
          for (int k = m; k <= nn - 1; ++k) {  // double QR step on rows l..nn, columns m..nn
            if (k != m) {
              p = a(k, k - 1);
              q = a(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = a(k + 2, k - 1);
              x = fabs(p) + fabs(q) + fabs(r);
              if (x != 0.0) { p /= x; q /= x; r /= x; }
            }
            const double s = sign_of(sqrt(p * p + q * q + r * r), p);
            if (s != 0.0) {
              if (k == m) {
                if (l != m) a(k, k - 1) = -a(k, k - 1);
              } else {
                a(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s;
              y = q / s;
              z = r / s;
              q /= p;
              r /= p;
              const bool three = k != nn - 1;
              for (int j = k; j <= nn; ++j) {
                const double a0 = a(k, j), a1 = a(k + 1, j);
                p = a0 + q * a1;
                if (three) {
                  const double a2 = a(k + 2, j);
                  p += r * a2;
                  a(k + 2, j) = a2 - p * z;
                }
                a(k + 1, j) = a1 - p * y;
                a(k, j) = a0 - p * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                const double a0 = a(i, k), a1 = a(i, k + 1);
                p = x * a0 + y * a1;
                if (three) {
                  const double a2 = a(i, k + 2);
                  p += z * a2;
                  a(i, k + 2) = a2 - p * r;
                }
                a(i, k + 1) = a1 - p * q;
                a(i, k) = a0 - p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}

// The companion matrix of (M2 + mu M1 + mu^2 M0) m = 0 (mu = 1/w = f^2), with the four structurally zero eigenvalues
// deflated exactly.  Every row of M2 is a multiple of the (3,3) entry F22 = a x + b y + c of F:
//   M2[1 + 3i + j] = F22 * (2 F_i2 F_2j - F22 F_ij),  M2[0] = 0            (constraint_matrices)
// so the row space of M2 lies in the six-dimensional span of F22 * {x^2, xy, y^2, x, y, 1}, whatever the data: M2 = C B
// with B (6 x 10) made of a, b, c only.  With B^T = Q [R; 0] (Householder), the last four columns of Q span null(B), hence
// (M2 Q)[:, 6..9] = 0 up to rounding, and in the coordinates m = Q m' the companion matrix
//   [ 0, I ; -A0^-1 A2, -A0^-1 A1 ],  Ak = Mk Q
// has four zero columns: deleting those rows and columns leaves a 16 x 16 matrix with the same non-zero eigenvalues.  Q is
// orthogonal, so the deflation is backward stable (it perturbs M2 by rounding errors only); no rank decision is taken
// from the data.  Layout of T: index p < 6 <-> m'_p, index 6 + i <-> (mu m')_i.
template <class Mat>
SSFM_HD_NOINLINE bool companion16(const double Fb[3][9], const double M[3][10][10], Mat& T) {
  const double fa = Fb[0][8], fb = Fb[1][8], fc = Fb[2][8];
  // B^T, 10 x 6: column p = coefficients of F22 * (p-th quadratic monomial) in the cubic monomial basis
  double Bt[10][6];
  for (int i = 0; i < 10; ++i)
    for (int p = 0; p < 6; ++p) Bt[i][p] = 0.0;
  Bt[0][0] = fa; Bt[1][0] = fb; Bt[4][0] = fc;
  Bt[1][1] = fa; Bt[2][1] = fb; Bt[5][1] = fc;
  Bt[2][2] = fa; Bt[3][2] = fb; Bt[6][2] = fc;
  Bt[4][3] = fa; Bt[5][3] = fb; Bt[7][3] = fc;
  Bt[5][4] = fa; Bt[6][4] = fb; Bt[8][4] = fc;
  Bt[7][5] = fa; Bt[8][5] = fb; Bt[9][5] = fc;
  double V[6][10], beta[6];  // Householder vectors (entries k..9 of V[k]) and 2 / |v|^2
  for (int k = 0; k < 6; ++k) {
    double nrm = 0.0;
    for (int i = k; i < 10; ++i) nrm += Bt[i][k] * Bt[i][k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) return false;
    const double alpha = Bt[k][k] > 0 ? -nrm : nrm;
    for (int i = 0; i < k; ++i) V[k][i] = 0.0;
    V[k][k] = Bt[k][k] - alpha;
    double vn = V[k][k] * V[k][k];
    for (int i = k + 1; i < 10; ++i) { V[k][i] = Bt[i][k]; vn += V[k][i] * V[k][i]; }
    beta[k] = vn > 0.0 ? 2.0 / vn : 0.0;
    for (int j = k + 1; j < 6; ++j) {
      double d = 0.0;
      for (int i = k; i < 10; ++i) d += V[k][i] * Bt[i][j];
      d *= beta[k];
      for (int i = k; i < 10; ++i) Bt[i][j] -= d * V[k][i];
    }
  }
  // rows of Mk Q = ((row H0) H1) ... H5
  double L[10][10];
  for (int which = 0; which < 3; ++which)
    for (int e = 0; e < 10; ++e) {
      double row[10];
      for (int q = 0; q < 10; ++q) row[q] = M[which][e][q];
      for (int k = 0; k < 6; ++k) {
        double d = 0.0;
        for (int i = k; i < 10; ++i) d += row[i] * V[k][i];
        d *= beta[k];
        for (int i = k; i < 10; ++i) row[i] -= d * V[k][i];
      }
      if (which == 0) {
        for (int q = 0; q < 10; ++q) L[e][q] = row[q];
      } else if (which == 1) {
        for (int q = 0; q < 10; ++q) T(6 + e, 6 + q) = -row[q];
      } else {
        for (int q = 0; q < 6; ++q) T(6 + e, q) = -row[q];  // columns 6..9 of M2 Q vanish (rounding only)
      }
    }
  // rows 6..15 of T are the right-hand sides X = [-A2[:, :6] | -A1], solved in place: Gaussian elimination with partial
  // pivoting on [A0 | X]
  for (int k = 0; k < 10; ++k) {
    int piv = k;
    double best = fabs(L[k][k]);
    for (int i = k + 1; i < 10; ++i)
      if (fabs(L[i][k]) > best) { best = fabs(L[i][k]); piv = i; }
    if (!(best > 1e-300)) return false;
    if (piv != k) {
      for (int j = 0; j < 10; ++j) { const double tt = L[k][j]; L[k][j] = L[piv][j]; L[piv][j] = tt; }
      for (int j = 0; j < kN; ++j) { const double tt = T(6 + k, j); T(6 + k, j) = T(6 + piv, j); T(6 + piv, j) = tt; }
    }
    const double inv = 1.0 / L[k][k];
    for (int i = k + 1; i < 10; ++i) {
      const double f = L[i][k] * inv;
      if (f != 0.0) {
        for (int j = k + 1; j < 10; ++j) L[i][j] -= f * L[k][j];
        for (int j = 0; j < kN; ++j) T(6 + i, j) -= f * T(6 + k, j);
      }
    }
  }
  for (int j = 0; j < kN; ++j)
    for (int i = 9; i >= 0; --i) {
      double v = T(6 + i, j);
      for (int q = i + 1; q < 10; ++q) v -= L[i][q] * T(6 + q, j);
      T(6 + i, j) = v / L[i][i];
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < kN; ++j) T(i, j) = (j == 6 + i) ? 1.0 : 0.0;
  for (int i = 6; i < kN; ++i)
    for (int j = 0; j < kN; ++j)
      if (!(fabs(T(i, j)) < 1e300)) return false;
  return true;
}

// min || A x - b ||, A rows x cols (cols <= 9, rows = 10), by Householder QR.  A and b are destroyed.
template <int COLS>
SSFM_HD_NOINLINE bool least_squares10(double (*A)[COLS], double* b, double* x) {
  constexpr int ROWS = 10;
  for (int k = 0; k < COLS; ++k) {
    double nrm = 0.0;
    for (int i = k; i < ROWS; ++i) nrm += A[i][k] * A[i][k];
    nrm = sqrt(nrm);
    if (!(nrm > 0.0)) return false;
    const double alpha = A[k][k] > 0 ? -nrm : nrm;
    const double v0 = A[k][k] - alpha;
    double vn = v0 * v0;
    for (int i = k + 1; i < ROWS; ++i) vn += A[i][k] * A[i][k];
    if (vn > 0.0) {
      const double beta = 2.0 / vn;
      for (int j = k + 1; j < COLS; ++j) {
        double d = v0 * A[k][j];
        for (int i = k + 1; i < ROWS; ++i) d += A[i][k] * A[i][j];
        d *= beta;
        A[k][j] -= d * v0;
        for (int i = k + 1; i < ROWS; ++i) A[i][j] -= d * A[i][k];
      }
      double d = v0 * b[k];
      for (int i = k + 1; i < ROWS; ++i) d += A[i][k] * b[i];
      d *= beta;
      b[k] -= d * v0;
      for (int i = k + 1; i < ROWS; ++i) b[i] -= d * A[i][k];
    }
    A[k][k] = alpha;
  }
  for (int k = COLS - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < COLS; ++j) s -= A[k][j] * x[j];
    x[k] = s / A[k][k];
  }
  return true;
}

SSFM_HD void monomials(double x, double y, double* m, double* dx, double* dy) {
  const double x2 = x * x, y2 = y * y;
  m[0] = x2 * x; m[1] = x2 * y; m[2] = x * y2; m[3] = y2 * y; m[4] = x2; m[5] = x * y; m[6] = y2; m[7] = x; m[8] = y; m[9] = 1.0;
  dx[0] = 3 * x2; dx[1] = 2 * x * y; dx[2] = y2; dx[3] = 0; dx[4] = 2 * x; dx[5] = y; dx[6] = 0; dx[7] = 1; dx[8] = 0; dx[9] = 0;
  dy[0] = 0; dy[1] = x2; dy[2] = 2 * x * y; dy[3] = 3 * y2; dy[4] = 0; dy[5] = x; dy[6] = 2 * y; dy[7] = 0; dy[8] = 1; dy[9] = 0;
}

// PoseLib's check_cheirality for unit bearings: both depths positive.
SSFM_HD bool in_front(const double* R, const double* t, const double* a1, const double* a2) {
  const double Rx[3] = {R[0] * a1[0] + R[1] * a1[1] + R[2] * a1[2], R[3] * a1[0] + R[4] * a1[1] + R[5] * a1[2],
                        R[6] * a1[0] + R[7] * a1[1] + R[8] * a1[2]};
  const double a = -(Rx[0] * a2[0] + Rx[1] * a2[1] + Rx[2] * a2[2]);
  const double b1 = -(Rx[0] * t[0] + Rx[1] * t[1] + Rx[2] * t[2]);
  const double b2 = a2[0] * t[0] + a2[1] * t[1] + a2[2] * t[2];
  return (b1 - a * b2 > 0.0) && (-a * b1 + b2 > 0.0);
}

}  // namespace sixpt

// rays: six correspondences, each (u.xyz, v.xyz).  Returns the number of models written (sorted by focal).  T: where this
// instance keeps its kN x kN matrix (sixpt::LocalMat / sixpt::StridedMat).
template <class Mat>
SSFM_HD_NOINLINE int solve_sixpt_focal_in(const double (*c)[6], SixPointModel* out, Mat& T) {
  using namespace sixpt;
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
  if (!nullspace_6x9(x1, x2, Fb)) return 0;
  double M[3][10][10];
  constraint_matrices(Fb, M);
  if (!companion16(Fb, M, T)) return 0;
  balance(T);
  to_hessenberg(T);
  double wr[kN], wi[kN];
  if (!hessenberg_eigenvalues(T, wr, wi)) return 0;

  int n_out = 0;
  double sx[kSixMaxModels], sy[kSixMaxModels], sw[kSixMaxModels];
  int n_sol = 0;
  for (int k = 0; k < kN; ++k) {
    const double mu = wr[k];
    if (!(mu > 0.0) || !(mu < 1e300) || fabs(wi[k]) > 1e-4 * mu) continue;  // nearly real: the polish + residual test decide
    // A nearly-real conjugate pair may be two close real roots: start the polish on either side.
    double w = 1.0 / (mu + wi[k]);
    if (!(w > 0.0)) w = 1.0 / mu;
    // (x, y) at this w: least squares for the nine non-constant monomials
    double x, y;
    {
      double A[10][9], b[10], sol[9];
      for (int e = 0; e < 10; ++e) {
        for (int q = 0; q < 9; ++q) A[e][q] = M[0][e][q] + w * (M[1][e][q] + w * M[2][e][q]);
        b[e] = -(M[0][e][9] + w * (M[1][e][9] + w * M[2][e][9]));
      }
      if (!least_squares10<9>(A, b, sol)) continue;
      x = sol[7];
      y = sol[8];
    }
    bool ok = true, converged = false;
    for (int itn = 0; itn < 12 && ok && !converged; ++itn) {  // Gauss-Newton on the ten equations in (x, y, w)
      double m[10], dx[10], dy[10], J[10][3], res[10], step[3];
      monomials(x, y, m, dx, dy);
      for (int e = 0; e < 10; ++e) {
        double r0 = 0, jx = 0, jy = 0, jw = 0;
        for (int q = 0; q < 10; ++q) {
          const double mw = M[0][e][q] + w * (M[1][e][q] + w * M[2][e][q]);
          r0 += mw * m[q];
          jx += mw * dx[q];
          jy += mw * dy[q];
          jw += (M[1][e][q] + 2.0 * w * M[2][e][q]) * m[q];
        }
        res[e] = -r0; J[e][0] = jx; J[e][1] = jy; J[e][2] = jw;
      }
      if (!least_squares10<3>(J, res, step)) { ok = false; break; }
      x += step[0]; y += step[1]; w += step[2];
      if (!(fabs(x) < 1e300) || !(fabs(y) < 1e300) || !(fabs(w) < 1e300)) ok = false;
      // accepted only once the iteration has settled (a start that wanders is a spurious eigenvalue)
      converged = fabs(step[0]) <= 1e-12 * (1.0 + fabs(x)) && fabs(step[1]) <= 1e-12 * (1.0 + fabs(y)) &&
                  fabs(step[2]) <= 1e-12 * fabs(w);
    }
    ok = ok && converged;
    if (!ok || !(w > 0.0)) continue;
    {  // residual test against the scale of each equation
      double m[10], dx[10], dy[10];
      monomials(x, y, m, dx, dy);
      for (int e = 0; e < 10 && ok; ++e) {
        double r0 = 0, sc = 0;
        for (int q = 0; q < 10; ++q) {
          const double mw = M[0][e][q] + w * (M[1][e][q] + w * M[2][e][q]);
          r0 += mw * m[q];
          sc += fabs(mw) * fabs(m[q]);
        }
        if (fabs(r0) > 1e-8 * sc) ok = false;
      }
    }
    if (!ok) continue;
    for (int q = 0; q < n_sol; ++q)
      if (fabs(x - sx[q]) + fabs(y - sy[q]) < 1e-6 * (1.0 + fabs(x) + fabs(y)) && fabs(w - sw[q]) < 1e-6 * w) ok = false;
    if (!ok || n_sol >= kSixMaxModels) continue;
    sx[n_sol] = x; sy[n_sol] = y; sw[n_sol] = w;
    ++n_sol;
  }
  for (int q = 0; q < n_sol; ++q) {
    const double f = 1.0 / sqrt(sw[q]);
    double E[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        const double Fv = sx[q] * Fb[0][3 * r + cc] + sy[q] * Fb[1][3 * r + cc] + Fb[2][3 * r + cc];
        E[3 * r + cc] = Fv * (r < 2 ? f : 1.0) * (cc < 2 ? f : 1.0);
      }
    double U[9], sv[3], V[9];
    svd3(E, U, sv, V);
    if (mat3_det(U) < 0) { U[2] = -U[2]; U[5] = -U[5]; U[8] = -U[8]; }
    if (mat3_det(V) < 0) { V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8]; }
    // unit bearings of the sample in both cameras
    double b1[6][3], b2[6][3];
    for (int i = 0; i < 6; ++i) {
      double a[3] = {x1[i][0] / f, x1[i][1] / f, x1[i][2]}, b[3] = {x2[i][0] / f, x2[i][1] / f, x2[i][2]};
      const double na = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = 1.0 / sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
      for (int d = 0; d < 3; ++d) { b1[i][d] = a[d] * na; b2[i][d] = b[d] * nb; }
    }
    const double Wm[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double Vt[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) Vt[3 * r + cc] = V[3 * cc + r];
    for (int which = 0; which < 2; ++which) {
      double tmp[9], R[9];
      mat3_mul(U, which == 0 ? Wm : Wt, tmp);
      mat3_mul(tmp, Vt, R);
      for (int sg = 0; sg < 2; ++sg) {
        const double t[3] = {sg == 0 ? U[2] : -U[2], sg == 0 ? U[5] : -U[5], sg == 0 ? U[8] : -U[8]};
        bool front = true;
        for (int i = 0; i < 6 && front; ++i) front = in_front(R, t, b1[i], b2[i]);
        if (!front || n_out >= kSixMaxModels) continue;
        SixPointModel mdl;
        for (int d = 0; d < 3; ++d) mdl.t[d] = t[d];
        so3ln(R, mdl.r);
        mdl.f = f * s;
        int pos = n_out;  // insertion sort by focal
        while (pos > 0 && out[pos - 1].f > mdl.f) { out[pos] = out[pos - 1]; --pos; }
        out[pos] = mdl;
        ++n_out;
      }
    }
  }
  return n_out;
}

SSFM_HD_NOINLINE int solve_sixpt_focal(const double (*c)[6], SixPointModel* out) {
  sixpt::LocalMat T;
  return solve_sixpt_focal_in(c, out, T);
}

// The matrix the estimator scores with.  focal_scoring == 0: E = skew3(t) so3exp(r), evaluated on the raw rays,
// exactly what SixPointEstimator::EvaluateModelOnPoint does (six_point_estimator.cpp:78-91 -- the focal is not
// used there).  focal_scoring != 0: F = Kinv E Kinv with Kinv = diag(1,1,focal), the model of the reference's own
// refit functor (:62-70), which is the geometrically meaningful error for pixel-unit rays.
SSFM_HD void sixpt_scoring_matrix(const SixPointModel& m, int focal_scoring, double* G) {
  double R[9];
  so3exp(m.r, R);
  const double S[9] = {0, -m.t[2], m.t[1], m.t[2], 0, -m.t[0], -m.t[1], m.t[0], 0};
  mat3_mul(S, R, G);
  if (focal_scoring) {
    G[2] *= m.f; G[5] *= m.f; G[6] *= m.f; G[7] *= m.f; G[8] *= m.f * m.f;
  }
}

// ---------------------------------------------------------------------------------------------
// SixPointEstimator::LeastSquares (examples/six_point_estimator.cpp:146-192): Ceres trust-region LM over
// r1 (3), t1 on the unit sphere (ceres::SphereManifold<3>, 2 tangent dimensions) and the focal (1), residual =
// the SampsonError functor (:25-76) with r0 = t0 = 0: E = [t1]x R(r1), F = Kinv E Kinv, Kinv = diag(1,1,focal).
// Ceres is absent from the reference tree: Solver::Options defaults and the SphereManifold Plus / PlusJacobian
// (Householder form) are restated from its documentation, as in the 3-point refit.
// ---------------------------------------------------------------------------------------------
namespace sixpt {

// Householder vector of ceres::internal::ComputeHouseholderVector for a 3-vector: H = I - beta v v^T maps x to |x| e3.
SSFM_HD void householder3(const double* x, double* v, double* beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1];
  v[0] = x[0]; v[1] = x[1]; v[2] = 1.0;
  *beta = 0.0;
  const double xp = x[2];
  if (sigma <= 2.220446049250313e-16) {
    if (xp < 0.0) *beta = 2.0;
    return;
  }
  const double mu = sqrt(xp * xp + sigma);
  const double vp = xp <= 0.0 ? xp - mu : -sigma / (xp + mu);
  *beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp; v[1] /= vp;
}
// SphereManifold<3>::Plus: x_plus = |x| H [sin(|d|) d / |d| ; cos(|d|)]
SSFM_HD void sphere_plus(const double* x, const double* d, double* out) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1]);
  if (nd == 0.0) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; return; }
  double v[3], beta;
  householder3(x, v, &beta);
  const double sbd = sin(nd) / nd;
  const double y[3] = {sbd * d[0], sbd * d[1], cos(nd)};
  const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double vy = beta * (v[0] * y[0] + v[1] * y[1] + v[2] * y[2]);
  for (int i = 0; i < 3; ++i) out[i] = nx * (y[i] - v[i] * vy);
}
// SphereManifold<3>::PlusJacobian (3 x 2): |x| times the first two columns of H
SSFM_HD void sphere_plus_jacobian(const double* x, double (*J)[2]) {
  double v[3], beta;
  householder3(x, v, &beta);
  const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 2; ++j) J[i][j] = nx * ((i == j ? 1.0 : 0.0) - beta * v[i] * v[j]);
}

}  // namespace sixpt

struct SixLmSummary {
  int iterations;
  double initial_cost, final_cost;
};

// One-lane execution context (the interface of SerialCtx / WarpCtx, ssfm_chain.cuh, as far as this file needs it).
struct SixOneLane {
  SSFM_HD int lane() const { return 0; }
  SSFM_HD int width() const { return 1; }
  SSFM_HD double sum(double x) const { return x; }
  template <int K>
  SSFM_HD void sum_vec(double (&v)[K]) const { (void)v; }
};

// rays: the estimator's correspondences (6 doubles each); sample: indices of the residuals.  With a multi-lane context the
// residuals (and their contributions to the normal equations) are spread over the lanes and reduced; everything else is
// computed redundantly by every lane from identical inputs, so control flow stays uniform.
template <class Ctx>
SSFM_HD_NOINLINE SixLmSummary sixpt_least_squares(const Ctx& cx, const double* rays, const int* sample, int n, SixPointModel& model) {
  using namespace sixpt;
  double x[7] = {model.r[0], model.r[1], model.r[2], model.t[0], model.t[1], model.t[2], model.f};
  double H[21], g[6], scale[6], diagonal[6] = {0, 0, 0, 0, 0, 0}, gmax = 0.0;
  bool have_scale = false;
  // residuals (the functor returns d^2/den itself as the residual, :72) and the local Jacobian at xx
  auto evaluate = [&](const double* xx, bool with_jac) -> double {
    double c = 0.0;
    Jet6 Ej[9];
    double Pt[3][2];
    if (with_jac) {
      const Jet6 r1[3] = {jvar(xx[0], 0), jvar(xx[1], 1), jvar(xx[2], 2)};
      const Jet6 t1[3] = {jvar(xx[3], 3), jvar(xx[4], 4), jvar(xx[5], 5)};
      spherical_E_of_params<Jet6>(r1, t1, 0.0, Ej);
      sphere_plus_jacobian(xx + 3, Pt);
      for (int a = 0; a < 21; ++a) H[a] = 0.0;
      for (int a = 0; a < 6; ++a) g[a] = 0.0;
    } else {
      double Ed[9];
      spherical_E_of_params<double>(xx, xx + 3, 0.0, Ed);
      for (int q = 0; q < 9; ++q) Ej[q].a = Ed[q];
    }
    const double f = xx[6];
    const double S[9] = {1, 1, f, 1, 1, f, f, f, f * f}, dS[9] = {0, 0, 1, 0, 0, 1, 1, 1, 2 * f};
    double F[9];
    for (int q = 0; q < 9; ++q) F[q] = Ej[q].a * S[q];
    for (int k = cx.lane(); k < n; k += cx.width()) {
      const double* u = rays + 6 * (size_t)sample[k];
      const double* v = u + 3;
      const double Fu0 = F[0] * u[0] + F[1] * u[1] + F[2] * u[2];
      const double Fu1 = F[3] * u[0] + F[4] * u[1] + F[5] * u[2];
      const double Fu2 = F[6] * u[0] + F[7] * u[1] + F[8] * u[2];
      const double Ft0 = F[0] * v[0] + F[3] * v[1] + F[6] * v[2];
      const double Ft1 = F[1] * v[0] + F[4] * v[1] + F[7] * v[2];
      const double d = v[0] * Fu0 + v[1] * Fu1 + v[2] * Fu2;
      const double den = Fu0 * Fu0 + Fu1 * Fu1 + Ft0 * Ft0 + Ft1 * Ft1;
      const double inv = 1.0 / den;
      const double r = d * d * inv;
      c += r * r;
      if (!with_jac) continue;
      const double a2 = 2.0 * d * inv, b2 = 2.0 * r * inv;
      double w[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double dd = 0.0;
          if (i < 2) dd += (i == 0 ? Fu0 : Fu1) * u[j];
          if (j < 2) dd += (j == 0 ? Ft0 : Ft1) * v[i];
          w[3 * i + j] = a2 * v[i] * u[j] - b2 * dd;  // d residual / d F_ij
        }
      double ga[7];
      for (int p7 = 0; p7 < 6; ++p7) {
        double sacc = 0.0;
        for (int q = 0; q < 9; ++q) sacc += w[q] * Ej[q].v[p7] * S[q];
        ga[p7] = sacc;
      }
      ga[6] = 0.0;
      for (int q = 0; q < 9; ++q) ga[6] += w[q] * Ej[q].a * dS[q];
      const double jl[6] = {ga[0], ga[1], ga[2], ga[3] * Pt[0][0] + ga[4] * Pt[1][0] + ga[5] * Pt[2][0],
                            ga[3] * Pt[0][1] + ga[4] * Pt[1][1] + ga[5] * Pt[2][1], ga[6]};
      int hk = 0;
      for (int a = 0; a < 6; ++a) {
        g[a] += jl[a] * r;
        for (int b = 0; b <= a; ++b) H[hk++] += jl[a] * jl[b];
      }
    }
    c = cx.sum(c);
    if (with_jac) {
      cx.sum_vec(H);
      cx.sum_vec(g);
      gmax = 0.0;
      for (int a = 0; a < 6; ++a) gmax = fmax(gmax, fabs(g[a]));  // gradient of the unscaled problem
      if (!have_scale) {
        for (int a = 0; a < 6; ++a) scale[a] = 1.0 / (1.0 + sqrt(H[a * (a + 1) / 2 + a]));
        have_scale = true;
      }
      int hk = 0;
      for (int a = 0; a < 6; ++a) {
        g[a] *= scale[a];
        for (int b = 0; b <= a; ++b) H[hk++] *= scale[a] * scale[b];
      }
    }
    return 0.5 * c;
  };
  SixLmSummary sum;
  double x_cost = evaluate(x, true);
  sum.iterations = 0;
  sum.initial_cost = sum.final_cost = x_cost;
  if (!isfinite(x_cost)) return sum;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid = 0, iteration = 0;
  for (;;) {
    if (iteration >= 200) break;
    if (gmax <= 1e-10) break;
    if (radius < 1e-32) break;
    ++iteration;
    if (!reuse_diagonal)
      for (int a = 0; a < 6; ++a) diagonal[a] = fmin(fmax(H[a * (a + 1) / 2 + a], 1e-6), 1e32);
    double Hd[21], step[6];
    for (int a = 0; a < 21; ++a) Hd[a] = H[a];
    for (int a = 0; a < 6; ++a) Hd[a * (a + 1) / 2 + a] += diagonal[a] / radius;
    bool valid = cholesky_solve6(Hd, g, step);
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (valid) {
      double sg = 0.0, sHs = 0.0;
      int hk = 0;
      for (int a = 0; a < 6; ++a) {
        step[a] = -step[a];
        sg += step[a] * g[a];
      }
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b <= a; ++b) sHs += (a == b ? 1.0 : 2.0) * H[hk++] * step[a] * step[b];
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
    double delta[6], cand[7];
    for (int a = 0; a < 6; ++a) delta[a] = step[a] * scale[a];
    cand[0] = x[0] + delta[0]; cand[1] = x[1] + delta[1]; cand[2] = x[2] + delta[2];
    sphere_plus(x + 3, delta + 3, cand + 3);
    cand[6] = x[6] + delta[5];
    double step_norm = 0.0, x_norm = 0.0;
    for (int a = 0; a < 7; ++a) {
      step_norm += (cand[a] - x[a]) * (cand[a] - x[a]);
      x_norm += x[a] * x[a];
    }
    step_norm = sqrt(step_norm);
    x_norm = sqrt(x_norm);
    double cand_cost = evaluate(cand, false);
    if (!isfinite(cand_cost)) cand_cost = kDblMax;
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) break;
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= 1e-6 * x_cost) break;
    const double rho = cost_change / model_cost_change;
    if (rho > 1e-3) {
      for (int a = 0; a < 7; ++a) x[a] = cand[a];
      x_cost = evaluate(x, true);
      const double tt = 2.0 * rho - 1.0;
      radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - tt * tt * tt));
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  for (int i = 0; i < 3; ++i) { model.r[i] = x[i]; model.t[i] = x[3 + i]; }
  model.f = x[6];
  return sum;
}
SSFM_HD SixLmSummary sixpt_least_squares(const double* rays, const int* sample, int n, SixPointModel& model) {
  return sixpt_least_squares(SixOneLane(), rays, sample, n, model);
}

}  // namespace ssfm
