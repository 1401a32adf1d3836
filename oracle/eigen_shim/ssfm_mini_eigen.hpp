// ssfm_mini_eigen.hpp -- a tiny stand-in for the subset of Eigen that the reference's
// src/spherical_solvers.cpp, src/so3.cpp and src/spherical_utils.cpp use, so those files can be
// compiled UNMODIFIED, where they lie under /root/reference, into oracle/_ref (Eigen itself is not
// installed and there is no network).  TEST INFRASTRUCTURE ONLY; this is not Eigen and shares no code
// with it.  Everything is eager (no expression templates), row-major, heap-backed.
//
// What it reproduces faithfully: column-pivoted Householder QR with Eigen's pivot rule and
// makeHouseholder sign convention (so the null-space basis B, and with it the polynomial variant's
// root set, follows the reference), partial-pivot LU, a 3x3 SVD.  EigenSolver is Eigen 3.4's algorithm
// restated step by step (ssfm_oracle::eigen34_eigensolver_4x4: Householder Hessenberg reduction, Francis
// QR with Eigen's shift/deflation rules, hqr2 back-substitution, unit-norm complex columns), because the
// reference keeps the real part of COMPLEX eigenvectors, whose phase is fixed by that iteration history.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <type_traits>
#include <vector>

#include "../ssfm_oracle.hpp"

#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

namespace Eigen {

const int Dynamic = -1;
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <class T, int R, int C>
class Matrix;

template <class T>
struct BlockRef;

template <class M, class T>
struct CommaInit {
  M& m;
  int k;
  CommaInit(M& mm, T first) : m(mm), k(0) { put(first); }
  void put(T v) {
    m.at_flat(k) = v;
    ++k;
  }
  CommaInit& operator,(T v) {
    put(v);
    return *this;
  }
};

template <class T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  int r, c;
  std::vector<T> d;
  Matrix() : r(R > 0 ? R : 0), c(C > 0 ? C : 0), d((size_t)r * c, T(0)) {}
  Matrix(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, T(0)) {}
  Matrix(T a, T b, T cc) : r(3), c(1), d{a, b, cc} {}
  Matrix(T a, T b) : r(2), c(1), d{a, b} {}  // Vector2d(x, y); integer arguments select (rows, cols) above
  template <int R2, int C2>
  Matrix(const Matrix<T, R2, C2>& o) : r(o.r), c(o.c), d(o.d) {}
  Matrix(const BlockRef<T>& b);
  template <int R2, int C2>
  Matrix& operator=(const Matrix<T, R2, C2>& o) {
    r = o.r; c = o.c; d = o.d;
    return *this;
  }
  Matrix& operator=(const BlockRef<T>& b);
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < std::min(m.r, m.c); ++i) m(i, i) = T(1);
    return m;
  }
  int rows() const { return r; }
  int cols() const { return c; }
  int size() const { return r * c; }
  T& operator()(int i, int j) { return d[(size_t)i * c + j]; }
  const T& operator()(int i, int j) const { return d[(size_t)i * c + j]; }
  T& operator()(int i) { return d[i]; }
  const T& operator()(int i) const { return d[i]; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
  T& at_flat(int k) { return d[k]; }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
  CommaInit<Matrix, T> operator<<(T v) { return CommaInit<Matrix, T>(*this, v); }
  BlockRef<T> block(int r0, int c0, int nr, int nc);
  Matrix<T, Dynamic, Dynamic> block(int r0, int c0, int nr, int nc) const;
  template <int NR, int NC>
  BlockRef<T> block(int r0, int c0) { return block(r0, c0, NR, NC); }
  template <int NR, int NC>
  Matrix<T, NR, NC> block(int r0, int c0) const { return Matrix<T, NR, NC>(block(r0, c0, NR, NC)); }
  BlockRef<T> row(int i);
  Matrix<T, Dynamic, Dynamic> row(int i) const { return block(i, 0, 1, c); }
  BlockRef<T> col(int j);
  Matrix<T, Dynamic, 1> col(int j) const { return Matrix<T, Dynamic, 1>(block(0, j, r, 1)); }
  BlockRef<T> head(int n);
  Matrix<T, Dynamic, 1> head(int n) const { return Matrix<T, Dynamic, 1>(block(0, 0, n, 1)); }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j);
    return t;
  }
  const Matrix& eval() const { return *this; }
  T norm() const { return std::sqrt(squaredNorm()); }
  T squaredNorm() const {
    T s = 0;
    for (const T& v : d) s += v * v;
    return s;
  }
  template <int R2, int C2>
  T dot(const Matrix<T, R2, C2>& o) const {
    T s = 0;
    for (size_t i = 0; i < d.size(); ++i) s += d[i] * o.d[i];
    return s;
  }
  T determinant() const {
    assert(r == 3 && c == 3);
    const Matrix& a = *this;
    return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
           a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
  }
  Matrix operator-() const {
    Matrix m(*this);
    for (T& v : m.d) v = -v;
    return m;
  }
  Matrix& operator*=(T s) {
    for (T& v : d) v *= s;
    return *this;
  }
  Matrix& operator/=(T s) {
    for (T& v : d) v /= s;
    return *this;
  }
  template <int R2, int C2>
  Matrix& operator+=(const Matrix<T, R2, C2>& o) {
    for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i];
    return *this;
  }
  struct ColPivQR;
  struct PartialLU;
  struct Svd;
  ColPivQR colPivHouseholderQr() const;
  PartialLU lu() const;
  Svd jacobiSvd(unsigned int opts = 0) const;
};

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;

template <class T>
struct BlockRef {
  std::vector<T>* d;
  int stride, r0, c0, nr, nc;
  T& ref(int i, int j) const { return (*d)[(size_t)(r0 + i) * stride + c0 + j]; }
  T& at_flat(int k) { return ref(k / nc, k % nc); }
  Matrix<T, Dynamic, Dynamic> eval() const {
    Matrix<T, Dynamic, Dynamic> m(nr, nc);
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) m(i, j) = ref(i, j);
    return m;
  }
  template <int R, int C>
  BlockRef& operator=(const Matrix<T, R, C>& m) {
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) ref(i, j) = m.d[(size_t)i * nc + j];
    return *this;
  }
  BlockRef& operator=(const BlockRef& o) { return *this = o.eval(); }
  CommaInit<BlockRef, T> operator<<(T v) { return CommaInit<BlockRef, T>(*this, v); }
  Matrix<T, Dynamic, Dynamic> operator-() const { return -eval(); }
  BlockRef& operator*=(T s) {
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) ref(i, j) *= s;
    return *this;
  }
  BlockRef& operator+=(const Matrix<T, Dynamic, Dynamic>& m) {
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) ref(i, j) += m(i, j);
    return *this;
  }
  Matrix<T, Dynamic, Dynamic> transpose() const { return eval().transpose(); }
  typename Matrix<T, Dynamic, Dynamic>::PartialLU lu() const;
  T squaredNorm() const { return eval().squaredNorm(); }
  T norm() const { return eval().norm(); }
  Matrix<T, Dynamic, Dynamic> operator/(T s) const {
    Matrix<T, Dynamic, Dynamic> m = eval();
    for (T& v : m.d) v /= s;
    return m;
  }
  Matrix<T, Dynamic, Dynamic> operator*(T s) const {
    Matrix<T, Dynamic, Dynamic> m = eval();
    for (T& v : m.d) v *= s;
    return m;
  }
};

template <class T, int R, int C>
Matrix<T, R, C>::Matrix(const BlockRef<T>& b) : r(b.nr), c(b.nc), d((size_t)b.nr * b.nc) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) (*this)(i, j) = b.ref(i, j);
}
template <class T, int R, int C>
Matrix<T, R, C>& Matrix<T, R, C>::operator=(const BlockRef<T>& b) {
  return *this = Matrix<T, R, C>(b);
}
template <class T, int R, int C>
BlockRef<T> Matrix<T, R, C>::block(int r0, int c0, int nr, int nc) {
  return BlockRef<T>{&d, c, r0, c0, nr, nc};
}
template <class T, int R, int C>
Matrix<T, Dynamic, Dynamic> Matrix<T, R, C>::block(int r0, int c0, int nr, int nc) const {
  Matrix<T, Dynamic, Dynamic> m(nr, nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) m(i, j) = (*this)(r0 + i, c0 + j);
  return m;
}
template <class T, int R, int C>
BlockRef<T> Matrix<T, R, C>::row(int i) { return block(i, 0, 1, c); }
template <class T, int R, int C>
BlockRef<T> Matrix<T, R, C>::col(int j) { return block(0, j, r, 1); }
template <class T, int R, int C>
BlockRef<T> Matrix<T, R, C>::head(int n) { return block(0, 0, n, 1); }

// ---- Map<const Matrix<T,R,C>>(ptr): read-only view of COLUMN-major memory (Eigen's default order), here a copy ----
template <class M>
class Map : public std::remove_const<M>::type {
 public:
  typedef typename std::remove_const<M>::type Base;
  typedef typename Base::Scalar T;
  explicit Map(const T* p) : Base() {
    for (int j = 0; j < this->c; ++j)
      for (int i = 0; i < this->r; ++i) (*this)(i, j) = p[(size_t)j * this->r + i];
  }
};

// ---- arithmetic (result dimensions are dynamic; converting constructors restore fixed types) ----
template <class T, int R, int C, int R2, int C2>
Matrix<T, R, C2> operator*(const Matrix<T, R, C>& a, const Matrix<T, R2, C2>& b) {
  Matrix<T, R, C2> m(a.r, b.c);
  for (int i = 0; i < a.r; ++i)
    for (int j = 0; j < b.c; ++j) {
      T s = 0;
      for (int k = 0; k < a.c; ++k) s += a(i, k) * b(k, j);  // left-to-right, like Eigen's coeff-based product
      m(i, j) = s;
    }
  return m;
}
template <class T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& a) { Matrix<T, R, C> m(a); m *= s; return m; }
template <class T, int R, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, T s) { Matrix<T, R, C> m(a); m *= s; return m; }
template <class T, int R, int C>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, T s) { Matrix<T, R, C> m(a); m /= s; return m; }
template <class T, int R, int C, int R2, int C2>
Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R2, C2>& b) {
  Matrix<T, R, C> m(a);
  for (size_t i = 0; i < m.d.size(); ++i) m.d[i] += b.d[i];
  return m;
}
template <class T, int R, int C, int R2, int C2>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R2, C2>& b) {
  Matrix<T, R, C> m(a);
  for (size_t i = 0; i < m.d.size(); ++i) m.d[i] -= b.d[i];
  return m;
}
template <class T, int R, int C>
Matrix<T, R, C> operator-(const BlockRef<T>& a, const Matrix<T, R, C>& b) { return Matrix<T, R, C>(a.eval()) - b; }
template <class T, int R, int C>
Matrix<T, R, C> operator+(const BlockRef<T>& a, const Matrix<T, R, C>& b) { return Matrix<T, R, C>(a.eval()) + b; }
template <class T, int R, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, const BlockRef<T>& b) { return a * b.eval(); }
template <class T, int R, int C>
Matrix<T, Dynamic, Dynamic> operator*(const BlockRef<T>& a, const Matrix<T, R, C>& b) { return a.eval() * b; }
template <class T>
Matrix<T, Dynamic, Dynamic> operator*(const BlockRef<T>& a, const BlockRef<T>& b) { return a.eval() * b.eval(); }

// ---- ColPivHouseholderQR (pivot = largest remaining column norm, Eigen's householder convention) ----
template <class T, int R, int C>
struct Matrix<T, R, C>::ColPivQR {
  int rows, steps;
  std::vector<std::vector<T>> vs;
  std::vector<T> taus;
  Matrix<T, Dynamic, Dynamic> householderQ() const {
    Matrix<T, Dynamic, Dynamic> Q(rows, rows);
    for (int j = 0; j < rows; ++j) {
      std::vector<T> q(rows, T(0));
      q[j] = 1;
      for (int k = steps - 1; k >= 0; --k) {
        T dd = 0;
        for (int i = 0; i < rows; ++i) dd += vs[k][i] * q[i];
        dd *= taus[k];
        for (int i = 0; i < rows; ++i) q[i] -= dd * vs[k][i];
      }
      for (int i = 0; i < rows; ++i) Q(i, j) = q[i];
    }
    return Q;
  }
};
template <class T, int R, int C>
typename Matrix<T, R, C>::ColPivQR Matrix<T, R, C>::colPivHouseholderQr() const {
  ColPivQR qr;
  Matrix<T, Dynamic, Dynamic> M(*this);
  const int m = r, n = c;
  qr.rows = m;
  qr.steps = std::min(m, n);
  for (int k = 0; k < qr.steps; ++k) {
    int piv = k;
    T best = -1;
    for (int cc = k; cc < n; ++cc) {
      T s = 0;
      for (int i = k; i < m; ++i) s += M(i, cc) * M(i, cc);
      if (s > best) { best = s; piv = cc; }
    }
    if (piv != k)
      for (int i = 0; i < m; ++i) std::swap(M(i, k), M(i, piv));
    const T c0 = M(k, k);
    T tail2 = 0;
    for (int i = k + 1; i < m; ++i) tail2 += M(i, k) * M(i, k);
    std::vector<T> v(m, T(0));
    v[k] = 1;
    T tau, beta;
    if (tail2 <= std::numeric_limits<T>::min()) {
      tau = 0;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < m; ++i) v[i] = M(i, k) / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    for (int cc = k; cc < n; ++cc) {
      T dd = 0;
      for (int i = k; i < m; ++i) dd += v[i] * M(i, cc);
      dd *= tau;
      for (int i = k; i < m; ++i) M(i, cc) -= dd * v[i];
    }
    qr.vs.push_back(v);
    qr.taus.push_back(tau);
  }
  return qr;
}

// ---- PartialPivLU ----
template <class T, int R, int C>
struct Matrix<T, R, C>::PartialLU {
  Matrix<T, Dynamic, Dynamic> a;
  std::vector<int> perm;
  template <class RHS>
  Matrix<T, Dynamic, Dynamic> solve(const RHS& rhs_in) const {
    const Matrix<T, Dynamic, Dynamic> rhs(rhs_in);
    const int n = a.r, m = rhs.c;
    Matrix<T, Dynamic, Dynamic> x(n, m);
    for (int j = 0; j < m; ++j) {
      std::vector<T> y(n);
      for (int i = 0; i < n; ++i) {
        T s = rhs(perm[i], j);
        for (int k = 0; k < i; ++k) s -= a(i, k) * y[k];
        y[i] = s;
      }
      for (int i = n - 1; i >= 0; --i) {
        T s = y[i];
        for (int k = i + 1; k < n; ++k) s -= a(i, k) * x(k, j);
        x(i, j) = s / a(i, i);
      }
    }
    return x;
  }
  Matrix<T, Dynamic, Dynamic> inverse() const {
    Matrix<T, Dynamic, Dynamic> id(a.r, a.r);
    for (int i = 0; i < a.r; ++i) id(i, i) = T(1);
    return solve(id);
  }
};
template <class T, int R, int C>
typename Matrix<T, R, C>::PartialLU Matrix<T, R, C>::lu() const {
  PartialLU f;
  f.a = Matrix<T, Dynamic, Dynamic>(*this);
  const int n = r;
  f.perm.resize(n);
  for (int i = 0; i < n; ++i) f.perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    T best = std::abs(f.a(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::abs(f.a(i, k)) > best) { best = std::abs(f.a(i, k)); piv = i; }
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(f.a(k, j), f.a(piv, j));
      std::swap(f.perm[k], f.perm[piv]);
    }
    for (int i = k + 1; i < n; ++i) {
      f.a(i, k) /= f.a(k, k);
      for (int j = k + 1; j < n; ++j) f.a(i, j) -= f.a(i, k) * f.a(k, j);
    }
  }
  return f;
}
template <class T>
typename Matrix<T, Dynamic, Dynamic>::PartialLU BlockRef<T>::lu() const { return eval().lu(); }

// ---- 3x3 SVD ----
template <class T, int R, int C>
struct Matrix<T, R, C>::Svd {
  Matrix<T, 3, 3> U, V;
  Matrix<T, 3, 1> S;
  const Matrix<T, 3, 3>& matrixU() const { return U; }
  const Matrix<T, 3, 3>& matrixV() const { return V; }
  const Matrix<T, 3, 1>& singularValues() const { return S; }
};
template <class T, int R, int C>
typename Matrix<T, R, C>::Svd Matrix<T, R, C>::jacobiSvd(unsigned int) const {
  assert(r == 3 && c == 3);
  ssfm_oracle::Mat3 A, U, V;
  double s[3];
  for (int i = 0; i < 9; ++i) A.m[i] = d[i];
  ssfm_oracle::svd3(A, U, s, V);
  Svd out;
  for (int i = 0; i < 9; ++i) { out.U.d[i] = U.m[i]; out.V.d[i] = V.m[i]; }
  for (int i = 0; i < 3; ++i) out.S.d[i] = s[i];
  return out;
}
template <class M>
class JacobiSVD {
 public:
  typename M::Svd f;
  JacobiSVD(const M& m, unsigned int opts = 0) : f(m.jacobiSvd(opts)) {}
  const Matrix<double, 3, 3>& matrixU() const { return f.U; }
  const Matrix<double, 3, 3>& matrixV() const { return f.V; }
};

// JacobiSVD of a tall dynamic matrix (src/triangulation_estimator.cpp:81: the 2N x 4 DLT matrix, ComputeFullV):
// one-sided Jacobi (Hestenes) on the columns of A; singular values descending, V orthogonal.
template <>
class JacobiSVD<MatrixXd> {
 public:
  MatrixXd V;
  VectorXd S;
  JacobiSVD(const MatrixXd& A_in, unsigned int = 0) {
    MatrixXd A(A_in);
    const int m = A.rows(), n = A.cols();
    V = MatrixXd(n, n);
    for (int i = 0; i < n; ++i) V(i, i) = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
      bool rotated = false;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          double alpha = 0, beta = 0, gamma = 0;
          for (int i = 0; i < m; ++i) { alpha += A(i, p) * A(i, p); beta += A(i, q) * A(i, q); gamma += A(i, p) * A(i, q); }
          if (gamma == 0.0 || std::fabs(gamma) <= 1e-16 * std::sqrt(alpha * beta)) continue;
          rotated = true;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
          for (int i = 0; i < m; ++i) { const double a = A(i, p), b = A(i, q); A(i, p) = c * a - s * b; A(i, q) = s * a + c * b; }
          for (int i = 0; i < n; ++i) { const double a = V(i, p), b = V(i, q); V(i, p) = c * a - s * b; V(i, q) = s * a + c * b; }
        }
      if (!rotated) break;
    }
    std::vector<double> sv(n);
    std::vector<int> order(n);
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int i = 0; i < m; ++i) s += A(i, j) * A(i, j);
      sv[j] = std::sqrt(s);
      order[j] = j;
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
    MatrixXd Vs(n, n);
    S = VectorXd(n, 1);
    for (int j = 0; j < n; ++j) {
      S(j) = sv[order[j]];
      for (int i = 0; i < n; ++i) Vs(i, j) = V(i, order[j]);
    }
    V = Vs;
  }
  const MatrixXd& matrixV() const { return V; }
  const VectorXd& singularValues() const { return S; }
};

// ---- EigenSolver (real 4x4): eigenvalues by Hessenberg-QR, unit-norm eigenvectors ----
template <class M>
class EigenSolver {
 public:
  typedef Matrix<std::complex<double>, Dynamic, Dynamic> EigenvectorsType;
  typedef Matrix<std::complex<double>, Dynamic, 1> EigenvalueType;
  explicit EigenSolver(const M& m) : vecs_(4, 4), vals_(4, 1) {
    assert(m.rows() == 4 && m.cols() == 4);
    double a[4][4];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) a[i][j] = m(i, j);
    std::complex<double> ev[4], V[4][4];
    const bool ok = ssfm_oracle::eigen34_eigensolver_4x4(a, ev, V);
    const double nanv = std::numeric_limits<double>::quiet_NaN();
    const bool skip = ssfm_oracle::eigen_shim_skip_complex() != 0;
    for (int k = 0; k < 4; ++k) {
      const bool masked = skip && ok && ev[k].imag() != 0.0;
      for (int i = 0; i < 4; ++i) vecs_(i, k) = (ok && !masked) ? V[i][k] : std::complex<double>(nanv, nanv);
      vals_(k) = ok ? ev[k] : std::complex<double>(nanv, nanv);
    }
  }
  const EigenvectorsType& eigenvectors() const { return vecs_; }
  const EigenvalueType& eigenvalues() const { return vals_; }

 private:
  EigenvectorsType vecs_;
  EigenvalueType vals_;
};

}  // namespace Eigen
