// tri_oracle.hpp -- CPU ORACLE (test infrastructure only) for SfM::Retriangulate's per-point robust triangulation
// (SURVEY.md 8f rank 3): float64 restatement of
//   sphericalsfm::TriangulationEstimator   src/triangulation_estimator.cpp:46-127,
//                                          include/sphericalsfm/triangulation_estimator.h:8-40
//   TriangulationError functor             src/triangulation_estimator.cpp:17-44 (ceres::AngleAxisRotatePoint)
//   Pose (P = [so3exp(r) | t])             src/sfm_types.cpp:14-41
//   the call site                          src/sfm.cpp:156-192 (LocallyOptimizedMSAC, RansacLib default LO options,
//                                          squared_inlier_threshold_ 4, final_least_squares_ true, >= 3 observations,
//                                          >= 3 inliers or the point stays at zero)
// driven by the generic lo_msac of lomsac.hpp (itself pinned bit-exact against the reference's RansacLib headers).
//
// Parity: PARTLY PINNED.  The driver loop is the pinned one.  The estimator's own arithmetic is restated:
// the reference takes the last right-singular vector of the 2N x 4 DLT matrix with Eigen::JacobiSVD (Eigen is
// absent here); this file takes the eigenvector of A^T A with the smallest eigenvalue (cyclic Jacobi), the same
// vector mathematically.  The refit is Ceres 2.2's trust-region LM restated (as in ssfm_oracle.hpp).  Pins that
// exist: synthetic ground truth (noise-free points recovered to 1e-9) and the reprojection error definition.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "ssfm_oracle.hpp"

namespace ssfm_oracle {

struct TriObservation {  // TriangulationObservation (triangulation_estimator.h:8-15) with Pose flattened
  double t[3], r[3];
  double R[9];  // so3exp(r), the rotation block of Pose::P (sfm_types.cpp:17)
  double x[2];
  double focal;
};

inline TriObservation make_tri_observation(const double t[3], const double r[3], const double x[2], double focal) {
  TriObservation o;
  for (int i = 0; i < 3; ++i) { o.t[i] = t[i]; o.r[i] = r[i]; }
  const Mat3 Rm = so3exp(r);
  for (int i = 0; i < 9; ++i) o.R[i] = Rm.m[i];
  o.x[0] = x[0]; o.x[1] = x[1];
  o.focal = focal;
  return o;
}

// ceres::AngleAxisRotatePoint (ceres/rotation.h)
template <typename T>
inline void angle_axis_rotate_point(const double aa[3], const T pt[3], T out[3]) {
  using std::sqrt; using std::sin; using std::cos;
  const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double theta = std::sqrt(theta2), costheta = std::cos(theta), sintheta = std::sin(theta), ti = 1.0 / theta;
    const double w[3] = {aa[0] * ti, aa[1] * ti, aa[2] * ti};
    const T wxp[3] = {T(w[1]) * pt[2] - T(w[2]) * pt[1], T(w[2]) * pt[0] - T(w[0]) * pt[2], T(w[0]) * pt[1] - T(w[1]) * pt[0]};
    const T tmp = (T(w[0]) * pt[0] + T(w[1]) * pt[1] + T(w[2]) * pt[2]) * T(1.0 - costheta);
    for (int i = 0; i < 3; ++i) out[i] = pt[i] * T(costheta) + wxp[i] * T(sintheta) + T(w[i]) * tmp;
  } else {
    const T wxp[3] = {T(aa[1]) * pt[2] - T(aa[2]) * pt[1], T(aa[2]) * pt[0] - T(aa[0]) * pt[2], T(aa[0]) * pt[1] - T(aa[1]) * pt[0]};
    for (int i = 0; i < 3; ++i) out[i] = pt[i] + wxp[i];
  }
}

// TriangulationError::operator() (triangulation_estimator.cpp:19-43): two residuals
template <typename T>
inline void triangulation_residual(const TriObservation& o, const T X[3], T res[2]) {
  T PX[3];
  angle_axis_rotate_point<T>(o.r, X, PX);
  PX[0] = PX[0] + T(o.t[0]);
  PX[1] = PX[1] + T(o.t[1]);
  PX[2] = PX[2] + T(o.t[2]);
  res[0] = T(o.focal) * (PX[0] / PX[2]) - T(o.x[0]);
  res[1] = T(o.focal) * (PX[1] / PX[2]) - T(o.x[1]);
}

// Smallest-eigenvalue eigenvector of a symmetric 4x4 matrix (cyclic Jacobi).
inline void smallest_eigenvector4(double S[4][4], double v[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) {
      diag += S[i][i] * S[i][i];
      for (int j = i + 1; j < 4; ++j) off += S[i][j] * S[i][j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (S[p][q] == 0.0) continue;
        const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double a = S[k][p], b = S[k][q];
          S[k][p] = c * a - s * b;
          S[k][q] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = S[p][k], b = S[q][k];
          S[p][k] = c * a - s * b;
          S[q][k] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = V[k][p], b = V[k][q];
          V[k][p] = c * a - s * b;
          V[k][q] = s * a + c * b;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (S[i][i] < S[best][best]) best = i;
  for (int k = 0; k < 4; ++k) v[k] = V[k][best];
}

// Ceres trust-region LM (same restatement as lm_refit in ssfm_oracle.hpp) over the 3 coordinates of the point.
inline LMSummary lm_triangulate(const TriObservation* obs, const int* sample, int n, double x[3]) {
  LMSummary sum;
  const int NP = 3, nres = 2 * n;
  typedef Jet<3> J3;
  std::vector<double> res(nres), cand_res(nres), jac((size_t)nres * NP);
  double scale[NP], gradient[NP], gmax = 0.0;
  auto eval_cost = [&](const double* xx, std::vector<double>& r) {
    double c = 0.0;
    for (int i = 0; i < n; ++i) {
      triangulation_residual<double>(obs[sample[i]], xx, &r[2 * i]);
      c += r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1];
    }
    return 0.5 * c;
  };
  auto eval_jac = [&](const double* xx, bool first) {
    double c = 0.0;
    for (int k = 0; k < NP; ++k) gradient[k] = 0.0;
    for (int i = 0; i < n; ++i) {
      const J3 X[3] = {J3(xx[0], 0), J3(xx[1], 1), J3(xx[2], 2)};
      J3 r[2];
      triangulation_residual<J3>(obs[sample[i]], X, r);
      for (int q = 0; q < 2; ++q) {
        res[2 * i + q] = r[q].a;
        c += r[q].a * r[q].a;
        for (int k = 0; k < NP; ++k) {
          jac[(size_t)(2 * i + q) * NP + k] = r[q].v[k];
          gradient[k] += r[q].v[k] * r[q].a;
        }
      }
    }
    if (first)
      for (int k = 0; k < NP; ++k) {
        double s = 0.0;
        for (int i = 0; i < nres; ++i) s += jac[(size_t)i * NP + k] * jac[(size_t)i * NP + k];
        scale[k] = 1.0 / (1.0 + std::sqrt(s));
      }
    for (int i = 0; i < nres; ++i)
      for (int k = 0; k < NP; ++k) jac[(size_t)i * NP + k] *= scale[k];
    gmax = 0.0;
    for (int k = 0; k < NP; ++k) gmax = std::max(gmax, std::fabs(gradient[k]));
    return 0.5 * c;
  };
  double x_cost = eval_jac(x, true);
  sum.initial_cost = sum.final_cost = x_cost;
  if (!std::isfinite(x_cost)) { sum.termination = 5; return sum; }
  double radius = 1e4, decrease_factor = 2.0, diagonal[NP];
  bool reuse_diagonal = false;
  int invalid = 0, iteration = 0;
  while (true) {
    if (iteration >= 200) { sum.termination = 4; break; }
    if (gmax <= 1e-10) { sum.termination = 1; break; }
    if (radius < 1e-32) { sum.termination = 6; break; }
    ++iteration;
    double H[NP][NP], g[NP], step[NP];
    for (int a = 0; a < NP; ++a) {
      g[a] = 0.0;
      for (int b = 0; b < NP; ++b) H[a][b] = 0.0;
    }
    for (int i = 0; i < nres; ++i) {
      const double* ji = &jac[(size_t)i * NP];
      for (int a = 0; a < NP; ++a) {
        g[a] += ji[a] * res[i];
        for (int b = 0; b <= a; ++b) H[a][b] += ji[a] * ji[b];
      }
    }
    if (!reuse_diagonal)
      for (int k = 0; k < NP; ++k) diagonal[k] = std::min(std::max(H[k][k], 1e-6), 1e32);
    for (int a = 0; a < NP; ++a) {
      for (int b = 0; b < a; ++b) H[b][a] = H[a][b];
    }
    double Hd[NP][NP];
    for (int a = 0; a < NP; ++a)
      for (int b = 0; b < NP; ++b) Hd[a][b] = H[a][b] + (a == b ? diagonal[a] / radius : 0.0);
    // Cholesky 3x3
    bool valid = true;
    {
      double L[NP][NP] = {};
      for (int j = 0; j < NP && valid; ++j) {
        double d = Hd[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) { valid = false; break; }
        L[j][j] = std::sqrt(d);
        for (int i = j + 1; i < NP; ++i) {
          double s = Hd[i][j];
          for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
          L[i][j] = s / L[j][j];
        }
      }
      if (valid) {
        double y[NP];
        for (int i = 0; i < NP; ++i) {
          double s = g[i];
          for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
          y[i] = s / L[i][i];
        }
        for (int i = NP - 1; i >= 0; --i) {
          double s = y[i];
          for (int k = i + 1; k < NP; ++k) s -= L[k][i] * step[k];
          step[i] = s / L[i][i];
        }
        for (int i = 0; i < NP; ++i)
          if (!std::isfinite(step[i])) valid = false;
      }
    }
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    if (valid) {
      for (int k = 0; k < NP; ++k) step[k] = -step[k];
      for (int i = 0; i < nres; ++i) {
        double m = 0.0;
        for (int k = 0; k < NP; ++k) m += jac[(size_t)i * NP + k] * step[k];
        model_cost_change -= m * (res[i] + m / 2.0);
      }
      if (!(model_cost_change > 0.0)) valid = false;
    }
    if (!valid) {
      if (++invalid >= 10) { sum.termination = 5; break; }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    double cand[NP], step_norm = 0.0, x_norm = 0.0;
    for (int k = 0; k < NP; ++k) {
      const double d = step[k] * scale[k];
      cand[k] = x[k] + d;
      step_norm += d * d;
      x_norm += x[k] * x[k];
    }
    step_norm = std::sqrt(step_norm);
    x_norm = std::sqrt(x_norm);
    double cand_cost = eval_cost(cand, cand_res);
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { sum.termination = 2; break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) { sum.termination = 3; break; }
    const double rho = cost_change / model_cost_change;
    if (rho > 1e-3) {
      for (int k = 0; k < NP; ++k) x[k] = cand[k];
      x_cost = eval_jac(x, false);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  return sum;
}

struct Point3 {
  double v[3] = {0, 0, 0};
};

class TriangulationEstimator {
 public:
  typedef Point3 Model;
  typedef std::vector<Point3> ModelVector;
  TriangulationEstimator(const TriObservation* obs, int n, uint32_t point_id) : obs_(obs), n_(n), id_(point_id) {}
  uint32_t pair_id() const { return id_; }
  int min_sample_size() const { return 2; }
  int non_minimal_sample_size() const { return 2; }
  int num_data() const { return n_; }
  int MinimalSolver(const std::vector<int>& sample, std::vector<Point3>* pts) const {  // :56-63
    Point3 pt;
    if (!NonMinimalSolver(sample, &pt)) return 0;
    pts->clear();
    pts->push_back(pt);
    return 1;
  }
  int NonMinimalSolver(const std::vector<int>& sample, Point3* pt) const {  // :65-86
    double S[4][4] = {};
    for (size_t k = 0; k < sample.size(); ++k) {
      const TriObservation& o = obs_[sample[k]];
      const double px = o.x[0] / o.focal, py = o.x[1] / o.focal;
      const double P0[4] = {o.R[0], o.R[1], o.R[2], o.t[0]}, P1[4] = {o.R[3], o.R[4], o.R[5], o.t[1]},
                   P2[4] = {o.R[6], o.R[7], o.R[8], o.t[2]};
      double a0[4], a1[4];
      for (int c = 0; c < 4; ++c) { a0[c] = P2[c] * px - P0[c]; a1[c] = P2[c] * py - P1[c]; }
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i][j] += a0[i] * a0[j] + a1[i] * a1[j];
    }
    double v[4];
    smallest_eigenvector4(S, v);
    for (int i = 0; i < 3; ++i) pt->v[i] = v[i] / v[3];
    return 1;
  }
  double EvaluateModelOnPoint(const Point3& pt, int i) const {  // :46-54
    ++evals_;
    const TriObservation& o = obs_[i];
    const double PX0 = (o.R[0] * pt.v[0] + o.R[1] * pt.v[1] + o.R[2] * pt.v[2]) + o.t[0];
    const double PX1 = (o.R[3] * pt.v[0] + o.R[4] * pt.v[1] + o.R[5] * pt.v[2]) + o.t[1];
    const double PX2 = (o.R[6] * pt.v[0] + o.R[7] * pt.v[1] + o.R[8] * pt.v[2]) + o.t[2];
    if (PX2 < 0) return std::numeric_limits<double>::max();
    const double r0 = o.focal * (PX0 / PX2) - o.x[0], r1 = o.focal * (PX1 / PX2) - o.x[1];
    return r0 * r0 + r1 * r1;
  }
  void LeastSquares(const std::vector<int>& sample, Point3* pt) const {  // :88-126
    lm_triangulate(obs_, sample.data(), (int)sample.size(), pt->v);
  }
  mutable long long evals_ = 0;

 private:
  const TriObservation* obs_;
  int n_;
  uint32_t id_;
};

}  // namespace ssfm_oracle
