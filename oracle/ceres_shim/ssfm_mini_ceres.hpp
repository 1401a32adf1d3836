// ssfm_mini_ceres.hpp -- a tiny stand-in for the subset of Ceres 2.2 that the reference's
// src/spherical_estimator.cpp uses (AutoDiffCostFunction, Problem, Solver with TRUST_REGION /
// LEVENBERG_MARQUARDT / DENSE_NORMAL_CHOLESKY, AngleAxisToRotationMatrix), so that file compiles
// UNMODIFIED into oracle/_ref.  TEST INFRASTRUCTURE ONLY; this is not Ceres and shares no code with it.
// The minimiser restates Ceres' documented algorithm and Solver::Options defaults
// (function_tolerance 1e-6, gradient_tolerance 1e-10, parameter_tolerance 1e-8, initial trust region
// radius 1e4, min_relative_decrease 1e-3, min/max_lm_diagonal 1e-6/1e32, jacobi_scaling, monotonic
// steps): "parity unpinned" for the LM trajectory itself, see DESIGN.md.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <map>
#include <memory>
#include <utility>
#include <vector>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(const T& x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(int x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }        // NOLINT
  Jet& operator+=(const Jet& y) { a += y.a; for (int i = 0; i < N; ++i) v[i] += y.v[i]; return *this; }
  Jet& operator-=(const Jet& y) { a -= y.a; for (int i = 0; i < N; ++i) v[i] -= y.v[i]; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};
#define SSFM_JET template <typename T, int N> inline Jet<T, N>
SSFM_JET operator+(const Jet<T, N>& x, const Jet<T, N>& y) { Jet<T, N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
SSFM_JET operator-(const Jet<T, N>& x, const Jet<T, N>& y) { Jet<T, N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
SSFM_JET operator-(const Jet<T, N>& x) { Jet<T, N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
SSFM_JET operator*(const Jet<T, N>& x, const Jet<T, N>& y) { Jet<T, N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
SSFM_JET operator/(const Jet<T, N>& x, const Jet<T, N>& y) { Jet<T, N> r; const T inv = T(1) / y.a; r.a = x.a * inv; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
SSFM_JET operator+(const Jet<T, N>& x, T s) { Jet<T, N> r(x); r.a += s; return r; }
SSFM_JET operator+(T s, const Jet<T, N>& x) { Jet<T, N> r(x); r.a += s; return r; }
SSFM_JET operator-(const Jet<T, N>& x, T s) { Jet<T, N> r(x); r.a -= s; return r; }
SSFM_JET operator-(T s, const Jet<T, N>& x) { Jet<T, N> r(-x); r.a += s; return r; }
SSFM_JET operator*(const Jet<T, N>& x, T s) { Jet<T, N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
SSFM_JET operator*(T s, const Jet<T, N>& x) { return x * s; }
SSFM_JET operator/(const Jet<T, N>& x, T s) { return x * (T(1) / s); }
SSFM_JET sqrt(const Jet<T, N>& x) { Jet<T, N> r; r.a = std::sqrt(x.a); const T f = T(0.5) / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * f; return r; }
SSFM_JET sin(const Jet<T, N>& x) { Jet<T, N> r; r.a = std::sin(x.a); const T c = std::cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = c * x.v[i]; return r; }
SSFM_JET cos(const Jet<T, N>& x) { Jet<T, N> r; r.a = std::cos(x.a); const T s = -std::sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = s * x.v[i]; return r; }
#undef SSFM_JET
template <typename T, int N> inline bool operator>(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a > y.a; }
template <typename T, int N> inline bool operator<(const Jet<T, N>& x, const Jet<T, N>& y) { return x.a < y.a; }

template <typename T>
inline T DotProduct(const T x[3], const T y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

// ceres/rotation.h
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  using std::cos; using std::sin; using std::sqrt;
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta), sintheta = sin(theta), theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// ceres/rotation.h: column-major 3x3 output.
template <typename T>
inline void AngleAxisToRotationMatrix(const T* angle_axis, T* R) {
  using std::cos; using std::sin; using std::sqrt;
  static const T kOne = T(1.0);
  const T theta2 = DotProduct(angle_axis, angle_axis);
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T wx = angle_axis[0] / theta, wy = angle_axis[1] / theta, wz = angle_axis[2] / theta;
    const T costheta = cos(theta), sintheta = sin(theta);
    R[0] = costheta + wx * wx * (kOne - costheta);
    R[1] = wz * sintheta + wx * wy * (kOne - costheta);
    R[2] = -wy * sintheta + wx * wz * (kOne - costheta);
    R[3] = wx * wy * (kOne - costheta) - wz * sintheta;
    R[4] = costheta + wy * wy * (kOne - costheta);
    R[5] = wx * sintheta + wy * wz * (kOne - costheta);
    R[6] = wy * sintheta + wx * wz * (kOne - costheta);
    R[7] = -wx * sintheta + wy * wz * (kOne - costheta);
    R[8] = costheta + wz * wz * (kOne - costheta);
  } else {
    R[0] = kOne; R[1] = angle_axis[2]; R[2] = -angle_axis[1];
    R[3] = -angle_axis[2]; R[4] = kOne; R[5] = angle_axis[0];
    R[6] = angle_axis[1]; R[7] = -angle_axis[0]; R[8] = kOne;
  }
}

class LossFunction {};

class CostFunction {
 public:
  virtual ~CostFunction() {}
  // jacobians[b] (row-major num_residuals x block_size) may be null.
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  std::vector<int> block_sizes;
  int num_residuals = 0;
};

template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
 public:
  static constexpr int kTotal = (Ns + ...);
  static constexpr int kBlocks = sizeof...(Ns);
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {
    block_sizes = {Ns...};
    num_residuals = kNumResiduals;
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    typedef Jet<double, kTotal> J;
    std::vector<J> x(kTotal);
    const J* ptr[kBlocks];
    int off = 0;
    for (int b = 0; b < kBlocks; ++b) {
      ptr[b] = &x[off];
      for (int i = 0; i < block_sizes[b]; ++i) {
        x[off + i].a = parameters[b][i];
        x[off + i].v[off + i] = 1.0;
      }
      off += block_sizes[b];
    }
    J res[kNumResiduals];
    if (!call(ptr, res, std::make_index_sequence<kBlocks>())) return false;
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = res[r].a;
    if (jacobians) {
      off = 0;
      for (int b = 0; b < kBlocks; ++b) {
        if (jacobians[b])
          for (int r = 0; r < kNumResiduals; ++r)
            for (int i = 0; i < block_sizes[b]; ++i) jacobians[b][r * block_sizes[b] + i] = res[r].v[off + i];
        off += block_sizes[b];
      }
    }
    return true;
  }

 private:
  template <class J, size_t... I>
  bool call(const J* const* p, J* res, std::index_sequence<I...>) const { return (*f_)(p[I]..., res); }
  std::unique_ptr<Functor> f_;
};

enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

class Problem {
 public:
  struct Block {
    std::unique_ptr<CostFunction> cost;
    std::vector<double*> params;
  };
  template <typename... Ps>
  void AddResidualBlock(CostFunction* cost, LossFunction*, Ps... ps) {
    Block b;
    b.cost.reset(cost);
    b.params = {ps...};
    for (size_t i = 0; i < b.params.size(); ++i)
      if (!sizes.count(b.params[i])) { sizes[b.params[i]] = cost->block_sizes[i]; order.push_back(b.params[i]); }
    blocks.push_back(std::move(b));
  }
  void SetParameterBlockConstant(double* p) { constant[p] = true; }
  std::vector<Block> blocks;
  std::map<double*, int> sizes;
  std::map<double*, bool> constant;
  std::vector<double*> order;
};

class Solver {
 public:
  struct Options {
    MinimizerType minimizer_type = TRUST_REGION;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    int max_num_consecutive_invalid_steps = 5;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  };
  struct Summary {
    int num_iterations = 0;
    double initial_cost = 0, final_cost = 0;
    int termination = 0;
  };
};

namespace detail {
inline bool cholesky_solve(std::vector<double> A, const std::vector<double>& b, int n, std::vector<double>* x) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    A[j * n + j] = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / A[j * n + j];
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * y[k];
    y[i] = s / A[i * n + i];
  }
  x->assign(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * (*x)[k];
    (*x)[i] = s / A[i * n + i];
  }
  for (double v : *x)
    if (!std::isfinite(v)) return false;
  return true;
}
}  // namespace detail

// TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseNormalCholeskySolver, restated.
inline void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
  Problem& P = *problem;
  // free parameter blocks in order of first appearance
  std::vector<double*> free_blocks;
  std::map<double*, int> offset;
  int n = 0;
  for (double* p : P.order)
    if (!P.constant.count(p)) { offset[p] = n; n += P.sizes[p]; free_blocks.push_back(p); }
  int m = 0;
  for (auto& b : P.blocks) m += b.cost->num_residuals;
  Solver::Summary sum;
  if (n == 0 || m == 0) { if (summary) *summary = sum; return; }
  std::vector<double> x(n), res(m), jac((size_t)m * n), cand_res(m), scale(n), diagonal(n), gradient(n);
  auto gather = [&](std::vector<double>& xx) { for (double* p : free_blocks) for (int i = 0; i < P.sizes[p]; ++i) xx[offset[p] + i] = p[i]; };
  auto scatter = [&](const std::vector<double>& xx) { for (double* p : free_blocks) for (int i = 0; i < P.sizes[p]; ++i) p[i] = xx[offset[p] + i]; };
  auto evaluate = [&](std::vector<double>& r, std::vector<double>* J) -> double {
    if (J) std::fill(J->begin(), J->end(), 0.0);
    int row = 0;
    double cost = 0.0;
    for (auto& b : P.blocks) {
      const int nb = (int)b.params.size(), nr = b.cost->num_residuals;
      std::vector<std::vector<double>> jb(nb);
      std::vector<double*> jp(nb, nullptr);
      if (J)
        for (int k = 0; k < nb; ++k)
          if (!P.constant.count(b.params[k])) { jb[k].assign((size_t)nr * b.cost->block_sizes[k], 0.0); jp[k] = jb[k].data(); }
      b.cost->Evaluate(b.params.data(), &r[row], J ? jp.data() : nullptr);
      for (int q = 0; q < nr; ++q) cost += r[row + q] * r[row + q];
      if (J)
        for (int k = 0; k < nb; ++k)
          if (jp[k])
            for (int q = 0; q < nr; ++q)
              for (int i = 0; i < b.cost->block_sizes[k]; ++i)
                (*J)[(size_t)(row + q) * n + offset[b.params[k]] + i] += jb[k][q * b.cost->block_sizes[k] + i];
      row += nr;
    }
    return 0.5 * cost;
  };
  gather(x);
  double gmax = 0.0;
  bool have_scale = false;
  auto eval_jac = [&]() -> double {
    const double c = evaluate(res, &jac);
    for (int k = 0; k < n; ++k) {
      double g = 0.0;
      for (int i = 0; i < m; ++i) g += jac[(size_t)i * n + k] * res[i];
      gradient[k] = g;
    }
    if (!have_scale) {
      for (int k = 0; k < n; ++k) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += jac[(size_t)i * n + k] * jac[(size_t)i * n + k];
        scale[k] = 1.0 / (1.0 + std::sqrt(s));
      }
      have_scale = true;
    }
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < n; ++k) jac[(size_t)i * n + k] *= scale[k];
    gmax = 0.0;
    for (int k = 0; k < n; ++k) gmax = std::max(gmax, std::fabs(gradient[k]));
    return c;
  };
  double x_cost = eval_jac();
  sum.initial_cost = sum.final_cost = x_cost;
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid = 0, iteration = 0;
  if (std::isfinite(x_cost)) {
    for (;;) {
      if (iteration >= opt.max_num_iterations) { sum.termination = 4; break; }
      if (gmax <= opt.gradient_tolerance) { sum.termination = 1; break; }
      if (radius < opt.min_trust_region_radius) { sum.termination = 6; break; }
      ++iteration;
      if (!reuse_diagonal)
        for (int k = 0; k < n; ++k) {
          double s = 0.0;
          for (int i = 0; i < m; ++i) s += jac[(size_t)i * n + k] * jac[(size_t)i * n + k];
          diagonal[k] = std::min(std::max(s, opt.min_lm_diagonal), opt.max_lm_diagonal);
        }
      std::vector<double> H((size_t)n * n, 0.0), g(n, 0.0), step;
      for (int i = 0; i < m; ++i)
        for (int a = 0; a < n; ++a) {
          g[a] += jac[(size_t)i * n + a] * res[i];
          for (int b = 0; b <= a; ++b) H[(size_t)a * n + b] += jac[(size_t)i * n + a] * jac[(size_t)i * n + b];
        }
      for (int a = 0; a < n; ++a) {
        for (int b = 0; b < a; ++b) H[(size_t)b * n + a] = H[(size_t)a * n + b];
        H[(size_t)a * n + a] += diagonal[a] / radius;
      }
      bool valid = detail::cholesky_solve(H, g, n, &step);
      reuse_diagonal = true;
      double model_cost_change = 0.0;
      if (valid) {
        for (int k = 0; k < n; ++k) step[k] = -step[k];
        for (int i = 0; i < m; ++i) {
          double mm = 0.0;
          for (int k = 0; k < n; ++k) mm += jac[(size_t)i * n + k] * step[k];
          model_cost_change -= mm * (res[i] + mm / 2.0);
        }
        if (!(model_cost_change > 0.0)) valid = false;
      }
      if (!valid) {
        if (++invalid >= opt.max_num_consecutive_invalid_steps) { sum.termination = 5; break; }
        radius /= decrease_factor;
        decrease_factor *= 2.0;
        continue;
      }
      invalid = 0;
      std::vector<double> cand(n);
      double step_norm = 0.0, x_norm = 0.0;
      for (int k = 0; k < n; ++k) {
        const double dlt = step[k] * scale[k];
        cand[k] = x[k] + dlt;
        step_norm += dlt * dlt;
        x_norm += x[k] * x[k];
      }
      step_norm = std::sqrt(step_norm);
      x_norm = std::sqrt(x_norm);
      scatter(cand);
      double cand_cost = evaluate(cand_res, nullptr);
      if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
      if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { scatter(x); sum.termination = 2; break; }
      const double cost_change = x_cost - cand_cost;
      if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { scatter(x); sum.termination = 3; break; }
      const double rho = cost_change / model_cost_change;
      if (rho > opt.min_relative_decrease) {
        x = cand;
        x_cost = eval_jac();
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0;
        reuse_diagonal = false;
      } else {
        scatter(x);
        radius /= decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = true;
      }
    }
  }
  scatter(x);
  sum.num_iterations = iteration;
  sum.final_cost = x_cost;
  if (summary) *summary = sum;
}

}  // namespace ceres
