// Stand-in (test infrastructure) for jonathanventura/polynomial's Polynomial<deg> as used at
// src/spherical_fast_estimator.cpp:10-11,220-223: constructed from deg+1 coefficients, highest degree first;
// realRootsSturm(lo, hi, roots) appends the real roots in [lo, hi].  The library itself is un-vendored (cloned at
// an unpinned HEAD by docker/Dockerfile:72-76), so its bracketing order / tolerances cannot be reproduced: this
// stand-in returns the roots in increasing order to full double precision (Sturm chain + bisection).
#pragma once
#include <vector>

#include "../../ssfm_oracle.hpp"

namespace polynomial {
template <int deg>
class Polynomial {
 public:
  explicit Polynomial(const double* coeffs) {
    for (int i = 0; i <= deg; ++i) c_[i] = coeffs[i];
  }
  void realRootsSturm(double lo, double hi, std::vector<double>& roots) const {
    static_assert(deg == 4, "only the quartic is needed");
    double r[4];
    const int n = ssfm_oracle::sturm_real_roots(c_, lo, hi, r);
    for (int i = 0; i < n; ++i) roots.push_back(r[i]);
  }

 private:
  double c_[deg + 1];
};
}  // namespace polynomial
