// ref_fast.cpp -- oracle/_ref/libssfm_reffast.so: the reference's orphan SphericalFastEstimator
// (/root/reference/src/spherical_fast_estimator.cpp, include/sphericalsfm/spherical_fast_estimator.h), compiled
// where it lies and UNMODIFIED, against oracle/eigen_shim plus two stand-ins in oracle/fast_shim (the old
// non-template Estimator base it was written for, and Polynomial<4>::realRootsSturm).  Test infrastructure only:
// pins the restated FAST_STURM solver (constraint rows, monomial order, elimination, det N(y) quartic, x by Cramer,
// E assembly and normalisation) and the estimator's score / decomposeE against the reference's own source.
#include <cstring>
#include <vector>

#include <Eigen/Core>
#include <sphericalsfm/spherical_fast_estimator.h>

extern "C" {

// rays: three correspondences (18 doubles).  E: up to 4 x 9 row-major.  Returns the number of solutions.
int orc_fast_compute(const double* rays18, double* E36) {
  using namespace sphericalsfm;
  RayPairList list(3);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) { list[i].first(k) = rays18[6 * i + k]; list[i].second(k) = rays18[6 * i + 3 + k]; }
  SphericalFastEstimator est;
  const int n = est.compute(list.begin(), list.end());
  for (int s = 0; s < n && s < 4; ++s)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) E36[9 * s + 3 * r + c] = est.Esolns[s](r, c);
  return n;
}

// score (:23-32) of E on n correspondences, and decomposeE (:290-341)
void orc_fast_score(const double* E9, const double* rays, int n, double* out) {
  using namespace sphericalsfm;
  SphericalFastEstimator est;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) est.E(r, c) = E9[3 * r + c];
  RayPairList list(n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { list[i].first(k) = rays[6 * i + k]; list[i].second(k) = rays[6 * i + 3 + k]; }
  for (int i = 0; i < n; ++i) out[i] = est.score(list.begin() + i);
}

void orc_fast_decompose(const double* E9, int inward, double* r3, double* t3) {
  using namespace sphericalsfm;
  SphericalFastEstimator est;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) est.E(r, c) = E9[3 * r + c];
  Eigen::Vector3d r, t;
  est.decomposeE(inward != 0, r, t);
  for (int k = 0; k < 3; ++k) { r3[k] = r(k); t3[k] = t(k); }
}
}
