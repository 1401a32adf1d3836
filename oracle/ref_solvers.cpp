// ref_solvers.cpp -- extern "C" surface over the reference's OWN solver / geometry sources, compiled
// unmodified from /root/reference (src/spherical_solvers.cpp, src/so3.cpp, src/spherical_utils.cpp)
// against oracle/eigen_shim (a stand-in for the Eigen subset they use; Eigen is not installed).
// Built by `make ref` into oracle/_ref/libssfm_refsolvers.so.  TEST INFRASTRUCTURE ONLY: it pins the
// oracle's restated solvers (tests/test_oracle.py::test_restated_solvers_match_reference_sources).
#include <sphericalsfm/so3.h>
#include <sphericalsfm/spherical_solvers.h>
#include <sphericalsfm/spherical_utils.h>

#include <cstring>

using namespace sphericalsfm;

extern "C" {

// rays: n x 6 doubles (RayPair memory); models: up to 4 row-major 3x3.  Returns the number of models.
int refsrc_solve(const double* rays, int n, const int* sample, int ns, int use_poly, double* models9) {
  RayPairList corr(n);
  for (int i = 0; i < n; ++i) {
    corr[i].first = Eigen::Vector3d(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    corr[i].second = Eigen::Vector3d(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
  }
  std::vector<int> s(sample, sample + ns);
  std::vector<Eigen::Matrix3d> Es;
  const int nm = use_poly ? spherical_solver_polynomial(corr, s, &Es) : spherical_solver_action_matrix(corr, s, &Es);
  for (int k = 0; k < nm && k < 4; ++k)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) models9[9 * k + 3 * i + j] = Es[k](i, j);
  return nm;
}

void refsrc_decompose(const double* E9, int inward, double* r, double* t) {
  Eigen::Matrix3d E;
  for (int i = 0; i < 9; ++i) E.d[i] = E9[i];
  Eigen::Vector3d rr, tt;
  decompose_spherical_essential_matrix(E, inward != 0, rr, tt);
  for (int i = 0; i < 3; ++i) { r[i] = rr(i); t[i] = tt(i); }
}

void refsrc_make_E(const double* r, int inward, double* E9) {
  Eigen::Matrix3d E;
  make_spherical_essential_matrix(so3exp(Eigen::Vector3d(r[0], r[1], r[2])), inward != 0, E);
  for (int i = 0; i < 9; ++i) E9[i] = E.d[i];
}

void refsrc_so3(const double* r, double* R9, double* r_back) {
  const Eigen::Matrix3d R = so3exp(Eigen::Vector3d(r[0], r[1], r[2]));
  for (int i = 0; i < 9; ++i) R9[i] = R.d[i];
  const Eigen::Vector3d rb = so3ln(R);
  for (int i = 0; i < 3; ++i) r_back[i] = rb(i);
}

}  // extern "C"
