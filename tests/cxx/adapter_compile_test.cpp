// Compiles the RansacLib-concept adapters against a reference-shaped call site (the body of
// estimate_pairwise, examples/spherical_sfm_tools.cpp:378-392) with stand-in Eigen-like types.
// Exit codes: 0 ok (GPU present), 3 engine reported "no device" (CPU-only box), 1 wrong result.
#include <array>
#include <cstdio>
#include <random>
#include <utility>
#include <vector>

#include "ssfm_ransaclib.hpp"

struct Vec3 {  // stand-in for Eigen::Vector3d (3 packed doubles)
  double d[3];
  double& operator[](int i) { return d[i]; }
  double operator[](int i) const { return d[i]; }
};
typedef std::pair<Vec3, Vec3> RayPair;
typedef std::vector<RayPair> RayPairList;

int main() {
  using namespace ssfm_b200;
  // a small synthetic pair: rotation about y, t = R e3 - e3, noise-free, 20% gross outliers
  const double a = 0.1, c = std::cos(a), s = std::sin(a);
  const double R[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
  const double t[3] = {R[2], R[5], R[8] - 1};
  std::mt19937 g(1);
  std::normal_distribution<double> nd;
  std::uniform_real_distribution<double> ud(4, 8);
  RayPairList rays(200);
  for (size_t i = 0; i < rays.size(); ++i) {
    Vec3 u{{nd(g) * 0.3, nd(g) * 0.3, 1.0}};
    const double dep = ud(g);
    double X[3] = {u[0] * dep, u[1] * dep, dep}, Y[3];
    for (int r = 0; r < 3; ++r) Y[r] = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + t[r];
    Vec3 v{{Y[0] / Y[2], Y[1] / Y[2], 1.0}};
    if (i % 5 == 0) { v[0] = nd(g); v[1] = nd(g); }
    rays[i] = std::make_pair(u, v);
  }
  try {
    Engine eng(0);
    LORansacOptions options;
    options.squared_inlier_threshold_ = 1e-6;
    options.num_lo_steps_ = 0;
    options.num_lsq_iterations_ = 0;
    options.final_least_squares_ = true;
    GpuSphericalEstimator<Mat3d> estimator(eng, rays, false, false);
    LocallyOptimizedMSAC<Mat3d, std::vector<Mat3d>, GpuSphericalEstimator<Mat3d>> ransac;
    RansacStatistics stats;
    Mat3d E;
    const int ninliers = ransac.EstimateModel(options, estimator, &E, &stats);
    int recount = 0;
    for (size_t i = 0; i < rays.size(); ++i)
      recount += estimator.EvaluateModelOnPoint(E, (int)i) < options.squared_inlier_threshold_;
    std::vector<Mat3d> Es;
    const int nm = estimator.MinimalSolver({1, 2, 3}, &Es);
    Mat3d Rm;
    Vec3 tt;
    estimator.Decompose(E, stats.inlier_indices, &Rm, &tt);
    std::vector<SsfmPairResult> res;
    EstimatePairs(eng, options, std::vector<RayPairList>{rays, rays}, false, false, &res);
    // the legacy drivers with their upstream signatures (msac.h:67, preemptive_ransac.h:46)
    typedef GpuSphericalFastEstimator<Mat3d> FastEst;
    std::vector<FastEst> pool(256);
    std::vector<FastEst*> ptrs;
    for (auto& e : pool) ptrs.push_back(&e);
    FastEst* best = nullptr;
    std::vector<bool> inl;
    MSAC<RayPairList, FastEst> msac(eng);
    msac.inlier_threshold = 1e-3;
    const int n_msac = msac.compute(rays.begin(), rays.end(), ptrs, &best, inl);
    PreemptiveRANSAC<RayPairList, FastEst> pre(eng, 10);
    pre.inlier_threshold = 1e-3;
    FastEst* best2 = nullptr;
    const int n_pre = pre.compute(rays.begin(), rays.end(), ptrs, &best2, inl);
    Vec3 rr, tr;
    if (best2) best2->decomposeE(false, rr, tr);
    // the six-point shared-focal estimator under VanillaMSAC (examples/six_point_estimator.h).  Spherical motion is a
    // degenerate configuration for two-view focal estimation (all optical axes meet at the sphere centre), so this
    // part uses a general motion; the rays are calibrated (focal 1), so the recovered focal must be ~1.
    RayPairList rays6(200);
    {
      const double tg[3] = {0.5, 0.1, 0.2}, b = 0.15, cb = std::cos(b), sb = std::sin(b);
      const double Rg[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
      for (size_t i = 0; i < rays6.size(); ++i) {
        Vec3 u{{nd(g) * 0.3, nd(g) * 0.3, 1.0}};
        const double dep = ud(g);
        double X[3] = {u[0] * dep, u[1] * dep, dep}, Y[3];
        for (int r = 0; r < 3; ++r) Y[r] = Rg[3 * r] * X[0] + Rg[3 * r + 1] * X[1] + Rg[3 * r + 2] * X[2] + tg[r];
        Vec3 v{{Y[0] / Y[2], Y[1] / Y[2], 1.0}};
        if (i % 5 == 0) { v[0] = nd(g); v[1] = nd(g); }
        rays6[i] = std::make_pair(u, v);
      }
    }
    GpuSixPointEstimator six(eng, rays6, /*focal_scoring=*/true);
    VanillaMSAC<SixPointSolution, std::vector<SixPointSolution>, GpuSixPointEstimator> ransac6;
    RansacStatistics stats6;
    SixPointSolution sol6;
    const int n_six = ransac6.EstimateModel(options, six, &sol6, &stats6);
    std::vector<SixPointSolution> sols6;
    const int nm6 = six.MinimalSolver({1, 2, 3, 4, 6, 7}, &sols6);
    SixPointSolution refit = sol6;
    refit.focal *= 1.05;  // perturb, then LeastSquares must bring it back
    six.LeastSquares(stats6.inlier_indices, &refit);
    // SfM::Retriangulate: eight cameras translated along x (identity rotations), two points, one gross outlier each
    std::vector<double> cam_tr;
    for (int c = 0; c < 8; ++c) { const double tr[6] = {-0.2 * c, 0, 0, 0, 0, 0}; cam_tr.insert(cam_tr.end(), tr, tr + 6); }
    const double Xs[2][3] = {{0.3, -0.2, 5.0}, {-0.5, 0.4, 7.0}};
    std::vector<std::vector<TrackObservation>> tracks(2);
    for (int j = 0; j < 2; ++j)
      for (int c = 0; c < 8; ++c) {
        const double px = Xs[j][0] - 0.2 * c, py = Xs[j][1], pz = Xs[j][2];
        TrackObservation ob{c, 600.0 * px / pz, 600.0 * py / pz};
        if (c == 3) ob.x += 80.0;
        tracks[j].push_back(ob);
      }
    std::vector<Point3d> pts3;
    std::vector<int32_t> tri_inl, tri_status;
    Retriangulate(eng, cam_tr, tracks, 600.0, &pts3, &tri_inl, &tri_status);
    const bool tri_ok = tri_status[0] == SSFM_PAIR_OK && tri_status[1] == SSFM_PAIR_OK && tri_inl[0] == 7 && tri_inl[1] == 7 &&
                        std::fabs(pts3[0].z - 5.0) < 1e-6 && std::fabs(pts3[1].x + 0.5) < 1e-6;
    std::printf("retriangulate: inliers %d %d point0 (%.6f %.6f %.6f)\n", tri_inl[0], tri_inl[1], pts3[0].x, pts3[0].y, pts3[0].z);
    std::printf("six-point: inliers %d focal %.6f minimal models %d refit focal %.6f\n", n_six, sol6.focal, nm6, refit.focal);
    std::printf("inliers %d recount %d models %d R02 %.6f (want %.6f) batched %d %d legacy msac %d (iter %d) preemptive %d ry %.6f\n",
                ninliers, recount, nm, Rm(0, 2), s, res[0].best_num_inliers, res[1].best_num_inliers, n_msac, msac.iter, n_pre,
                best2 ? rr[1] : 0.0);
    const bool ok = ninliers == 160 && recount == ninliers && nm == 4 && std::fabs(Rm(0, 2) - s) < 1e-6 &&
                    res[0].best_num_inliers == 160 && n_msac == 160 && best != nullptr && n_pre == 160 && best2 != nullptr &&
                    std::fabs(rr[1] - a) < 1e-6 && (int)inl.size() == 200 && n_six == 160 &&
                    std::fabs(sol6.focal - 1.0) < 1e-6 && nm6 >= 1 && std::fabs(refit.focal - 1.0) < 5e-3 && tri_ok;  // the refit stops at Ceres' function tolerance
    return ok ? 0 : 1;
  } catch (const Error& e) {
    std::printf("engine error %d: %s\n", e.code(), e.what());
    return e.code() == SSFM_ERR_NO_DEVICE ? 3 : 2;
  }
}
