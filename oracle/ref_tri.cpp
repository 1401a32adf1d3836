// ref_tri.cpp -- oracle/_ref/libssfm_reftri.so: the reference's OWN triangulation path, compiled where it lies:
//   /root/reference/src/triangulation_estimator.cpp   (TriangulationEstimator: DLT, reprojection error, Ceres refit)
//   /root/reference/src/sfm_types.cpp, src/so3.cpp    (Pose, so3exp)
//   /root/reference/include/RansacLib/ransac.h        (LocallyOptimizedMSAC)
// against the Eigen / Ceres stand-ins of this directory (oracle/eigen_shim, oracle/ceres_shim), with the Philox
// sampler as RansacLib's Sampler template parameter.  Test infrastructure only: it pins oracle/tri_oracle.hpp.
#include <cstring>
#include <vector>

#include <Eigen/Core>
#include <RansacLib/ransac.h>
#include <sphericalsfm/sfm_types.h>
#include <sphericalsfm/triangulation_estimator.h>

#include "oracle_capi.h"
#include "ssfm_oracle.hpp"

namespace {
struct CountingTri : public sphericalsfm::TriangulationEstimator {
  uint32_t id;
  CountingTri(const sphericalsfm::TriangulationObservationList& o, uint32_t id_) : sphericalsfm::TriangulationEstimator(o), id(id_) {}
  uint32_t pair_id() const { return id; }
};
}  // namespace

extern "C" {

int orc_is_reference(void) { return 3; }

int orc_triangulate(const double* cam_tr, const double* obs_xy, int n, double focal, const OrcOptions* o, uint32_t point_id,
                    OrcResult* out, int* inlier_idx) {
  using namespace sphericalsfm;
  std::memset(out, 0, sizeof(*out));
  out->status = 3;
  out->best_model_score = std::numeric_limits<double>::max();
  if (n < 3) return 0;  // src/sfm.cpp:173
  TriangulationObservationList obs;
  for (int i = 0; i < n; ++i) {
    Eigen::Vector3d t, r;
    for (int k = 0; k < 3; ++k) { t(k) = cam_tr[6 * i + k]; r(k) = cam_tr[6 * i + 3 + k]; }
    Eigen::Vector2d x;
    x(0) = obs_xy[2 * i]; x(1) = obs_xy[2 * i + 1];
    obs.push_back(TriangulationObservation(Pose(t, r), x, focal));
  }
  ransac_lib::LORansacOptions ro;
  ro.min_num_iterations_ = o->min_num_iterations; ro.max_num_iterations_ = o->max_num_iterations;
  ro.success_probability_ = o->success_probability; ro.squared_inlier_threshold_ = o->squared_inlier_threshold;
  ro.random_seed_ = o->random_seed; ro.num_lo_steps_ = o->num_lo_steps; ro.threshold_multiplier_ = o->threshold_multiplier;
  ro.num_lsq_iterations_ = o->num_lsq_iterations; ro.min_sample_multiplicator_ = o->min_sample_multiplicator;
  ro.non_min_sample_multiplier_ = o->non_min_sample_multiplier; ro.lo_starting_iterations_ = o->lo_starting_iterations;
  ro.final_least_squares_ = o->final_least_squares != 0;
  ransac_lib::RansacStatistics rs;
  CountingTri est(obs, point_id);
  ransac_lib::LocallyOptimizedMSAC<Point, std::vector<Point>, CountingTri, ssfm_oracle::PhiloxSampling<CountingTri> > ransac;
  Point X;
  X(0) = X(1) = X(2) = 0.0;
  const int ninl = ransac.EstimateModel(ro, est, &X, &rs);
  out->num_iterations = rs.num_iterations;
  out->best_num_inliers = rs.best_num_inliers;
  out->best_model_score = rs.best_model_score;
  out->inlier_ratio = rs.inlier_ratio;
  out->number_lo_iterations = rs.number_lo_iterations;
  if (ninl < 3) {  // src/sfm.cpp:186
    out->status = 2;
  } else {
    out->status = 0;
    for (int k = 0; k < 3; ++k) out->E[k] = X(k);
  }
  if (inlier_idx)
    for (size_t i = 0; i < rs.inlier_indices.size(); ++i) inlier_idx[i] = rs.inlier_indices[i];
  return ninl;
}
}
