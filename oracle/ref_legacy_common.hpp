// shared by ref_legacy_msac.cpp / ref_legacy_preemptive.cpp (test infrastructure)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <vector>

#include <Eigen/Core>
#include <sphericalsfm/spherical_fast_estimator.h>

#include "oracle_capi.h"
#include "pinned_rand.hpp"

// The reference's SphericalFastEstimator with two hooks: compute() marks the end of a random_sample() call for the
// pinned rand() stream, score() counts evaluations.
struct HookedFastEstimator : public sphericalsfm::SphericalFastEstimator {
  long long* evals = nullptr;
  int compute(sphericalsfm::RayPairList::iterator b, sphericalsfm::RayPairList::iterator e) {
    pinned_rand::next_hypothesis();
    const int n = sphericalsfm::SphericalFastEstimator::compute(b, e);
    // Upstream leaves E unset until chooseSolution() is called, and PreemptiveRANSAC never calls it when there is
    // exactly one solution (preemptive_ransac.h:75).  Define that case as the restatement does: the single
    // solution is the hypothesis; no solution -> a matrix that is never an inlier.
    if (n >= 1) E = Esolns[0];
    else E = Eigen::Matrix3d::Zero();
    return n;
  }
  double score(sphericalsfm::RayPairList::iterator it) {
    if (evals) ++*evals;
    return sphericalsfm::SphericalFastEstimator::score(it);
  }
};

inline void fill_list(const double* rays, int n, sphericalsfm::RayPairList* list) {
  list->resize(n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { (*list)[i].first(k) = rays[6 * i + k]; (*list)[i].second(k) = rays[6 * i + 3 + k]; }
}

// Fill an OrcResult from what the legacy drivers return (best estimator, inlier mask, count).
inline int finish_legacy(HookedFastEstimator* best, const std::vector<bool>& inl, int ninl, sphericalsfm::RayPairList& list,
                         double thr2, int inward, uint32_t iterations, long long evals, OrcResult* out, int* inlier_idx) {
  std::memset(out, 0, sizeof(*out));
  out->num_iterations = iterations;
  out->evals = evals;
  out->best_model_score = std::numeric_limits<double>::max();
  out->status = 2;
  if (!best) return 0;
  const int n = (int)list.size();
  double cost = 0.0;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const double sc = best->sphericalsfm::SphericalFastEstimator::score(list.begin() + i);
    cost += (sc <= thr2) ? sc : thr2;
    if (inl[i] && inlier_idx) inlier_idx[k++] = i;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out->E[3 * r + c] = best->E(r, c);
  out->best_num_inliers = ninl;
  out->best_model_score = cost;
  out->inlier_ratio = (double)ninl / (double)n;
  out->status = 0;
  Eigen::Vector3d r, t;
  best->decomposeE(inward != 0, r, t);
  for (int d = 0; d < 3; ++d) { out->r[d] = r(d); out->t[d] = t(d); }
  return ninl;
}
