// oracle/_ref/libssfm_reflegacy.so, part 2: the reference's PreemptiveRANSAC
// (/root/reference/include/sphericalsfm/preemptive_ransac.h) around its SphericalFastEstimator.  (A separate
// translation unit: msac.h and preemptive_ransac.h both define sphericalsfm::random_sample.)
#include "ref_legacy_common.hpp"

#define rand() pinned_rand::next()
#include <sphericalsfm/preemptive_ransac.h>
#undef rand

extern "C" int orc_legacy_preemptive(const double* rays, int n, const OrcOptions* o, uint32_t pair_id, OrcResult* out,
                                     int* inlier_idx) {
  using namespace sphericalsfm;
  std::memset(out, 0, sizeof(*out));
  out->status = 1;
  out->best_model_score = std::numeric_limits<double>::max();
  if (n < 4 || o->legacy_budget <= 0 || o->preemptive_block <= 0) return 0;
  RayPairList list;
  fill_list(rays, n, &list);
  long long evals = 0;
  std::vector<HookedFastEstimator> pool(o->legacy_budget);
  std::vector<HookedFastEstimator*> ptrs(o->legacy_budget);
  for (int i = 0; i < o->legacy_budget; ++i) {
    pool[i].evals = &evals;
    pool[i].E = Eigen::Matrix3d::Zero();
    ptrs[i] = &pool[i];
  }
  pinned_rand::start(o->random_seed, pair_id);
  PreemptiveRANSAC<RayPairList, HookedFastEstimator> pr((size_t)o->preemptive_block);
  pr.inlier_threshold = std::sqrt(o->squared_inlier_threshold);
  HookedFastEstimator* best = nullptr;
  std::vector<bool> inl;
  const int ninl = pr.compute(list.begin(), list.end(), ptrs, &best, inl);
  return finish_legacy(best, inl, ninl, list, o->squared_inlier_threshold, o->inward, (uint32_t)o->legacy_budget, evals, out,
                       inlier_idx);
}
