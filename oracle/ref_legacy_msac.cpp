// oracle/_ref/libssfm_reflegacy.so, part 1: the reference's legacy fixed-budget MSAC driver
// (/root/reference/include/sphericalsfm/msac.h) around its SphericalFastEstimator
// (/root/reference/src/spherical_fast_estimator.cpp), compiled where they lie.  This is config C2's path as upstream
// wrote it.  Stand-ins: oracle/eigen_shim, oracle/fast_shim; rand() -> pinned_rand (see pinned_rand.hpp).
#include "ref_legacy_common.hpp"

#define rand() pinned_rand::next()
#include <sphericalsfm/msac.h>
#undef rand

extern "C" int orc_legacy_msac(const double* rays, int n, const OrcOptions* o, uint32_t pair_id, OrcResult* out, int* inlier_idx) {
  using namespace sphericalsfm;
  std::memset(out, 0, sizeof(*out));
  out->status = 1;
  out->best_model_score = std::numeric_limits<double>::max();
  if (n < 3 || o->legacy_budget <= 0) return 0;
  RayPairList list;
  fill_list(rays, n, &list);
  long long evals = 0;
  std::vector<HookedFastEstimator> pool(o->legacy_budget);
  std::vector<HookedFastEstimator*> ptrs(o->legacy_budget);
  for (int i = 0; i < o->legacy_budget; ++i) { pool[i].evals = &evals; ptrs[i] = &pool[i]; }
  pinned_rand::start(o->random_seed, pair_id);
  MSAC<RayPairList, HookedFastEstimator> msac(o->legacy_prob_success);
  msac.inlier_threshold = std::sqrt(o->squared_inlier_threshold);
  HookedFastEstimator* best = nullptr;
  std::vector<bool> inl(n, false);
  const int ninl = msac.compute(list.begin(), list.end(), ptrs, &best, inl);
  return finish_legacy(best, inl, ninl, list, o->squared_inlier_threshold, o->inward, (uint32_t)msac.iter, evals, out, inlier_idx);
}
