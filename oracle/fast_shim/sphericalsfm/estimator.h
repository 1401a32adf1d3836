// Stand-in (test infrastructure) for the OLD, non-template sphericalsfm::Estimator interface that the orphan
// sources include/sphericalsfm/spherical_fast_estimator.h, msac.h and preemptive_ransac.h were written against
// (sampleSize / compute / chooseSolution / score / canRefine -- see their uses at msac.h:71,85,91,
// preemptive_ransac.h:50,69,84).  The header of that name in today's reference tree declares the templated
// RansacLib-style concept instead, which is why upstream no longer builds spherical_fast_estimator.cpp.
// This directory is put in front of the reference's include path only for oracle/_ref/libssfm_reffast.so.
#pragma once
#include <sphericalsfm/ray.h>

namespace sphericalsfm {
struct Estimator {
  virtual ~Estimator() {}
  virtual int sampleSize() = 0;
  virtual double score(RayPairList::iterator it) = 0;
  virtual void chooseSolution(int soln) = 0;
  virtual int compute(RayPairList::iterator begin, RayPairList::iterator end) = 0;
  virtual bool canRefine() = 0;
};
}  // namespace sphericalsfm
