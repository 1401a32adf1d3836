// pinned_rand.hpp -- test infrastructure: the stream that replaces the C library's rand() when the reference's
// legacy drivers (include/sphericalsfm/msac.h:18, preemptive_ransac.h:18) are compiled into oracle/_ref: the j-th
// call made while drawing the sample of hypothesis h is ssfm_oracle::philox_rand31(seed, pair, h, j).  The headers
// are not modified; the call expression `rand()` is redirected by a macro around the #include.
#pragma once
#include <cstdint>

#include "ssfm_oracle.hpp"

namespace pinned_rand {
inline thread_local uint32_t seed = 0, pair = 0, hyp = 0, draw = 0;
inline int next() { return ssfm_oracle::philox_rand31(seed, pair, hyp, draw++); }
inline void start(uint32_t s, uint32_t p) { seed = s; pair = p; hyp = 0; draw = 0; }
inline void next_hypothesis() { ++hyp; draw = 0; }  // call right after each random_sample(): compute() follows it
}  // namespace pinned_rand
