// ssfm_match.cu -- descriptor matching for many image pairs: the step in front of the
// relative-pose path, and the one dense contraction on it.
//
// Reference: match() (examples/spherical_sfm_tools.cpp:235-251) = cv::BFMatcher (NORM_L2, no cross check) ::knnMatch(
// query = features1.descs, train = features0.descs, k = 2), Lowe's ratio test `d0 < ratio * d1`, and
// `m01[trainIdx] = queryIdx` in query order (a later query overwrites an earlier one on the same train index);
// match_exhaustive() (:575-600) runs it for every image pair under `#pragma omp parallel for`.
//
// B200 mapping.  For one pair the 2-NN search is a running top-2 per query row of d^2 = |q|^2 + |t|^2 - 2 q.t over all train
// rows.  SIFT descriptors are integers 0..255 stored as float (cv::SIFT), so in fp16 they are exact, every product is exact
// in fp32 and every partial sum is an integer below 2^24: the tensor-core result is the exact integer OpenCV's float
// accumulation also produces, and dist = sqrtf(d^2) is the same float.  Hence bit-exact index pairs, ties included (OpenCV
// keeps the lower train index on equal float distances: strict '<' in its insertion).
//
// The whole of |t|^2 - 2 q.t comes out of the tensor core: the contraction is K = 144 wide, 128 descriptor bins plus one
// 16-wide block that carries |t|^2 split into three fp16-exact pieces (train side: -2 t, p0, p1, p2; query side: q, 1, 64,
// 4096 with |t|^2 = p0 + 64 p1 + 4096 p2), so the epilogue is one compare per distance.
//   k_desc_pack   float descriptors -> fp16 in the tensor-core "core matrix" order (8 rows x 16 bytes contiguous, 18 such
//                 blocks per 8-row group), once in query form and once in train form, |x|^2 per row, rows of every image
//                 padded to a multiple of 256; a tile of a packed buffer is ONE contiguous 1-D TMA bulk copy, no swizzle
//   k_match_2nn   persistent, one CTA per SM walking (pair, 256 query rows) work items; warp 0 = TMA producer (+ TMEM alloc),
//                 warp 1 = tcgen05.mma issuer (two M=128 halves x N=128 x K=16, fp16 -> fp32 in TMEM, accumulators double
//                 buffered = all 512 TMEM columns), warps 2-9 = epilogue (tcgen05.ld, one query row per thread, running top-2
//                 with OpenCV's insertion rule, ratio test, atomicMax of the query index on the winning train index = the
//                 reference's overwrite order).  256 query rows per CTA keep the train-tile stream at 32 B/clk/SM, under the
//                 L2 limit (128 rows would need 64 B/clk/SM against ~43 available).
//   k_match_count / k_match_write   ordered compaction of the owner table into the Matches list (sorted by index in image 0,
//                 the iteration order of the reference's std::map)
// There is no CPU fallback; the entry fails with SSFM_ERR_NO_DEVICE / SSFM_ERR_CUDA like the rest of the library.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/ssfm.h"

extern "C" void ssfm_internal_set_error(const char* msg);  // ssfm_engine.cu
extern "C" int ssfm_internal_device(ssfm_handle h);
extern "C" int ssfm_internal_num_sms(ssfm_handle h);
extern "C" cudaStream_t ssfm_internal_stream(ssfm_handle h);
extern "C" void** ssfm_internal_ctx_slot(ssfm_handle h, void (*deleter)(void*));  // per-engine context owned by this TU

namespace {

constexpr int kD = 128;            // descriptor length (SIFT)
constexpr int kKp = 144;           // contraction length: 128 bins + one 16-wide block carrying |t|^2
constexpr int kQTile = 256;        // query rows per work item = two UMMA M=128 halves
constexpr int kTTile = 128;        // train rows per MMA tile = UMMA N
constexpr int kRowPad = 256;       // rows of every image are padded to a multiple of this in the packed buffers
constexpr int kGroupBytes = (kKp / 8) * 128;  // one 8-row group: 18 core matrices (8 rows x 16 B) = 2304 bytes
constexpr int kStages = 3;         // train-tile ring in shared memory
constexpr int kAccBufs = 2;        // accumulator buffers in TMEM (each: 2 halves x 128 fp32 columns)
#ifndef SSFM_MATCH_COLSPLIT
#define SSFM_MATCH_COLSPLIT 2
#endif
constexpr int kColSplit = SSFM_MATCH_COLSPLIT;  // epilogue threads per query row: each scans kTTile / kColSplit columns of every tile
constexpr int kEpiWarps = 8 * kColSplit;        // (two epilogue warps per scheduler cannot hide the TMEM-load and ALU latencies)
constexpr int kMatchThreads = 32 * (2 + kEpiWarps);  // warp 0 producer, warp 1 MMA, the rest epilogue
constexpr int kEpiChunks = kTTile / 32 / kColSplit;  // 32-column TMEM loads per tile and thread
static_assert(kColSplit == 1 || kColSplit == 2, "column split");
#ifndef SSFM_MATCH_DOUBLEBUF
#define SSFM_MATCH_DOUBLEBUF 1
#endif
constexpr uint32_t kQBytes = kQTile / 8 * kGroupBytes;  // 73 728
constexpr uint32_t kTBytes = kTTile / 8 * kGroupBytes;  // 36 864
constexpr uint32_t kTmemCols = 512;
constexpr float kFarSq = 1.0e8f;   // "no neighbour yet": above any real d^2 (<= 2 * 128 * 255^2), below a padded train row's
constexpr int kPadP2 = 60000;      // p2 of a padded train row: its d^2 - |q|^2 is 4096 * 60000 = 2.4576e8 > kFarSq

int mfail(int code, const std::string& msg) {
  ssfm_internal_set_error(msg.c_str());
  return code;
}
#define MCK(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return mfail(e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA,                      \
                   std::string(#call) + ": " + cudaGetErrorString(e__));                                 \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// device helpers (mbarrier / 1-D TMA / tcgen05), raw PTX
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for the phase with the given parity.  A pipeline bug must not hang the GPU: after ~2 s of spinning the kernel traps
// (the launch then fails with an error instead of never returning).
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if ((spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, fp16 in, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id); asynchronous until tmem_wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// Wait for this thread's outstanding tcgen05.ld; the registers are passed through so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// Shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes; the next core matrix along K
// is 128 bytes further (leading byte offset), the next 8-row group kGroupBytes further (stride byte offset); version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128u >> 4) << 16;
  d |= (uint64_t)((uint32_t)kGroupBytes >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------
// k_desc_pack: 16 threads per padded row (one per 8-bin chunk); the first of them also writes the |t|^2 block
// ---------------------------------------------------------------------------------------------------------
__global__ void k_desc_pack(const float* __restrict__ desc, const long long* __restrict__ desc_off, const long long* __restrict__ prow_off,
                            int num_images, long long total_prows, __half* __restrict__ packed_q, __half* __restrict__ packed_t,
                            float* __restrict__ norms, int* __restrict__ not_integer) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long prow = gid >> 4;
  const int chunk = (int)(gid & 15);
  float part = 0.f;
  bool real = false;
  const bool in_range = prow < total_prows;
  char *dq = nullptr, *dt = nullptr;
  if (in_range) {
    int lo = 0, hi = num_images;  // image that owns this padded row
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (prow_off[mid] <= prow) lo = mid; else hi = mid;
    }
    const long long r = prow - prow_off[lo];
    const long long n = desc_off[lo + 1] - desc_off[lo];
    __align__(16) __half hq[8], ht[8];
    if (r < n) {
      real = true;
      const float4* src = reinterpret_cast<const float4*>(desc + (desc_off[lo] + r) * kD + chunk * 8);
      const float4 a = src[0], b = src[1];
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      bool bad = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bad = bad || !(v[k] >= 0.f && v[k] <= 255.f && v[k] == rintf(v[k]));
        hq[k] = __float2half_rn(v[k]);
        ht[k] = __float2half_rn(-2.f * v[k]);  // even integers up to 510: exact in fp16
        part += v[k] * v[k];                    // integers < 2^24: exact
      }
      if (bad) *not_integer = 1;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) hq[k] = ht[k] = __float2half_rn(0.f);
    }
    const size_t off = (size_t)(prow >> 3) * kGroupBytes + (size_t)(prow & 7) * 16;
    dq = reinterpret_cast<char*>(packed_q) + off;
    dt = reinterpret_cast<char*>(packed_t) + off;
    *reinterpret_cast<uint4*>(dq + chunk * 128) = *reinterpret_cast<const uint4*>(hq);
    *reinterpret_cast<uint4*>(dt + chunk * 128) = *reinterpret_cast<const uint4*>(ht);
  }
  // |x|^2 of the row: the 16 chunk threads of a row are 16 consecutive lanes
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, 16);
  if (in_range && chunk == 0) {
    norms[prow] = real ? part : 0.f;
    const int n2 = (int)part;  // <= 128 * 255^2 < 2^23
    __align__(16) __half eq[8], et[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) eq[k] = et[k] = __float2half_rn(0.f);
    eq[0] = __float2half_rn(1.f);
    eq[1] = __float2half_rn(64.f);
    eq[2] = __float2half_rn(4096.f);
    et[0] = __float2half_rn(real ? (float)(n2 & 63) : 0.f);
    et[1] = __float2half_rn(real ? (float)((n2 >> 6) & 63) : 0.f);
    et[2] = __float2half_rn(real ? (float)(n2 >> 12) : (float)kPadP2);  // a padded train row can never be a neighbour
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(dq + 16 * 128) = *reinterpret_cast<const uint4*>(eq);
    *reinterpret_cast<uint4*>(dt + 16 * 128) = *reinterpret_cast<const uint4*>(et);
    *reinterpret_cast<uint4*>(dq + 17 * 128) = z;
    *reinterpret_cast<uint4*>(dt + 17 * 128) = z;
  }
}

// ---------------------------------------------------------------------------------------------------------
// k_match_2nn
// ---------------------------------------------------------------------------------------------------------
struct MatchSmem {
  unsigned long long full[kStages], empty[kStages], tfull[kAccBufs], tempty[kAccBufs], qfull, qempty;
  uint32_t tmem_base;
};
constexpr size_t kMatchSmemHeader = 1024;
static_assert(sizeof(MatchSmem) <= kMatchSmemHeader, "header");
struct MergeSlot { float a0, a1; int i0, i1; };  // top-2 of the upper column half of a query row, handed to its partner thread
constexpr size_t kMergeBytes = kColSplit > 1 ? 2 * kQTile * sizeof(MergeSlot) : 0;  // double buffered over work items
constexpr size_t kMatchSmemBytes = kMatchSmemHeader + kQBytes + kStages * kTBytes + kMergeBytes;

// Running two nearest neighbours of one query row.  Everything is kept in the accumulator's domain a = d^2 - |q|^2 (what
// the tensor core delivers: |t|^2 - 2 q.t, an exact integer), so a clean distance costs one min and the bookkeeping is
// min / max / select only.
struct Top2 {
  float a0, a1;  // nearest / second nearest (kFarSq: none yet); a1 is the candidate threshold
  int i0, i1;
};
// cv::batchDistance inserts train index j with `d < dist[K-1]`, shifting while `dist[k] > d`, on d = sqrtf(d^2), i.e. it
// keeps the two smallest (d, j) in lexicographic order.  sqrtf is monotone, and injective on integers below 2^22 (the gap
// sqrt(a+1) - sqrt(a) >= 1/(2*2048) is two ulps there), so while the second neighbour is below 2^22 the float comparisons
// equal the integer ones: four branch-free steps (top2_group).  Above -- only while the second neighbour is still far --
// the floats themselves are compared (top2_group_far).
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }  // one FMNMX3
__device__ __forceinline__ void top2_group(Top2& s, const float* x, int idx) {  // four consecutive train rows
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool first = x[j] < s.a0, cand = x[j] < s.a1;
    s.i1 = cand ? (first ? s.i0 : idx + j) : s.i1;
    s.i0 = first ? idx + j : s.i0;
    s.a1 = fminf(s.a1, fmaxf(s.a0, x[j]));
    s.a0 = fminf(s.a0, x[j]);
  }
}
__device__ __noinline__ Top2 top2_group_far(Top2 s, float x0, float x1, float x2, float x3, float qn, int idx) {
  const float x[4] = {x0, x1, x2, x3};
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    if (!(x[j] < s.a1)) continue;
    const float d = sqrtf(x[j] + qn);
    if (!(d < sqrtf(s.a1 + qn))) continue;
    if (sqrtf(s.a0 + qn) > d) { s.a1 = s.a0; s.i1 = s.i0; s.a0 = x[j]; s.i0 = idx + j; }
    else { s.a1 = x[j]; s.i1 = idx + j; }
  }
  return s;
}

// 32 consecutive train rows against the running top-2 of this thread's query row.  Groups of eight = two of four: one min
// tree per eight and ONE round of warp votes per 32 (against the threshold at the start: it only decreases, so a stale
// one lets a few more groups through, never fewer); the four-groups that hold a candidate are walked.
#ifndef SSFM_MATCH_DIAG
#define SSFM_MATCH_DIAG 0  // timing diagnostics only (WRONG results): 1 = clean scan without the exact insertions, 2 = no scan at all
#endif
__device__ __forceinline__ void top2_scan32(Top2& st, const float* x, int idx0, float qn, float far_a) {
#if SSFM_MATCH_DIAG == 2
  st.a0 = fminf(st.a0, x[0]);
  return;
#endif
  float m4[8], m8[4];
#pragma unroll
  for (int h = 0; h < 8; ++h) m4[h] = fminf(min3f(x[4 * h], x[4 * h + 1], x[4 * h + 2]), x[4 * h + 3]);
#pragma unroll
  for (int g = 0; g < 4; ++g) m8[g] = fminf(m4[2 * g], m4[2 * g + 1]);
  bool gate[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) gate[g] = __any_sync(0xffffffffu, m8[g] < st.a1);
  // the threshold only decreases: once every row of the warp is in the near regime (second neighbour below 2^22) it stays
  // there, and the per-group far-regime vote below is not needed any more
  const bool all_near = __all_sync(0xffffffffu, st.a1 < far_a);
#if SSFM_MATCH_DIAG == 1
#pragma unroll
  for (int g = 0; g < 4; ++g) st.a1 = fminf(st.a1, gate[g] ? m8[g] : st.a1);
  return;
#endif
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (gate[g]) {
#pragma unroll
      for (int h = 2 * g; h < 2 * g + 2; ++h) {
        const bool hit = m4[h] < st.a1;
        if (__any_sync(0xffffffffu, hit)) {
          if (!all_near && __any_sync(0xffffffffu, hit && !(st.a1 < far_a)))
            st = top2_group_far(st, x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3], qn, idx0 + 4 * h);
          else
            top2_group(st, x + 4 * h, idx0 + 4 * h);
        }
      }
    }
  }
}

// The two smallest (sqrtf(d^2), train index) pairs of two disjoint column sets = the two smallest of their four candidates:
// the comparison cv::batchDistance makes, on the floats it makes it on (exact in the near and in the far regime alike).
__device__ __forceinline__ void top2_insert(Top2& s, float a, int i, float qn) {
  const float d = sqrtf(a + qn), d0 = sqrtf(s.a0 + qn), d1 = sqrtf(s.a1 + qn);
  if (d < d0 || (d == d0 && (unsigned)i < (unsigned)s.i0)) {
    s.a1 = s.a0; s.i1 = s.i0; s.a0 = a; s.i0 = i;
  } else if (d < d1 || (d == d1 && (unsigned)i < (unsigned)s.i1)) {
    s.a1 = a; s.i1 = i;
  }
}

// 18 warps are allocated as 20 (granularity of four): 65536 / (20 x 32) = 102 -> 96 registers per thread with the column split
__global__ void __launch_bounds__(kMatchThreads, 1)
k_match_2nn(const __half* __restrict__ packed_q, const __half* __restrict__ packed_t, const float* __restrict__ norms,
            const long long* __restrict__ prow_off, const int* __restrict__ nrows, const int* __restrict__ pair_images,
            const int* __restrict__ work_pair, const int* __restrict__ work_qblock, int num_work, int pair_base,
            const long long* __restrict__ owner_off, int* __restrict__ owner, double ratio) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MatchSmem* sm = reinterpret_cast<MatchSmem*>(smem_raw);
  unsigned char* sQ = smem_raw + kMatchSmemHeader;
  unsigned char* sT = sQ + kQBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(&sm->tfull[b], 1);
      mbar_init(&sm->tempty[b], kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_init(&sm->qfull, 1);
    mbar_init(&sm->qempty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM: 2 accumulator buffers x 2 halves x 128 fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      const char* base_q = reinterpret_cast<const char*>(packed_q);
      const char* base_t = reinterpret_cast<const char*>(packed_t);
      uint32_t it = 0, n = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++n) {
        const int pair = work_pair[w], qb = work_qblock[w];
        const int img0 = pair_images[2 * pair], img1 = pair_images[2 * pair + 1];  // train = image 0, query = image 1
        const int ntiles = (nrows[img0] + kTTile - 1) / kTTile;
        const long long t0 = prow_off[img0], q0 = prow_off[img1] + (long long)qb * kQTile;
        // the first train tiles of this item are requested before its query block: they only need free ring slots, while
        // the query buffer is free only once the previous item's last MMA has retired
        const int qpos = ntiles < 3 ? ntiles - 1 : 2;
        for (int t = 0; t < ntiles; ++t, ++it) {
          if (t == qpos) {
            mbar_wait(&sm->qempty, (n & 1u) ^ 1u);
            mbar_expect_tx(&sm->qfull, kQBytes);
            tma_load_1d(sQ, base_q + (q0 >> 3) * kGroupBytes, kQBytes, &sm->qfull);
          }
          const uint32_t s = it % kStages;
          mbar_wait(&sm->empty[s], ((it / kStages) & 1u) ^ 1u);  // slot free (passes at once the first time round)
          mbar_expect_tx(&sm->full[s], kTBytes);
          tma_load_1d(sT + (size_t)s * kTBytes, base_t + ((t0 >> 3) + (long long)t * (kTTile / 8)) * kGroupBytes, kTBytes, &sm->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      uint32_t it = 0, n = 0;
      const uint64_t dq0 = umma_desc(smem_u32(sQ));
      const uint64_t dq1 = umma_desc(smem_u32(sQ + kQBytes / 2));
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++n) {
        const int pair = work_pair[w];
        const int ntiles = (nrows[pair_images[2 * pair]] + kTTile - 1) / kTTile;
        mbar_wait(&sm->qfull, n & 1u);
        for (int t = 0; t < ntiles; ++t, ++it) {
          const uint32_t s = it % kStages, b = it % kAccBufs;
          mbar_wait(&sm->tempty[b], ((it / kAccBufs) & 1u) ^ 1u);  // the epilogue has drained this accumulator buffer
          mbar_wait(&sm->full[s], (it / kStages) & 1u);            // the train tile has landed
          tc_fence_after();
          const uint64_t dt = umma_desc(smem_u32(sT + (size_t)s * kTBytes));
          const uint32_t acc = tmem + b * (2 * kTTile);
#pragma unroll
          for (int k = 0; k < kKp / 16; ++k)  // K = 16 per instruction = two core matrices = 256 bytes along K
            umma_f16(acc, dq0 + (uint64_t)(k * 16), dt + (uint64_t)(k * 16), kIdesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kKp / 16; ++k)
            umma_f16(acc + kTTile, dq1 + (uint64_t)(k * 16), dt + (uint64_t)(k * 16), kIdesc, k > 0 ? 1u : 0u);
          umma_commit(&sm->empty[s]);   // shared-memory slot reusable once these MMAs have read it
          umma_commit(&sm->tfull[b]);   // accumulators ready for the epilogue
        }
        umma_commit(&sm->qempty);       // query block reusable
      }
    }
  } else {
    // ===== epilogue: one query row per thread =====
    const int quarter = warp & 3;               // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int half = ((warp - 2) >> 2) & 1;     // which M=128 half of the query block
    const int colpart = (warp - 2) >> 3;        // which kTTile / kColSplit columns of every tile
    const int col0 = colpart * (kTTile / kColSplit);
    const int row = half * 128 + quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * kTTile + col0);
    MergeSlot* merge = reinterpret_cast<MergeSlot*>(sT + (size_t)kStages * kTBytes);
    uint32_t it = 0, item = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++item) {
      const int pair = work_pair[w], qb = work_qblock[w];
      const int img0 = pair_images[2 * pair], img1 = pair_images[2 * pair + 1];
      const int n0 = nrows[img0], n1 = nrows[img1];
      const int ntiles = (n0 + kTTile - 1) / kTTile;
      const long long qrow = (long long)qb * kQTile + row;  // row within image 1
      const bool live = qrow < n1;
      const float qn = live ? norms[prow_off[img1] + qrow] : 0.f;
      Top2 st;
      st.a0 = st.a1 = kFarSq;
      st.i0 = st.i1 = -1;
      const float far_a = 4194304.f - qn;  // second neighbour at or above it: d^2 >= 2^22, compare the floats
      for (int t = 0; t < ntiles; ++t, ++it) {
        const uint32_t b = it % kAccBufs;
        mbar_wait(&sm->tfull[b], (it / kAccBufs) & 1u);
        tc_fence_after();
        const uint32_t taddr = lane_addr + b * (2 * kTTile);
#if SSFM_MATCH_DOUBLEBUF
        // two register buffers: the next 32 columns are in flight while the current ones are scanned
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_wait(va);
#pragma unroll 1
        for (int c = 0; c < kEpiChunks; c += 2) {
          tmem_ld32(taddr + (uint32_t)((c + 1) * 32), vb);
          top2_scan32(st, reinterpret_cast<const float*>(va), t * kTTile + col0 + c * 32, qn, far_a);
          tmem_wait(vb);
          if (c + 2 < kEpiChunks) tmem_ld32(taddr + (uint32_t)((c + 2) * 32), va);
          top2_scan32(st, reinterpret_cast<const float*>(vb), t * kTTile + col0 + (c + 1) * 32, qn, far_a);
          if (c + 2 < kEpiChunks) tmem_wait(va);
        }
#else
        // one register buffer: with four epilogue warps per scheduler the other warps cover the TMEM-load latency
        uint32_t va[32];
#pragma unroll 1
        for (int c = 0; c < kEpiChunks; ++c) {
          tmem_ld32(taddr + (uint32_t)(c * 32), va);
          tmem_wait(va);
          top2_scan32(st, reinterpret_cast<const float*>(va), t * kTTile + col0 + c * 32, qn, far_a);
        }
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->tempty[b]);
      }
      // Lowe's ratio test in double, as `matches[i][0].distance < ratio * matches[i][1].distance` evaluates it (:246);
      // m01[trainIdx] = queryIdx with later queries overwriting earlier ones == the maximum query index per train index
      if (kColSplit > 1) {  // the upper column half hands its two candidates to the thread that owns the lower half
        MergeSlot* slot = merge + (size_t)(item & 1u) * kQTile + row;
        if (colpart) *slot = MergeSlot{st.a0, st.a1, st.i0, st.i1};
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // epilogue warps only
        if (colpart == 0) {
          const MergeSlot o = *slot;
          top2_insert(st, o.a0, o.i0, qn);
          top2_insert(st, o.a1, o.i1, qn);
        }
      }
      if (colpart == 0 && live && st.i1 >= 0) {
        const float d0 = sqrtf(st.a0 + qn), d1 = sqrtf(st.a1 + qn);  // exact d^2, IEEE sqrt: the floats cv::BFMatcher returns
        if ((double)d0 < ratio * (double)d1) atomicMax(&owner[owner_off[pair - pair_base] + st.i0], (int)qrow);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// owner[t] >= 0 <=> train keypoint t of the pair is matched (to query owner[t])
__global__ void k_match_count(const int* __restrict__ owner, const long long* __restrict__ owner_off, int npairs, int* __restrict__ counts) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= npairs) return;
  const long long a = owner_off[w], b = owner_off[w + 1];
  int c = 0;
  for (long long i = a + lane; i < b; i += 32) c += owner[i] >= 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[w] = c;
}
// matches of pair w, in increasing order of the image-0 index (the iteration order of the reference's std::map)
__global__ void k_match_write(const int* __restrict__ owner, const long long* __restrict__ owner_off, int npairs,
                              const long long* __restrict__ match_off, int2* __restrict__ matches) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= npairs) return;
  const long long a = owner_off[w], b = owner_off[w + 1];
  long long out = match_off[w];
  for (long long base = a; base < b; base += 32) {
    const long long i = base + lane;
    const int o = i < b ? owner[i] : -1;
    const unsigned m = __ballot_sync(0xffffffffu, o >= 0);
    if (o >= 0) matches[out + __popc(m & ((1u << lane) - 1u))] = make_int2((int)(i - a), o);
    out += __popc(m);
  }
}

// Grow-only device buffer (kept in the engine's match context across calls: no cudaMalloc in a warmed-up call).
template <class T>
struct Buf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t need(size_t n) {
    n = std::max<size_t>(n, 1);
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 8;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct MatchCtx {
  Buf<float> desc, norms;
  Buf<__half> packed_q, packed_t;
  Buf<long long> doff, prow, ooff, moff;
  Buf<int> nrows, pairs, wp, wq, owner, counts, flag;
  Buf<int2> matches;
  cudaEvent_t ev[6] = {};
  bool attr_set = false;
  SsfmMatchStats stats = {};
  ~MatchCtx() {
    desc.release(); norms.release(); packed_q.release(); packed_t.release(); doff.release(); prow.release(); ooff.release(); moff.release();
    nrows.release(); pairs.release(); wp.release(); wq.release(); owner.release(); counts.release(); flag.release(); matches.release();
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};
void match_ctx_delete(void* p) { delete static_cast<MatchCtx*>(p); }

MatchCtx* match_ctx(ssfm_handle h) {
  void** slot = ssfm_internal_ctx_slot(h, match_ctx_delete);
  if (!*slot) *slot = new MatchCtx();
  return static_cast<MatchCtx*>(*slot);
}

}  // namespace

extern "C" int ssfm_match_get_stats(ssfm_handle h, SsfmMatchStats* out) {
  if (!h || !out) return mfail(SSFM_ERR_INVALID, "ssfm_match_get_stats: NULL argument");
  *out = match_ctx(h)->stats;
  return SSFM_OK;
}

extern "C" int ssfm_match_pairs(ssfm_handle h, const SsfmDescriptorBatch* b, int64_t* match_offsets, int32_t* matches, int64_t capacity) {
  if (!h || !b || !match_offsets) return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: NULL argument");
  if (b->num_images < 0 || b->num_pairs < 0 || b->descriptor_length != kD)
    return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: bad sizes (descriptor_length must be 128)");
  if (!(b->ratio > 0.0)) return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: ratio must be > 0");
  const int NI = b->num_images, P = b->num_pairs;
  match_offsets[0] = 0;
  if (P == 0) return SSFM_OK;
  if (!b->desc_offsets || !b->descriptors || !b->pair_images || (capacity > 0 && !matches))
    return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: NULL array");
  if (b->desc_offsets[0] != 0) return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: desc_offsets[0] must be 0");
  const auto wall0 = std::chrono::steady_clock::now();
  std::vector<long long> prow(NI + 1, 0);
  std::vector<int> nrows(std::max(NI, 1), 0);
  for (int i = 0; i < NI; ++i) {
    const long long n = b->desc_offsets[i + 1] - b->desc_offsets[i];
    if (n < 0 || n > 0x3fffffffLL) return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: desc_offsets must be non-decreasing");
    nrows[i] = (int)n;
    prow[i + 1] = prow[i] + (n + kRowPad - 1) / kRowPad * kRowPad;
  }
  for (int p = 0; p < P; ++p) {
    const int i0 = b->pair_images[2 * p], i1 = b->pair_images[2 * p + 1];
    if (i0 < 0 || i1 < 0 || i0 >= NI || i1 >= NI) return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: pair image index out of range");
  }
  const long long rows = b->desc_offsets[NI], prows = prow[NI];
  cudaError_t e0 = cudaSetDevice(ssfm_internal_device(h));
  if (e0 != cudaSuccess) return mfail(SSFM_ERR_CUDA, cudaGetErrorString(e0));
  cudaStream_t st = ssfm_internal_stream(h);
  MatchCtx& c = *match_ctx(h);
  for (auto& e : c.ev)
    if (!e) MCK(cudaEventCreate(&e));
  c.stats = SsfmMatchStats{};

  MCK(c.desc.need((size_t)rows * kD));
  MCK(c.packed_q.need((size_t)prows * kKp));
  MCK(c.packed_t.need((size_t)prows * kKp));
  MCK(c.norms.need((size_t)prows));
  MCK(c.doff.need(NI + 1));
  MCK(c.prow.need(NI + 1));
  MCK(c.nrows.need(NI));
  MCK(c.pairs.need((size_t)2 * P));
  MCK(c.flag.need(1));
  MCK(cudaEventRecord(c.ev[0], st));
  if (rows > 0) MCK(cudaMemcpyAsync(c.desc.p, b->descriptors, sizeof(float) * (size_t)rows * kD, cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(c.doff.p, b->desc_offsets, sizeof(long long) * (NI + 1), cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(c.prow.p, prow.data(), sizeof(long long) * (NI + 1), cudaMemcpyHostToDevice, st));
  if (NI > 0) MCK(cudaMemcpyAsync(c.nrows.p, nrows.data(), sizeof(int) * NI, cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(c.pairs.p, b->pair_images, sizeof(int) * 2 * (size_t)P, cudaMemcpyHostToDevice, st));
  MCK(cudaMemsetAsync(c.flag.p, 0, sizeof(int), st));
  if (prows > 0) {
    const long long threads = prows * 16;
    k_desc_pack<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(c.desc.p, c.doff.p, c.prow.p, NI, prows, c.packed_q.p, c.packed_t.p,
                                                                   c.norms.p, c.flag.p);
    MCK(cudaGetLastError());
  }
  MCK(cudaEventRecord(c.ev[1], st));
  int bad = 0;
  MCK(cudaMemcpyAsync(&bad, c.flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  if (bad)
    return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: descriptors must be integers in [0, 255] stored as float (cv::SIFT); "
                                   "the exactness of the fp16 tensor-core path depends on it");
  {
    float ms = 0.f;
    MCK(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    c.stats.pack_ms = ms;
    c.stats.h2d_bytes = (int64_t)sizeof(float) * rows * kD + (int64_t)sizeof(int) * 2 * P;
  }
  if (!c.attr_set) {
    MCK(cudaFuncSetAttribute(k_match_2nn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemBytes));
    c.attr_set = true;
  }
  const int num_sms = std::max(1, ssfm_internal_num_sms(h));

  // passes of pairs: the owner table (one int per train keypoint per pair) is bounded to 64 Mi entries
  const long long kOwnerCap = 64ll << 20;
  long long total = 0;
  std::vector<long long> ooff, moff;
  std::vector<int> wp, wq, counts;
  for (int p0 = 0; p0 < P;) {
    int p1 = p0;
    long long own = 0;
    ooff.assign(1, 0);
    wp.clear();
    wq.clear();
    while (p1 < P) {
      const int n0 = nrows[b->pair_images[2 * p1]], n1 = nrows[b->pair_images[2 * p1 + 1]];
      if (p1 > p0 && own + n0 > kOwnerCap) break;
      own += n0;
      ooff.push_back(own);
      if (n0 >= 2) {  // with fewer than two train descriptors knnMatch(k=2) has no second neighbour: no match passes the test
        for (int qb = 0; qb * kQTile < n1; ++qb) { wp.push_back(p1); wq.push_back(qb); }
        c.stats.distance_evaluations += (int64_t)n0 * n1;
        c.stats.mma_tiles += (int64_t)((n0 + kTTile - 1) / kTTile) * ((n1 + kQTile - 1) / kQTile);
      }
      ++p1;
    }
    const int np = p1 - p0;
    MCK(c.owner.need((size_t)own));
    MCK(c.ooff.need(np + 1));
    MCK(c.wp.need(wp.size()));
    MCK(c.wq.need(wq.size()));
    MCK(c.counts.need(np));
    MCK(c.moff.need(np + 1));
    MCK(cudaMemsetAsync(c.owner.p, 0xff, sizeof(int) * (size_t)std::max<long long>(own, 1), st));
    MCK(cudaMemcpyAsync(c.ooff.p, ooff.data(), sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
    MCK(cudaEventRecord(c.ev[2], st));
    if (!wp.empty()) {
      MCK(cudaMemcpyAsync(c.wp.p, wp.data(), sizeof(int) * wp.size(), cudaMemcpyHostToDevice, st));
      MCK(cudaMemcpyAsync(c.wq.p, wq.data(), sizeof(int) * wq.size(), cudaMemcpyHostToDevice, st));
      const int grid = (int)std::min<size_t>(wp.size(), (size_t)num_sms);
      MCK(cudaEventRecord(c.ev[2], st));
      k_match_2nn<<<grid, kMatchThreads, kMatchSmemBytes, st>>>(c.packed_q.p, c.packed_t.p, c.norms.p, c.prow.p, c.nrows.p, c.pairs.p, c.wp.p,
                                                               c.wq.p, (int)wp.size(), p0, c.ooff.p, c.owner.p, b->ratio);
      MCK(cudaGetLastError());
      c.stats.knn_launches += 1;
      c.stats.ctas += grid;
    }
    MCK(cudaEventRecord(c.ev[3], st));
    k_match_count<<<(np + 3) / 4, 128, 0, st>>>(c.owner.p, c.ooff.p, np, c.counts.p);
    MCK(cudaGetLastError());
    counts.assign(np, 0);
    MCK(cudaMemcpyAsync(counts.data(), c.counts.p, sizeof(int) * np, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    moff.assign(np + 1, 0);
    for (int k = 0; k < np; ++k) {
      moff[k + 1] = moff[k] + counts[k];
      match_offsets[p0 + k + 1] = total + moff[k + 1];
    }
    if (total + moff[np] > capacity)
      return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: `capacity` is too small (sum over pairs of min(n0, n1) always suffices)");
    if (moff[np] > 0) {
      MCK(c.matches.need((size_t)moff[np]));
      MCK(cudaMemcpyAsync(c.moff.p, moff.data(), sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
      k_match_write<<<(np + 3) / 4, 128, 0, st>>>(c.owner.p, c.ooff.p, np, c.moff.p, c.matches.p);
      MCK(cudaGetLastError());
      MCK(cudaMemcpyAsync(matches + 2 * total, c.matches.p, sizeof(int2) * (size_t)moff[np], cudaMemcpyDeviceToHost, st));
    }
    MCK(cudaEventRecord(c.ev[4], st));
    MCK(cudaStreamSynchronize(st));
    {
      float ms = 0.f;
      MCK(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
      c.stats.knn_ms += ms;
      MCK(cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]));
      c.stats.compact_ms += ms;
      c.stats.d2h_bytes += (int64_t)sizeof(int2) * moff[np] + (int64_t)sizeof(int) * np;
    }
    total += moff[np];
    p0 = p1;
  }
  c.stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  return SSFM_OK;
}
