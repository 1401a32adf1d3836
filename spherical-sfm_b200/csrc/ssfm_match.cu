// ssfm_match.cu -- descriptor matching for many image pairs: the step in front of the
// relative-pose path, and the one dense contraction on it.
//
// Reference: match() (examples/spherical_sfm_tools.cpp:235-251) = cv::BFMatcher (NORM_L2, no cross check) ::knnMatch(
// query = features1.descs, train = features0.descs, k = 2), Lowe's ratio test `d0 < ratio * d1`, and
// `m01[trainIdx] = queryIdx` in query order (a later query overwrites an earlier one on the same train index);
// match_exhaustive() (:575-600) runs it for every image pair under `#pragma omp parallel for`.
//
// B200 mapping.  For one pair the 2-NN search is S = Q T^T (n1 x n0 x 128) followed by a running top-2 per query row of
// d^2 = |q|^2 + |t|^2 - 2 S.  SIFT descriptors are integers 0..255 stored as float (cv::SIFT), so in fp16 they are exact,
// every product is exact in fp32 and every partial sum is an integer below 2^24: the tensor-core result is the exact
// integer d^2 that OpenCV's float accumulation also produces, and dist = sqrtf(d^2) is the same float.  Hence bit-exact
// index pairs, ties included (OpenCV keeps the lower train index on equal float distances: strict '<' in its insertion).
//   k_desc_pack   float descriptors -> fp16 in the tensor-core "core matrix" order (8 rows x 16 bytes contiguous, 16 such
//                 blocks per 8-row group), |x|^2 per row, rows of every image padded to a multiple of 256; a tile of the
//                 packed buffer is then ONE contiguous 1-D TMA bulk copy and needs no swizzle
//   k_match_2nn   one CTA per (pair, 128 query rows): warp 0 = TMA producer (+ TMEM alloc), warp 1 = tcgen05.mma issuer
//                 (M=128, N=256, K=16, fp16 -> fp32 in TMEM, two accumulator buffers = all 512 TMEM columns), warps 2-5 =
//                 epilogue (tcgen05.ld, one query row per thread, running top-2 with OpenCV's insertion rule, ratio test,
//                 atomicMax of the query index on the winning train index = the reference's overwrite order)
//   k_match_count / k_match_write   ordered compaction of the owner table into the Matches list (sorted by index in image 0,
//                 the iteration order of the reference's std::map)
// There is no CPU fallback; the entry fails with SSFM_ERR_NO_DEVICE / SSFM_ERR_CUDA like the rest of the library.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ssfm.h"

extern "C" void ssfm_internal_set_error(const char* msg);  // ssfm_engine.cu
extern "C" int ssfm_internal_device(ssfm_handle h);
extern "C" cudaStream_t ssfm_internal_stream(ssfm_handle h);

namespace {

constexpr int kD = 128;            // descriptor length (SIFT)
constexpr int kQTile = 128;        // query rows per CTA = UMMA M
constexpr int kTTile = 256;        // train rows per MMA tile = UMMA N
constexpr int kRowPad = 256;       // rows of every image are padded to a multiple of this in the packed buffer
constexpr int kGroupBytes = 2048;  // one 8-row group: 16 core matrices (8 rows x 16 B) = 8 x 128 halfs
constexpr int kStages = 2;         // train-tile ring in shared memory (and accumulator buffers in TMEM)
constexpr int kNormBufs = 4;
constexpr int kMatchThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr uint32_t kQBytes = kQTile * kD * 2;  // 32 KB
constexpr uint32_t kTBytes = kTTile * kD * 2;  // 64 KB
constexpr uint32_t kTmemCols = 512;

int mfail(int code, const std::string& msg) {
  ssfm_internal_set_error(msg.c_str());
  return code;
}
#define MCK(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      release_all();                                                                                     \
      return mfail(e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA,                      \
                   std::string(#call) + ": " + cudaGetErrorString(e__));                                 \
    }                                                                                                    \
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
  const long long t0 = clock64();
  for (;;) {
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
    if (clock64() - t0 > 4000000000ll) __trap();
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
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes; the next core matrix along K
// is 128 bytes further (leading byte offset), the next 8-row group 2048 bytes further (stride byte offset); version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128u >> 4) << 16;
  d |= (uint64_t)((uint32_t)kGroupBytes >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTTile >> 3) << 17) | ((uint32_t)(kQTile >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------
// k_desc_pack: one thread per (padded row, 8-column chunk)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_desc_pack(const float* __restrict__ desc, const long long* __restrict__ desc_off, const long long* __restrict__ prow_off,
                            int num_images, long long total_prows, __half* __restrict__ packed, float* __restrict__ norms,
                            int* __restrict__ not_integer) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long prow = gid >> 4;
  const int chunk = (int)(gid & 15);
  float part = 0.f;
  bool real = false;
  if (prow < total_prows) {
    int lo = 0, hi = num_images;  // image that owns this padded row
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (prow_off[mid] <= prow) lo = mid; else hi = mid;
    }
    const long long r = prow - prow_off[lo];
    const long long n = desc_off[lo + 1] - desc_off[lo];
    __align__(16) __half h[8];
    if (r < n) {
      real = true;
      const float4* src = reinterpret_cast<const float4*>(desc + (desc_off[lo] + r) * kD + chunk * 8);
      const float4 a = src[0], b = src[1];
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      bool bad = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bad = bad || !(v[k] >= 0.f && v[k] <= 255.f && v[k] == rintf(v[k]));
        h[k] = __float2half_rn(v[k]);
        part += v[k] * v[k];  // integers < 2^24: exact
      }
      if (bad) *not_integer = 1;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) h[k] = __float2half_rn(0.f);
    }
    char* dst = reinterpret_cast<char*>(packed) + (prow >> 3) * kGroupBytes + chunk * 128 + (prow & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
  }
  // |x|^2 of the row: the 16 chunk threads of a row are 16 consecutive lanes
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, 16);
  if (prow < total_prows && chunk == 0) norms[prow] = real ? part : INFINITY;  // a padded train row can never be a neighbour
}

// ---------------------------------------------------------------------------------------------------------
// k_match_2nn
// ---------------------------------------------------------------------------------------------------------
struct MatchSmem {
  unsigned long long full[kStages], empty[kStages], tfull[kStages], tempty[kStages], qfull;
  uint32_t tmem_base;
  uint32_t pad_[13];
  float tnorm[kNormBufs][kTTile];  // |t|^2 of the tile's train rows
};
constexpr size_t kMatchSmemHeader = ((sizeof(MatchSmem) + 1023) / 1024) * 1024;
constexpr size_t kMatchSmemBytes = kMatchSmemHeader + kQBytes + kStages * kTBytes;

__global__ void __launch_bounds__(kMatchThreads, 1)
k_match_2nn(const __half* __restrict__ packed, const float* __restrict__ norms, const long long* __restrict__ prow_off,
            const int* __restrict__ nrows, const int* __restrict__ pair_images, const int* __restrict__ work_pair,
            const int* __restrict__ work_qblock, int pair_base, const long long* __restrict__ owner_off, int* __restrict__ owner,
            double ratio) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MatchSmem* sm = reinterpret_cast<MatchSmem*>(smem_raw);
  unsigned char* sQ = smem_raw + kMatchSmemHeader;
  unsigned char* sT = sQ + kQBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = work_pair[blockIdx.x];
  const int qb = work_qblock[blockIdx.x];
  const int img0 = pair_images[2 * pair], img1 = pair_images[2 * pair + 1];  // train = image 0, query = image 1
  const int n0 = nrows[img0], n1 = nrows[img1];
  const long long t0 = prow_off[img0], q0 = prow_off[img1] + (long long)qb * kQTile;
  const int ntiles = (n0 + kTTile - 1) / kTTile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], 1);
      mbar_init(&sm->tfull[s], 1);
      mbar_init(&sm->tempty[s], 4);  // one arrival per epilogue warp
    }
    mbar_init(&sm->qfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM: both accumulator buffers (2 x 256 fp32 columns x 128 lanes)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      const char* base = reinterpret_cast<const char*>(packed);
      mbar_expect_tx(&sm->qfull, kQBytes);
      tma_load_1d(sQ, base + (q0 >> 3) * kGroupBytes, kQBytes, &sm->qfull);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&sm->empty[s], (uint32_t)(((t / kStages) & 1) ^ 1));  // slot free (passes at once the first time round)
        mbar_expect_tx(&sm->full[s], kTBytes + (uint32_t)sizeof(float) * kTTile);
        tma_load_1d(sT + (size_t)s * kTBytes, base + ((t0 >> 3) + (long long)t * (kTTile / 8)) * kGroupBytes, kTBytes, &sm->full[s]);
        tma_load_1d(sm->tnorm[t % kNormBufs], norms + t0 + (long long)t * kTTile, (uint32_t)sizeof(float) * kTTile, &sm->full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      mbar_wait(&sm->qfull, 0u);
      const uint64_t dq = umma_desc(smem_u32(sQ));
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        const uint32_t ph = (uint32_t)((t / kStages) & 1);
        mbar_wait(&sm->tempty[s], ph ^ 1u);  // the epilogue has drained this accumulator buffer
        mbar_wait(&sm->full[s], ph);         // the train tile has landed
        tc_fence_after();
        const uint64_t dt = umma_desc(smem_u32(sT + (size_t)s * kTBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)  // K = 16 per instruction = two core matrices = 256 bytes along K
          umma_f16(tmem + (uint32_t)s * kTTile, dq + (uint64_t)(k * 16), dt + (uint64_t)(k * 16), kIdesc, k > 0 ? 1u : 0u);
        umma_commit(&sm->empty[s]);  // shared-memory slot reusable once these MMAs have read it
        umma_commit(&sm->tfull[s]);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: one query row per thread =====
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int row = quarter * 32 + lane;
    const long long qrow = (long long)qb * kQTile + row;  // row within image 1
    const bool live = qrow < n1;
    const float qn = live ? norms[q0 + row] : 0.f;
    // cv::batchDistance initialises dist = FLT_MAX, idx = -1 and inserts with `d < dist[K-1]`, shifting while `dist[k] > d`
    float d0 = FLT_MAX, d1 = FLT_MAX;
    int i0 = -1, i1 = -1;
    // Exact d^2 of the two kept neighbours.  sqrtf is monotone, so a candidate with d^2 >= q1 has sqrtf(d^2) >= d1 and can
    // never pass OpenCV's `d < dist[1]`: that is the one-compare fast path; only candidates below q1 pay for the sqrt.
    float q0d = INFINITY, q1d = INFINITY;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % kStages;
      mbar_wait(&sm->tfull[s], (uint32_t)((t / kStages) & 1));
      tc_fence_after();
      const float* tn = sm->tnorm[t % kNormBufs];
#pragma unroll 1
      for (int c = 0; c < kTTile / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * kTTile + c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dot = __uint_as_float(v[j]);
          const float dsq = fmaf(-2.f, dot, tn[c * 32 + j]) + qn;  // exact integers (< 2^24); +inf for padded train rows
          if (dsq < q1d) {
            const float d = sqrtf(dsq);  // IEEE sqrt, the float cv::BFMatcher compares
            if (d < d1) {
              const int idx = t * kTTile + c * 32 + j;
              if (d0 > d) { d1 = d0; i1 = i0; q1d = q0d; d0 = d; i0 = idx; q0d = dsq; }
              else { d1 = d; i1 = idx; q1d = dsq; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->tempty[s]);
    }
    // Lowe's ratio test in double, as `matches[i][0].distance < ratio * matches[i][1].distance` evaluates it (:246);
    // m01[trainIdx] = queryIdx with later queries overwriting earlier ones == the maximum query index per train index
    if (live && i1 >= 0 && (double)d0 < ratio * (double)d1) atomicMax(&owner[owner_off[pair - pair_base] + i0], (int)qrow);
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

template <class T>
struct Buf {
  T* p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
  void release() { if (p) cudaFree(p); p = nullptr; }
};

}  // namespace

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

  Buf<float> d_desc, d_norms;
  Buf<__half> d_packed;
  Buf<long long> d_doff, d_prow, d_ooff, d_moff;
  Buf<int> d_nrows, d_pairs, d_wp, d_wq, d_owner, d_counts, d_flag;
  Buf<int2> d_matches;
  auto release_all = [&]() {
    d_desc.release(); d_norms.release(); d_packed.release(); d_doff.release(); d_prow.release(); d_ooff.release(); d_moff.release();
    d_nrows.release(); d_pairs.release(); d_wp.release(); d_wq.release(); d_owner.release(); d_counts.release(); d_flag.release();
    d_matches.release();
  };
  MCK(d_desc.alloc((size_t)rows * kD));
  MCK(d_packed.alloc((size_t)prows * kD));
  MCK(d_norms.alloc((size_t)prows + kTTile));
  MCK(d_doff.alloc(NI + 1));
  MCK(d_prow.alloc(NI + 1));
  MCK(d_nrows.alloc(NI));
  MCK(d_pairs.alloc((size_t)2 * P));
  MCK(d_flag.alloc(1));
  if (rows > 0) MCK(cudaMemcpyAsync(d_desc.p, b->descriptors, sizeof(float) * (size_t)rows * kD, cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(d_doff.p, b->desc_offsets, sizeof(long long) * (NI + 1), cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(d_prow.p, prow.data(), sizeof(long long) * (NI + 1), cudaMemcpyHostToDevice, st));
  if (NI > 0) MCK(cudaMemcpyAsync(d_nrows.p, nrows.data(), sizeof(int) * NI, cudaMemcpyHostToDevice, st));
  MCK(cudaMemcpyAsync(d_pairs.p, b->pair_images, sizeof(int) * 2 * (size_t)P, cudaMemcpyHostToDevice, st));
  MCK(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
  if (prows > 0) {
    const long long threads = prows * 16;
    k_desc_pack<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_desc.p, d_doff.p, d_prow.p, NI, prows, d_packed.p, d_norms.p, d_flag.p);
    MCK(cudaGetLastError());
  }
  int bad = 0;
  MCK(cudaMemcpyAsync(&bad, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  d_desc.release();
  if (bad) {
    release_all();
    return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: descriptors must be integers in [0, 255] stored as float (cv::SIFT); "
                                   "the exactness of the fp16 tensor-core path depends on it");
  }
  MCK(cudaFuncSetAttribute(k_match_2nn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemBytes));

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
      if (n0 >= 2)  // with fewer than two train descriptors knnMatch(k=2) has no second neighbour: no match passes the test
        for (int qb = 0; qb * kQTile < n1; ++qb) { wp.push_back(p1); wq.push_back(qb); }
      ++p1;
    }
    const int np = p1 - p0;
    d_owner.release(); d_ooff.release(); d_wp.release(); d_wq.release(); d_counts.release(); d_moff.release(); d_matches.release();
    MCK(d_owner.alloc((size_t)own));
    MCK(d_ooff.alloc(np + 1));
    MCK(d_wp.alloc(wp.size()));
    MCK(d_wq.alloc(wq.size()));
    MCK(d_counts.alloc(np));
    MCK(d_moff.alloc(np + 1));
    MCK(cudaMemsetAsync(d_owner.p, 0xff, sizeof(int) * (size_t)std::max<long long>(own, 1), st));
    MCK(cudaMemcpyAsync(d_ooff.p, ooff.data(), sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
    if (!wp.empty()) {
      MCK(cudaMemcpyAsync(d_wp.p, wp.data(), sizeof(int) * wp.size(), cudaMemcpyHostToDevice, st));
      MCK(cudaMemcpyAsync(d_wq.p, wq.data(), sizeof(int) * wq.size(), cudaMemcpyHostToDevice, st));
      k_match_2nn<<<(unsigned)wp.size(), kMatchThreads, kMatchSmemBytes, st>>>(d_packed.p, d_norms.p, d_prow.p, d_nrows.p, d_pairs.p, d_wp.p,
                                                                              d_wq.p, p0, d_ooff.p, d_owner.p, b->ratio);
      MCK(cudaGetLastError());
    }
    k_match_count<<<(np + 3) / 4, 128, 0, st>>>(d_owner.p, d_ooff.p, np, d_counts.p);
    MCK(cudaGetLastError());
    counts.assign(np, 0);
    MCK(cudaMemcpyAsync(counts.data(), d_counts.p, sizeof(int) * np, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    moff.assign(np + 1, 0);
    for (int k = 0; k < np; ++k) {
      moff[k + 1] = moff[k] + counts[k];
      match_offsets[p0 + k + 1] = total + moff[k + 1];
    }
    if (total + moff[np] > capacity) {
      release_all();
      return mfail(SSFM_ERR_INVALID, "ssfm_match_pairs: `capacity` is too small (sum over pairs of min(n0, n1) always suffices)");
    }
    if (moff[np] > 0) {
      MCK(d_matches.alloc((size_t)moff[np]));
      MCK(cudaMemcpyAsync(d_moff.p, moff.data(), sizeof(long long) * (np + 1), cudaMemcpyHostToDevice, st));
      k_match_write<<<(np + 3) / 4, 128, 0, st>>>(d_owner.p, d_ooff.p, np, d_moff.p, d_matches.p);
      MCK(cudaGetLastError());
      MCK(cudaMemcpyAsync(matches + 2 * total, d_matches.p, sizeof(int2) * (size_t)moff[np], cudaMemcpyDeviceToHost, st));
      MCK(cudaStreamSynchronize(st));
    }
    total += moff[np];
    p0 = p1;
  }
  release_all();
  return SSFM_OK;
}
