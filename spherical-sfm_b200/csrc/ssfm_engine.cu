// ssfm_engine.cu -- host side of libssfm_b200.so: the extern "C" layer declared in include/ssfm.h,
// device-memory management and the round scheduler that drives the kernels in ssfm_kernels.cuh.
//
// Execution model (one handle = one GPU, one stream):
//   upload : H2D of the caller's RayPair memory (float64 AoS) + k_pack into float4 SoA planes
//   run    : for each pass of <= kMaxPassPairs pairs
//              k_init_pairs, k_finish_trivial
//              repeat until no pair is active ("look-ahead rounds"):
//                k_sample_solve  : every active pair speculates its next <= cap iterations
//                k_score_rounds  : FP32 scoring of all those hypotheses (the hot kernel)
//                k_chain         : FP64 certification + sequential LO-MSAC logic, termination,
//                                  compaction of the still-active pairs
//              (one 4-byte D2H per round: the number of active pairs)
//   download: D2H of the per-pair result table (+ optional inlier flags)
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ssfm_kernels.cuh"
#include "ssfm_preemptive.cuh"
#include "ssfm_sixpt_kernels.cuh"

using namespace ssfm;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define SSFM_CK(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      return fail(e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA,             \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                        \
    }                                                                                          \
  } while (0)

constexpr int kMaxPassPairs = 131072;
constexpr int kPipelinePassPairs = 16384;  // pass size when the upload is pipelined with the compute
constexpr int kMaxUploadChunks = 4096;
constexpr int kRoundCap = 256;
constexpr int kDeferMinPairs = 4096;  // batches below this run the LO refits inline (measured crossover)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// Environment knobs.  They exist for measurements and tests (A/B runs of tools/gpu_*.sh, the pre-filter margin test), not
// for users: every one of them is read in ONE place, read_knobs(), once per public call (refresh_knobs), and the rest of
// the engine only sees this struct.
// ---------------------------------------------------------------------------------------------------------
struct Knobs {
  float cand_margin = 2e-4f;   // SSFM_CAND_MARGIN      relative slack of the FP32 pre-filter
  bool no_unitz = false;       // SSFM_NO_UNITZ         always use the general (z != 1) scoring planes
  bool no_xy64 = false;        // SSFM_NO_XY64          exact passes stream the 48-byte records instead of the compact plane
  bool no_small_stage = false; // SSFM_NO_SMALL_STAGE   k_refit_small gathers through L2 on every pass
  bool no_pipeline = false;    // SSFM_NO_PIPELINE      blocking upload instead of the chunk-pipelined one
  bool no_defer = false;       // SSFM_NO_DEFER         run the LO refits inline whatever the batch size
  bool no_handover = false;    // SSFM_NO_HANDOVER      stragglers stay on their thread
  int handover_at = 0;         // SSFM_HANDOVER         iterations before a small refit moves to a warp (0: kHandover)
  int defer_min_pairs = -1;    // SSFM_DEFER_MIN_PAIRS  batches below this run the LO refits inline (-1: kDeferMinPairs)
  int refit_threads_min = 0;   // SSFM_REFIT_THREADS_MIN
  int sixpt_lo_r = 0;          // SSFM_SIXPT_LO_R       look-ahead of the six-point LO-MSAC path (0: adaptive)
  int first_chunk = 2048;      // SSFM_FIRST_CHUNK      pairs in the first upload chunk
  int round_cap = 0;           // SSFM_ROUND_CAP        look-ahead of rounds >= 1 (0: kRoundCap)
  int first_cap = 0;           // SSFM_FIRST_CAP        look-ahead of round 0 (0: round_up32(min_num_iterations))
  int workers = 0;             // SSFM_WORKERS          streams over the pair list (0: automatic)
  int e2e_workers = 0;         // SSFM_E2E_WORKERS      ... while an upload is in flight
  int e2e_parts = 0;           // SSFM_E2E_PARTS        parts dealt to those workers (0: 3 per worker)
};

static Knobs read_knobs() {
  Knobs k;
  auto flag = [](const char* name) { return getenv(name) != nullptr; };
  auto num = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  if (const char* e = getenv("SSFM_CAND_MARGIN")) k.cand_margin = (float)atof(e);
  k.no_unitz = flag("SSFM_NO_UNITZ");
  k.no_xy64 = flag("SSFM_NO_XY64");
  k.no_small_stage = flag("SSFM_NO_SMALL_STAGE");
  k.no_pipeline = flag("SSFM_NO_PIPELINE");
  k.no_defer = flag("SSFM_NO_DEFER");
  k.no_handover = flag("SSFM_NO_HANDOVER");
  k.handover_at = num("SSFM_HANDOVER", 0);
  k.defer_min_pairs = num("SSFM_DEFER_MIN_PAIRS", -1);
  k.refit_threads_min = num("SSFM_REFIT_THREADS_MIN", 0);
  k.sixpt_lo_r = num("SSFM_SIXPT_LO_R", 0);
  k.first_chunk = std::max(256, num("SSFM_FIRST_CHUNK", 2048));
  k.round_cap = num("SSFM_ROUND_CAP", 0);
  k.first_cap = num("SSFM_FIRST_CAP", 0);
  k.workers = num("SSFM_WORKERS", 0);
  k.e2e_workers = num("SSFM_E2E_WORKERS", 0);
  k.e2e_parts = num("SSFM_E2E_PARTS", 0);
  return k;
}

struct Worker;
struct ssfm_engine {
  Knobs knobs;  // refreshed at the top of every public call (refresh_knobs)
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};
  cudaEvent_t ev_tables = nullptr;  // offsets (+ upload-time flags) are in HBM: every worker stream waits on it
  // resident batch
  int P = 0;
  long long M = 0;
  std::vector<long long> h_offsets;
  const double* d_rays = nullptr;  // either rays_own.p or the caller's device pointer
  DevBuf<double> rays_own;
  DevBuf<float4> u4, v4, uv4;  // uv4: always; u4 / v4: only once a batch with z != 1 rays shows up (ensure_general_planes)
  DevBuf<double> xy64;         // (u.x, u.y, v.x, v.y) in float64: what the chain's exact passes stream when every z == 1
  std::mutex general_mu;
  bool unit_z = false;
  // ssfm_upload_matches: keypoints / pair table / matches / Kinv (grow-only, kept across calls)
  DevBuf<float> m_kp;
  DevBuf<long long> m_kpoff;
  DevBuf<int> m_pairs, m_matches;
  DevBuf<double> m_kinv;
  bool matches_pending_check = false;  // a pipelined match upload whose index-range flag has not been read yet
  DevBuf<long long> offsets;
  DevBuf<float> rays_f32;  // staging for SSFM_RAYS_F32 host input (24 bytes per correspondence)
  bool resident = false;
  // pipelined upload (ssfm_estimate_pairs): one event + one unit-z flag per pass of pairs
  std::vector<cudaEvent_t> up_ev;
  std::vector<cudaEvent_t> copy_ev;   // chunk k has landed in HBM (recorded on `stream`, waited on by `stream_pack`)
  cudaStream_t stream_pack = nullptr;  // high priority: the per-chunk pack / ray-build kernels.  On the copy stream they
                                       // would hold the NEXT chunk's H2D back until SMs free up under the workers' kernels.
  std::vector<int> up_bounds;  // pair bounds of the upload chunks (empty: not pipelined)
  DevBuf<int> up_flags;
  int* h_up_flags = nullptr;  // pinned
  // run buffers live in the workers (one stream each; see ssfm_run)
  DevBuf<int> counts;  // upload-time flags
  DevBuf<SsfmPairResult> results;
  DevBuf<unsigned char> flags;
  struct Worker* workers = nullptr;
  int num_workers = 0;
  bool have_results = false;
  SsfmRunStats stats = {};
  int* h_count = nullptr;  // pinned
  // context of the descriptor-matching translation unit (ssfm_match.cu), deleted with the engine
  void* match_ctx = nullptr;
  void (*match_ctx_delete)(void*) = nullptr;
};

// One worker = one stream + its own scratch.  ssfm_run splits the pair list between the workers and
// drives them from separate host threads, so one worker's latency-bound FP64 chain/refit kernels
// overlap the other worker's FP32-bound scoring kernel on the same GPU.
struct Worker {
  cudaStream_t stream = nullptr;     // sampling/solving and FP32 scoring (bulk, normal priority)
  cudaStream_t stream_hi = nullptr;  // FP64 chain + refit waves (small latency-bound grids, high priority:
                                     // they slip in between another worker's scoring CTAs)
  cudaEvent_t ev[4] = {};
  int* h_count = nullptr;  // pinned
  DevBuf<PairState> states;
  DevBuf<uint32_t> mt;
  DevBuf<int> active0, active1, ident, navail, list_a, list_b, counts, parked0, parked1;
  DevBuf<double> models, lm_E;
  DevBuf<float> s32, s32m;
  DevBuf<unsigned long long> counters;
  DevBuf<unsigned char> has;  // pre-emptive driver: hypothesis-has-a-model flags
  DevBuf<SixState> six_states;  // six-point estimator
  DevBuf<SixLoState> six_lo_states;  // ... under LO-MSAC
  DevBuf<uint32_t> six_it;           // iteration number of every pair's look-ahead slot 0
  DevBuf<double> six_M;         // the ten cubics of every look-ahead sample (global scratch of k_sixpt_sample_solve)
  DevBuf<double> small_stage;   // k_refit_small: each lane's contiguous copy of its refit's correspondences
  DevBuf<LMState> lm_states;    // stragglers handed from k_refit_small to k_refit_long
  DevBuf<int> long_list;
  DevBuf<int> six_nm, pk_id, pk_count;
  DevBuf<float> pk_G;
  // outputs of the last run
  double solve_ms = 0, score_ms = 0, chain_ms = 0;
  int rounds = 0, launches = 0;
  long long score_launches = 0, refit_waves = 0;
  unsigned long long hc[32] = {};
  int rc = SSFM_OK;
  std::string err;
  void release() {
    states.release(); mt.release(); active0.release(); active1.release(); ident.release(); navail.release(); list_a.release();
    list_b.release(); counts.release(); parked0.release(); parked1.release(); models.release(); lm_E.release();
    s32.release(); s32m.release(); counters.release(); has.release();
    lm_states.release(); long_list.release(); small_stage.release();
    six_states.release(); six_lo_states.release(); six_it.release(); six_M.release(); six_nm.release(); pk_id.release(); pk_count.release(); pk_G.release();
  }
};

namespace {

constexpr int kMaxWorkers = 4;

static void refresh_knobs(ssfm_engine* h) {
  if (h) h->knobs = read_knobs();
}

Params make_params(const SsfmOptions& o, const Knobs& knobs) {
  Params P;
  P.min_iters = o.min_num_iterations;
  P.max_iters = o.max_num_iterations;
  P.eta = 1.0 - o.success_probability;
  P.thr2 = o.squared_inlier_threshold;
  P.seed = o.random_seed;
  P.num_lo_steps = o.num_lo_steps;
  P.thr_mult = o.threshold_multiplier;
  P.num_lsq_iters = o.num_lsq_iterations;
  P.min_sample_mult = o.min_sample_multiplicator;
  P.non_min_mult = o.non_min_sample_multiplier;
  P.lo_start = o.lo_starting_iterations;
  P.final_lsq = o.final_least_squares;
  P.solver = o.solver;
  P.driver = o.driver;
  P.inward = o.inward;
  P.fixed_budget = o.fixed_budget;
  P.fixed_prob = o.fixed_prob_success;
  P.first_pair_id = o.first_pair_id;
  P.min_points = o.min_num_points;
  P.preempt_block = o.preemptive_block;
  P.sixpt_focal_scoring = o.sixpt_focal_scoring;
  P.skip_complex = o.complex_root_models == SSFM_COMPLEX_SKIP;
  P.cand_margin = knobs.cand_margin;
  P.inline_small_max = 0;
  P.inline_handover = 0;
  return P;
}

int check_options(const SsfmOptions* o) {
  if (!o) return fail(SSFM_ERR_INVALID, "options is NULL");
  if (o->solver < 0 || o->solver > 3) return fail(SSFM_ERR_INVALID, "unknown solver kind");
  if (o->solver == SSFM_SOLVER_SIXPT_FOCAL && o->driver != SSFM_DRIVER_VANILLA_MSAC && o->driver != SSFM_DRIVER_LO_MSAC)
    return fail(SSFM_ERR_INVALID, "the batched six-point shared-focal path runs SSFM_DRIVER_VANILLA_MSAC (config C4) or SSFM_DRIVER_LO_MSAC");
  if (o->driver < 0 || o->driver > 3) return fail(SSFM_ERR_INVALID, "unknown driver kind");
  if (o->complex_root_models != SSFM_COMPLEX_CANONICAL && o->complex_root_models != SSFM_COMPLEX_SKIP)
    return fail(SSFM_ERR_INVALID, "unknown complex_root_models");
  if (o->driver == SSFM_DRIVER_PREEMPTIVE && (o->fixed_budget <= 0 || o->fixed_budget > 8192 || o->preemptive_block <= 0))
    return fail(SSFM_ERR_INVALID, "pre-emptive driver needs 0 < fixed_budget <= 8192 hypotheses and preemptive_block > 0");
  if (!(o->squared_inlier_threshold > 0.0)) return fail(SSFM_ERR_INVALID, "squared_inlier_threshold must be > 0");
  if (o->driver == SSFM_DRIVER_MSAC_FIXED && o->fixed_budget <= 0)
    return fail(SSFM_ERR_INVALID, "fixed_budget must be > 0 for the fixed-budget MSAC driver");
  return SSFM_OK;
}

template <int KIND>
void launch_solve(ssfm_engine* h, Worker& w, const Params& P, int pair0, const int* active, int count, int cap, int R) {
  // a 32-slot round (the fixed-budget legacy driver) would leave half of every 64-thread block idle and, with the block count
  // fixed by the pair count, cost extra waves: one warp per block there
  const int threads = (cap <= 32 && SSFM_SOLVE_SYNC == 0) ? 32 : kSolveThreads;
  dim3 grid(count, (cap + threads - 1) / threads);
  k_sample_solve<KIND><<<grid, threads, 0, w.stream>>>(P, h->d_rays, h->offsets.p, pair0, active, w.navail.p, w.states.p, R,
                                                  w.models.p);
}

struct RunCfg {
  int first_cap, round_cap, R;
  int small_refit_threads_min;  // waves with more small refits than this use one thread per refit
  bool defer;
  bool handover;  // stragglers of the one-thread-per-refit kernel continue on a warp
  int handover_at;
  float thr32;
};

#define SSFM_WCK(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      w.err = std::string(#call) + ": " + cudaGetErrorString(e__);                             \
      return e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA;                  \
    }                                                                                          \
  } while (0)

// The general FP32 planes (u.xyz / v.xyz) for correspondences [c0, c1): built lazily, on the stream that needs them.
int ensure_general_planes(ssfm_engine* h, cudaStream_t st, long long c0, long long c1, std::string* err) {
  std::lock_guard<std::mutex> lock(h->general_mu);
  const size_t m = (size_t)std::max<long long>(h->M, 1);
  if (h->u4.cap < m || h->v4.cap < m) {
    // (re)allocation invalidates what was packed before: callers always pack the range they are about to read
    cudaError_t e = h->u4.ensure(m);
    if (e == cudaSuccess) e = h->v4.ensure(m);
    if (e != cudaSuccess) {
      *err = std::string("general planes: ") + cudaGetErrorString(e);
      return e == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA;
    }
  }
  if (c1 > c0) k_pack_general<<<(unsigned)((c1 - c0 + 255) / 256), 256, 0, st>>>(h->d_rays + 6 * c0, c1 - c0, h->u4.p + c0, h->v4.p + c0);
  return SSFM_OK;
}

struct PassDesc {
  int pair0, np;
  int pipelined;  // 1: the upload is still in flight; round 0 is launched chunk by chunk as the rays arrive
};

// All rounds of the given passes on worker w's streams.
int run_range(ssfm_engine* h, Worker& w, const Params& P, const RunCfg& cfg, const std::vector<PassDesc>& passes) {
  SSFM_WCK(cudaSetDevice(h->device));
  const int first_cap = cfg.first_cap, round_cap = cfg.round_cap, R = cfg.R;
  const bool defer = cfg.defer;
  const float thr32 = cfg.thr32;
  w.solve_ms = w.score_ms = w.chain_ms = 0;
  w.rounds = w.launches = 0;
  w.score_launches = w.refit_waves = 0;
  SSFM_WCK(w.counters.ensure(32));
  SSFM_WCK(w.counts.ensure(8));
  SSFM_WCK(cudaStreamWaitEvent(w.stream, h->ev_tables, 0));
  SSFM_WCK(cudaStreamWaitEvent(w.stream_hi, h->ev_tables, 0));
  SSFM_WCK(cudaMemsetAsync(w.counters.p, 0, 32 * sizeof(unsigned long long), w.stream));
  cudaEvent_t evA = w.ev[0], evB = w.ev[1], evC = w.ev[2], evD = w.ev[3];
  SSFM_WCK(cudaFuncSetAttribute(k_refit_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRefitBigSmem));
  int launches = 0;
  for (const PassDesc& pd : passes) {
    const int pair0 = pd.pair0, np = pd.np;
    bool unit_z = h->unit_z;
    const long long c0 = h->h_offsets[pair0], c1 = h->h_offsets[pair0 + np];
    const size_t mpass = (size_t)std::max<long long>(c1 - c0, 1);
    if (pd.pipelined && (P.solver == SSFM_SOLVER_SIXPT_FOCAL || P.driver == SSFM_DRIVER_PREEMPTIVE)) {
      // these paths do not launch chunk by chunk: wait for the whole upload of this pass
      bool all_unit = true;
      for (size_t k = 0; k + 1 < h->up_bounds.size(); ++k) {
        const int q0 = std::max(h->up_bounds[k], pair0), q1 = std::min(h->up_bounds[k + 1], pair0 + np);
        if (q1 <= q0) continue;
        SSFM_WCK(cudaEventSynchronize(h->up_ev[k]));
        all_unit = all_unit && h->h_up_flags[k] == 0;
      }
      unit_z = all_unit && !h->knobs.no_unitz;
    }
    if (!unit_z && (P.solver == SSFM_SOLVER_SIXPT_FOCAL || !pd.pipelined)) {
      if (int rc = ensure_general_planes(h, w.stream, c0, c1, &w.err)) return rc;
    }
    if (P.solver == SSFM_SOLVER_SIXPT_FOCAL) {
      // Six-point shared-focal estimator under VanillaMSAC (config C4) or LO-MSAC: look-ahead rounds like the 3-point
      // path, in sub-passes bounded by the model table (15 models x 16 doubles per look-ahead slot).
      const bool six_lo = P.driver == SSFM_DRIVER_LO_MSAC;
      // LO-MSAC: the LocalOptimization waves are one warp per parked pair, so they want many pairs in flight; a shorter
      // look-ahead (down to 64 slots instead of 256, when the pass has more than 2048 pairs) lets up to four times as many pairs share the
      // same 1 GB model table.
      int r_lo = cfg.R;
      while (six_lo && r_lo > 64 && (long long)np * r_lo > 2048LL * cfg.R) r_lo >>= 1;  // only when the pass has the pairs to fill it
      if (h->knobs.sixpt_lo_r > 0) r_lo = std::max(32, std::min(cfg.R, h->knobs.sixpt_lo_r & ~31));
      const int R = six_lo ? r_lo : cfg.R;
      const int first_cap = std::min(cfg.first_cap, R), round_cap = std::min(cfg.round_cap, R);
      const int kSub = 2048 * (cfg.R / R);
      if (six_lo) {
        SSFM_WCK(w.list_a.ensure(mpass + 16));
        SSFM_WCK(w.list_b.ensure(mpass + 16));
        SSFM_WCK(w.mt.ensure((size_t)std::min(np, kSub) * 625));
        SSFM_WCK(w.parked0.ensure(std::min(np, kSub)));
        SSFM_WCK(w.parked1.ensure(std::min(np, kSub)));
      }
      for (int s0 = 0; s0 < np; s0 += kSub) {
        const int q0 = pair0 + s0, nq = std::min(kSub, np - s0);
        if (six_lo) SSFM_WCK(w.six_lo_states.ensure(nq));
        else SSFM_WCK(w.six_states.ensure(nq));
        SSFM_WCK(w.six_it.ensure(nq));
        SSFM_WCK(w.active0.ensure(nq));
        SSFM_WCK(w.active1.ensure(nq));
        SSFM_WCK(w.navail.ensure(nq));
        SSFM_WCK(w.models.ensure((size_t)nq * R * kSixMaxModels * kSixRecord));
        SSFM_WCK(w.six_nm.ensure((size_t)nq * R));
        SSFM_WCK(w.six_M.ensure(((size_t)nq * R + kSixSamplesPerBlock) * sixc::kMSize));
        SSFM_WCK(w.s32m.ensure((size_t)nq * R * kSixSlotModels));
        SSFM_WCK(w.pk_G.ensure((size_t)nq * R * kSixMaxModels * 9));
        SSFM_WCK(w.pk_id.ensure((size_t)nq * R * kSixMaxModels));
        SSFM_WCK(w.pk_count.ensure(nq));
        SSFM_WCK(cudaMemsetAsync(w.pk_count.p, 0, sizeof(int) * nq, w.stream));
        if (six_lo) {
          k_sixpt_lo_init<<<(nq + 127) / 128, 128, 0, w.stream>>>(P, h->offsets.p, q0, nq, w.six_lo_states.p, w.mt.p, w.active0.p,
                                                                  w.navail.p, first_cap, w.counts.p, w.six_it.p);
          k_sixpt_lo_trivial<<<(nq + 127) / 128, 128, 0, w.stream>>>(P, h->offsets.p, q0, nq, w.six_lo_states.p, h->results.p + q0,
                                                                     h->flags.p);
          launches += 1;
        } else {
          k_sixpt_init<<<(nq + 127) / 128, 128, 0, w.stream>>>(P, h->offsets.p, q0, nq, w.six_states.p, w.active0.p, w.navail.p,
                                                               first_cap, w.counts.p, w.six_it.p);
        }
        launches += 1;
        int count = nq;
        int* act = w.active0.p;
        int* act_next = w.active1.p;
        int round = 0;
        SSFM_WCK(cudaFuncSetAttribute(k_sixpt_sample_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSixSolveSmem));
        while (count > 0) {
          const int cap = round == 0 ? first_cap : round_cap;
          SSFM_WCK(cudaEventRecord(evA, w.stream));
          dim3 gs(count, (cap + kSixSamplesPerBlock - 1) / kSixSamplesPerBlock);
          k_sixpt_sample_solve<<<gs, kSixSolveThreads, kSixSolveSmem, w.stream>>>(P, h->d_rays, h->offsets.p, q0, act, w.navail.p, w.six_it.p, R,
                                                        w.models.p, w.six_nm.p, w.pk_G.p, w.pk_id.p, w.pk_count.p, w.s32m.p,
                                                        w.six_M.p);
          SSFM_WCK(cudaGetLastError());
          SSFM_WCK(cudaEventRecord(evB, w.stream));
          dim3 gc(count, (cap * kSixMaxModels + 127) / 128);
          if (unit_z)
            k_sixpt_score<true><<<gc, 128, 0, w.stream>>>(h->uv4.p, nullptr, h->offsets.p, q0, act, R, w.pk_G.p, w.pk_id.p,
                                                          w.pk_count.p, thr32, w.s32m.p, w.counters.p);
          else
            k_sixpt_score<false><<<gc, 128, 0, w.stream>>>(h->u4.p, h->v4.p, h->offsets.p, q0, act, R, w.pk_G.p, w.pk_id.p,
                                                           w.pk_count.p, thr32, w.s32m.p, w.counters.p);
          SSFM_WCK(cudaGetLastError());
          SSFM_WCK(cudaEventRecord(evC, w.stream));
          SSFM_WCK(cudaMemsetAsync(w.counts.p + 1, 0, sizeof(int), w.stream));
          SixChainArgs A;
          A.rays = h->d_rays; A.offsets = h->offsets.p; A.pair0 = q0; A.list = act; A.nlist = count; A.navail = w.navail.p;
          A.states = w.six_states.p; A.R = R; A.models = w.models.p; A.nmodels = w.six_nm.p; A.s32m = w.s32m.p;
          A.flags = h->flags.p; A.results = h->results.p + q0; A.next_active = act_next; A.next_count = w.counts.p + 1;
          A.next_cap = round_cap; A.pk_count = w.pk_count.p; A.counters = w.counters.p; A.round_it = w.six_it.p;
          A.lo_states = w.six_lo_states.p; A.parked = w.parked0.p; A.parked_count = w.counts.p + 2;
          A.list_a = w.list_a.p; A.list_b = w.list_b.p; A.list_base = c0; A.mt = w.mt.p;
          if (!six_lo) {
            k_sixpt_chain<<<(count + kSixChainWarps - 1) / kSixChainWarps, kSixChainWarps * 32, 0, w.stream>>>(P, A);
            SSFM_WCK(cudaGetLastError());
            launches += 3;
          } else {
            // waves: walk -> (parked pairs) LocalOptimization -> walk the parked pairs on -> ... until nobody is parked
            int nlist = count;
            launches += 2;
            for (;;) {
              SSFM_WCK(cudaMemsetAsync(w.counts.p + 2, 0, sizeof(int), w.stream));
              A.nlist = nlist;
              k_sixpt_chain_lo<<<(nlist + kSixChainWarps - 1) / kSixChainWarps, kSixChainWarps * 32, 0, w.stream>>>(P, A);
              SSFM_WCK(cudaGetLastError());
              SSFM_WCK(cudaMemcpyAsync(w.h_count + 1, w.counts.p + 2, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
              SSFM_WCK(cudaStreamSynchronize(w.stream));
              const int nparked = w.h_count[1];
              launches += 1;
              if (nparked == 0) break;
              k_sixpt_lo<<<(nparked + kSixLoWarps - 1) / kSixLoWarps, kSixLoWarps * 32, 0, w.stream>>>(P, A, nparked);
              SSFM_WCK(cudaGetLastError());
              launches += 1;
              w.refit_waves += 1;
              // the next walk reads this wave's parked list while it appends the next one: hand it a copy
              SSFM_WCK(cudaMemcpyAsync(w.parked1.p, w.parked0.p, sizeof(int) * nparked, cudaMemcpyDeviceToDevice, w.stream));
              A.list = w.parked1.p;
              nlist = nparked;
            }
          }
          SSFM_WCK(cudaEventRecord(evD, w.stream));
          SSFM_WCK(cudaMemcpyAsync(w.h_count, w.counts.p + 1, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
          SSFM_WCK(cudaStreamSynchronize(w.stream));
          float t1 = 0, t2 = 0, t3 = 0;
          cudaEventElapsedTime(&t1, evA, evB);
          cudaEventElapsedTime(&t2, evB, evC);
          cudaEventElapsedTime(&t3, evC, evD);
          w.solve_ms += t1;
          w.score_ms += t2;
          w.chain_ms += t3;
          w.score_launches += 1;
          count = w.h_count[0];
          std::swap(act, act_next);
          ++round;
        }
        w.rounds = std::max(w.rounds, round);
      }
      continue;
    }
    if (P.driver == SSFM_DRIVER_PREEMPTIVE) {
      // No sequential dependence between hypotheses: generate all M per pair, then one CTA per pair runs the
      // block-wise elimination.  Two launches per pass.
      const int M = P.fixed_budget;
      int Mpad = 1;
      while (Mpad < M) Mpad <<= 1;
      SSFM_WCK(w.models.ensure((size_t)np * 6 * M));
      SSFM_WCK(w.has.ensure((size_t)np * M));
      SSFM_WCK(cudaEventRecord(evA, w.stream));
      SSFM_WCK(cudaMemsetAsync(w.has.p, 0, (size_t)np * M, w.stream));
      dim3 grid(np, (M + 63) / 64);
      if (P.solver == 0) k_preempt_hypotheses<0><<<grid, 64, 0, w.stream>>>(P, h->d_rays, h->offsets.p, pair0, M, w.models.p, w.has.p);
      else if (P.solver == 1) k_preempt_hypotheses<1><<<grid, 64, 0, w.stream>>>(P, h->d_rays, h->offsets.p, pair0, M, w.models.p, w.has.p);
      else k_preempt_hypotheses<2><<<grid, 64, 0, w.stream>>>(P, h->d_rays, h->offsets.p, pair0, M, w.models.p, w.has.p);
      SSFM_WCK(cudaGetLastError());
      SSFM_WCK(cudaEventRecord(evB, w.stream));
      const size_t smem = (size_t)Mpad * (sizeof(long long) + sizeof(int));
      SSFM_WCK(cudaFuncSetAttribute(k_preempt_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_preempt_select<<<np, kPreemptThreads, smem, w.stream>>>(P, h->d_rays, h->offsets.p, pair0, M, Mpad, P.preempt_block,
                                                                w.models.p, w.has.p, h->flags.p, h->results.p + pair0,
                                                                w.counters.p);
      SSFM_WCK(cudaGetLastError());
      SSFM_WCK(cudaEventRecord(evC, w.stream));
      SSFM_WCK(cudaStreamSynchronize(w.stream));
      float t1 = 0, t2 = 0;
      cudaEventElapsedTime(&t1, evA, evB);
      cudaEventElapsedTime(&t2, evB, evC);
      w.solve_ms += t1;
      w.chain_ms += t2;
      w.rounds += 1;
      launches += 2;
      continue;
    }
    SSFM_WCK(w.states.ensure(np));
    SSFM_WCK(w.active0.ensure(np));
    SSFM_WCK(w.active1.ensure(np));
    SSFM_WCK(w.ident.ensure(np));
    SSFM_WCK(w.navail.ensure(np));
    SSFM_WCK(w.models.ensure((size_t)np * 24 * R));
    SSFM_WCK(w.s32.ensure((size_t)np * R));
    SSFM_WCK(w.s32m.ensure((size_t)np * R * 4));
    SSFM_WCK(w.list_a.ensure(mpass + 16));
    SSFM_WCK(w.list_b.ensure(P.num_lo_steps > 0 ? mpass + 16 : 16));
    SSFM_WCK(w.mt.ensure(P.driver == SSFM_DRIVER_LO_MSAC ? (size_t)np * 625 : 625));
    SSFM_WCK(w.parked0.ensure(np));
    SSFM_WCK(w.parked1.ensure(np));
    SSFM_WCK(w.lm_E.ensure((size_t)np * 9));
    SSFM_WCK(w.lm_states.ensure(defer ? np : 1));
    SSFM_WCK(w.long_list.ensure(defer ? np : 1));

    k_init_pairs<<<(np + 127) / 128, 128, 0, w.stream>>>(P, h->offsets.p, pair0, np, w.states.p, w.mt.p, w.active0.p,
                                                         w.ident.p, w.navail.p, first_cap);
    SSFM_WCK(cudaMemsetAsync(w.counts.p, 0, 2 * sizeof(int), w.stream));
    k_finish_trivial<<<(np + 127) / 128, 128, 0, w.stream>>>(P, h->offsets.p, pair0, np, w.states.p, h->flags.p, 0,
                                                             h->results.p + pair0, w.active0.p, w.counts.p);
    launches += 2;
    SSFM_WCK(cudaMemcpyAsync(w.h_count, w.counts.p, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
    SSFM_WCK(cudaStreamSynchronize(w.stream));
    int count = w.h_count[0];
    int* act = w.active0.p;
    int* act_next = w.active1.p;
    int round = 0;
    while (count > 0) {
      const int cap = round == 0 ? first_cap : round_cap;
      SSFM_WCK(cudaEventRecord(evA, w.stream));
      auto launch_round = [&](const int* list, int n, bool uz, bool mark) -> cudaError_t {
        if (P.solver == 0) launch_solve<0>(h, w, P, pair0, list, n, cap, R);
        else if (P.solver == 1) launch_solve<1>(h, w, P, pair0, list, n, cap, R);
        else launch_solve<2>(h, w, P, pair0, list, n, cap, R);
        if (mark) cudaEventRecord(evB, w.stream);  // solve | score boundary for the stage timers
        dim3 grid(n, (cap + kScoreThreads - 1) / kScoreThreads);
        if (uz)
          k_score_rounds<true><<<grid, kScoreThreads, 0, w.stream>>>(h->uv4.p, nullptr, h->offsets.p, pair0, list, w.navail.p, R,
                                                                     w.models.p, thr32, w.s32.p, w.s32m.p);
        else
          k_score_rounds<false><<<grid, kScoreThreads, 0, w.stream>>>(h->u4.p, h->v4.p, h->offsets.p, pair0, list, w.navail.p,
                                                                      R, w.models.p, thr32, w.s32.p, w.s32m.p);
        return cudaGetLastError();
      };
      if (round == 0 && pd.pipelined) {
        // The upload is still streaming in: sample/solve/score each chunk of pairs as soon as its rays
        // are in HBM (identity list slice; pairs with nothing to do exit at once), so the H2D copy of
        // later chunks overlaps the first round's kernels.
        SSFM_WCK(cudaEventRecord(evB, w.stream));  // chunked round: everything is booked under 'score'
        bool all_unit = true;
        for (size_t k = 0; k + 1 < h->up_bounds.size(); ++k) {
          const int q0 = std::max(h->up_bounds[k], pair0), q1 = std::min(h->up_bounds[k + 1], pair0 + np);
          if (q1 <= q0) continue;
          SSFM_WCK(cudaEventSynchronize(h->up_ev[k]));
          const bool uz = h->h_up_flags[k] == 0 && !h->knobs.no_unitz;
          all_unit = all_unit && uz;
          if (!uz)
            if (int rc = ensure_general_planes(h, w.stream, h->h_offsets[q0], h->h_offsets[q1], &w.err)) return rc;
          SSFM_WCK(launch_round(w.ident.p + (q0 - pair0), q1 - q0, uz, false));
          launches += 2;
        }
        unit_z = all_unit;
        if (!unit_z)  // later rounds score the whole pass with the general kernel
          if (int rc = ensure_general_planes(h, w.stream, c0, c1, &w.err)) return rc;
        launches -= 2;
      } else {
        SSFM_WCK(launch_round(act, count, unit_z, true));
      }
      SSFM_WCK(cudaEventRecord(evC, w.stream));
      cudaStream_t hs = w.stream_hi;  // everything below runs at high priority, after the scoring kernel
      SSFM_WCK(cudaStreamWaitEvent(hs, evC, 0));
      SSFM_WCK(cudaMemsetAsync(w.counts.p + 1, 0, sizeof(int), hs));
      {
        ChainArgs A;
        A.rays = h->d_rays; A.offsets = h->offsets.p; A.pair0 = pair0;
        A.xy64 = (unit_z && !h->knobs.no_xy64) ? h->xy64.p : nullptr;
        A.navail = w.navail.p; A.states = w.states.p; A.R = R; A.models = w.models.p; A.s32 = w.s32.p; A.s32m = w.s32m.p;
        A.list_a = w.list_a.p; A.list_b = w.list_b.p; A.mt = w.mt.p; A.lm_E = w.lm_E.p; A.list_base = c0;
        A.flags = h->flags.p + c0; A.results = h->results.p + pair0; A.next_active = act_next; A.next_count = w.counts.p + 1;
        A.next_cap = round_cap; A.parked_small = w.counts.p + 4; A.parked_big = w.counts.p + 5; A.counters = w.counters.p;
        A.cap = np;
        if (!defer) {
          A.list = act; A.nlist = count; A.mode = 0; A.n_front = 0; A.parked = w.parked0.p;
          k_chain<false><<<(count + kChainWarps - 1) / kChainWarps, kChainWarps * 32, 0, hs>>>(P, A);
          SSFM_WCK(cudaGetLastError());
          launches += 1;
        } else {
          // waves: walk -> solve the parked refits -> resume, until no pair of this round is parked
          const int* cur = act;
          int ncur = count, mode = 0, nfront = 0;
          int* out = w.parked0.p;
          int* other = w.parked1.p;
          for (int wave = 0;; ++wave) {
            if (wave > R + 8) { w.err = "internal error: refit waves did not drain"; return SSFM_ERR_CUDA; }
            SSFM_WCK(cudaMemsetAsync(w.counts.p + 4, 0, 2 * sizeof(int), hs));
            A.list = cur; A.nlist = ncur; A.mode = mode; A.n_front = nfront; A.parked = out;
            k_chain<true><<<(ncur + kChainWarps - 1) / kChainWarps, kChainWarps * 32, 0, hs>>>(P, A);
            SSFM_WCK(cudaGetLastError());
            launches += 1;
            SSFM_WCK(cudaMemcpyAsync(w.h_count + 4, w.counts.p + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, hs));
            SSFM_WCK(cudaStreamSynchronize(hs));
            const int ns = w.h_count[4], nb = w.h_count[5];
            if (ns + nb == 0) break;
            if (ns > cfg.small_refit_threads_min) {
              // persistent lanes pulling from a queue: enough warps to fill the machine, not one per task
              SSFM_WCK(cudaMemsetAsync(w.counts.p + 6, 0, 2 * sizeof(int), hs));  // task queue head, straggler count
              const int blocks = std::min((ns + 63) / 64, h->num_sms * 8);
              SSFM_WCK(w.small_stage.ensure((size_t)h->num_sms * 8 * 64 * kSmallRefit * 6));
              const bool handover = cfg.handover;
              k_refit_small<<<blocks, 64, 0, hs>>>(P, h->d_rays, h->offsets.p, pair0, out, ns, w.counts.p + 6, w.states.p,
                                                   w.list_a.p, c0, w.lm_E.p, w.lm_states.p, handover ? w.long_list.p : nullptr,
                                                   w.counts.p + 7, cfg.handover_at,
                                                   h->knobs.no_small_stage ? nullptr : w.small_stage.p);
              launches += 1;
              if (handover) {
                SSFM_WCK(cudaMemsetAsync(w.counts.p + 6, 0, sizeof(int), hs));
                k_refit_long<<<h->num_sms * 2, 128, 0, hs>>>(P, h->d_rays, h->offsets.p, pair0, w.long_list.p, w.counts.p + 7,
                                                            w.counts.p + 6, w.states.p, w.list_a.p, c0, w.lm_E.p, w.lm_states.p);
                launches += 1;
              }
            } else if (ns > 0) {
              // too few to fill the machine one thread each: one warp per refit (lower latency; this is tail)
              k_refit_big<<<(ns + 3) / 4, 128, kRefitBigSmem, hs>>>(P, h->d_rays, h->offsets.p, pair0, out, np, ns, 1, w.states.p,
                                                        w.list_a.p, c0, w.lm_E.p);
              launches += 1;
            }
            if (nb > 0) {
              k_refit_big<<<(nb + 3) / 4, 128, kRefitBigSmem, hs>>>(P, h->d_rays, h->offsets.p, pair0, out, np, nb, 0, w.states.p,
                                                        w.list_a.p, c0, w.lm_E.p);
              launches += 1;
            }
            SSFM_WCK(cudaGetLastError());
            cur = out; ncur = ns + nb; mode = 1; nfront = ns;
            std::swap(out, other);
            w.refit_waves += 1;
          }
        }
      }
      SSFM_WCK(cudaEventRecord(evD, hs));
      SSFM_WCK(cudaMemcpyAsync(w.h_count, w.counts.p + 1, sizeof(int), cudaMemcpyDeviceToHost, hs));
      SSFM_WCK(cudaStreamSynchronize(hs));
      float t1 = 0, t2 = 0, t3 = 0;
      cudaEventElapsedTime(&t1, evA, evB);
      cudaEventElapsedTime(&t2, evB, evC);
      cudaEventElapsedTime(&t3, evC, evD);
      w.solve_ms += t1;
      w.score_ms += t2;
      w.chain_ms += t3;
      w.score_launches += 1;
      launches += 2;
      count = w.h_count[0];
      std::swap(act, act_next);
      ++round;
    }
    w.rounds += round;
  }
  SSFM_WCK(cudaMemcpyAsync(w.hc, w.counters.p, sizeof(w.hc), cudaMemcpyDeviceToHost, w.stream));
  SSFM_WCK(cudaStreamSynchronize(w.stream));
  w.launches = launches;
  return SSFM_OK;
}

}  // namespace

extern "C" {

int ssfm_abi_version(void) { return SSFM_ABI_VERSION; }

const char* ssfm_last_error(void) { return g_last_error.c_str(); }
void ssfm_internal_set_error(const char* msg) { g_last_error = msg ? msg : ""; }  // for the library's other translation units
int ssfm_internal_device(ssfm_handle h) { return h ? h->device : 0; }
cudaStream_t ssfm_internal_stream(ssfm_handle h) { return h ? h->stream : nullptr; }
int ssfm_internal_num_sms(ssfm_handle h) { return h ? h->num_sms : 0; }
void** ssfm_internal_ctx_slot(ssfm_handle h, void (*deleter)(void*)) {
  h->match_ctx_delete = deleter;
  return &h->match_ctx;
}

void ssfm_default_options(SsfmOptions* o) {
  if (!o) return;
  o->min_num_iterations = 100u;
  o->max_num_iterations = 10000u;
  o->success_probability = 0.9999;
  o->squared_inlier_threshold = 1.0;
  o->random_seed = 0u;
  o->num_lo_steps = 10;
  o->threshold_multiplier = 1.4142135623730951;
  o->num_lsq_iterations = 4;
  o->min_sample_multiplicator = 7;
  o->non_min_sample_multiplier = 3;
  o->lo_starting_iterations = 50u;
  o->final_least_squares = 0;
  o->solver = SSFM_SOLVER_ACTION_MATRIX;
  o->driver = SSFM_DRIVER_LO_MSAC;
  o->inward = 0;
  o->fixed_budget = 512;
  o->fixed_prob_success = 0.999;
  o->first_pair_id = 0u;
  o->min_num_points = 0;
  o->preemptive_block = 10; /* PreemptiveRANSAC( size_t _B = 10 ), preemptive_ransac.h:40 */
  o->sixpt_focal_scoring = 0;
  o->complex_root_models = SSFM_COMPLEX_CANONICAL;
}

static int create_resources(ssfm_engine* h);

int ssfm_create(int device, ssfm_handle* out) {
  if (!out) return fail(SSFM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SSFM_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                        " (this engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SSFM_ERR_INVALID, "device index out of range");
  SSFM_CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  SSFM_CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(SSFM_ERR_NO_DEVICE, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                        "; libssfm_b200 is built for sm_100a only");
  ssfm_engine* h = new ssfm_engine();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  const int rc = create_resources(h);
  if (rc != SSFM_OK) {
    ssfm_destroy(h);
    return rc;
  }
  *out = h;
  return SSFM_OK;
}

static int create_resources(ssfm_engine* h) {
  SSFM_CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    int prio_lo = 0, prio_hi = 0;
    SSFM_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SSFM_CK(cudaStreamCreateWithPriority(&h->stream_pack, cudaStreamNonBlocking, prio_hi));
  }
  for (auto& ev : h->ev) SSFM_CK(cudaEventCreate(&ev));
  SSFM_CK(cudaEventCreateWithFlags(&h->ev_tables, cudaEventDisableTiming));
  SSFM_CK(cudaMallocHost(&h->h_count, 64));
  SSFM_CK(cudaMallocHost(&h->h_up_flags, sizeof(int) * kMaxUploadChunks));
  h->num_workers = kMaxWorkers;
  h->workers = new Worker[kMaxWorkers];
  for (int k = 0; k < kMaxWorkers; ++k) {
    Worker& w = h->workers[k];
    int prio_lo = 0, prio_hi = 0;
    SSFM_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    SSFM_CK(cudaStreamCreateWithPriority(&w.stream, cudaStreamNonBlocking, prio_lo));
    SSFM_CK(cudaStreamCreateWithPriority(&w.stream_hi, cudaStreamNonBlocking, prio_hi));
    for (auto& ev : w.ev) SSFM_CK(cudaEventCreate(&ev));
    SSFM_CK(cudaMallocHost(&w.h_count, 64));
  }
  return SSFM_OK;
}

void ssfm_destroy(ssfm_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->match_ctx && h->match_ctx_delete) h->match_ctx_delete(h->match_ctx);
  h->match_ctx = nullptr;
  h->rays_own.release(); h->u4.release(); h->v4.release(); h->uv4.release(); h->xy64.release(); h->offsets.release();
  h->counts.release(); h->results.release(); h->flags.release();
  h->m_kp.release(); h->m_kpoff.release(); h->m_pairs.release(); h->m_matches.release(); h->m_kinv.release();
  for (int k = 0; k < h->num_workers; ++k) {
    Worker& w = h->workers[k];
    if (w.stream) cudaStreamSynchronize(w.stream);
    if (w.stream_hi) cudaStreamSynchronize(w.stream_hi);
    w.release();
    for (auto& ev : w.ev)
      if (ev) cudaEventDestroy(ev);
    if (w.h_count) cudaFreeHost(w.h_count);
    if (w.stream) cudaStreamDestroy(w.stream);
    if (w.stream_hi) cudaStreamDestroy(w.stream_hi);
  }
  delete[] h->workers;
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  if (h->ev_tables) cudaEventDestroy(h->ev_tables);
  if (h->h_count) cudaFreeHost(h->h_count);
  if (h->h_up_flags) cudaFreeHost(h->h_up_flags);
  for (auto& e : h->up_ev) cudaEventDestroy(e);
  for (auto& e : h->copy_ev) cudaEventDestroy(e);
  if (h->stream_pack) cudaStreamDestroy(h->stream_pack);
  h->up_flags.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// Chunked, asynchronous upload plan: graded chunk sizes (the first kernels start after a 2048-pair copy instead of a
// 16384-pair one), one event + one unit-z flag per chunk.
static int plan_upload_chunks(ssfm_handle h) {
  h->up_bounds.clear();
  int step = h->knobs.first_chunk;
  for (int p0 = 0; p0 < h->P;) {
    h->up_bounds.push_back(p0);
    p0 += std::min(step, kPipelinePassPairs);
    if (step < kPipelinePassPairs) step *= 2;
  }
  h->up_bounds.push_back(h->P);
  const int nchunks = (int)h->up_bounds.size() - 1;
  SSFM_CK(h->up_flags.ensure(nchunks));
  SSFM_CK(cudaMemsetAsync(h->up_flags.p, 0, sizeof(int) * nchunks, h->stream));
  while ((int)h->up_ev.size() < nchunks) {
    cudaEvent_t e;
    SSFM_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->up_ev.push_back(e);
  }
  while ((int)h->copy_ev.size() < nchunks) {
    cudaEvent_t e;
    SSFM_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->copy_ev.push_back(e);
  }
  return SSFM_OK;
}

static bool want_pipelined_upload(ssfm_handle h) {
  return h->P > 2 * kPipelinePassPairs && (h->P / kPipelinePassPairs) < kMaxUploadChunks - 1;
}

static int upload_impl(ssfm_handle h, const SsfmBatch* b, bool pipelined) {
  if (!h || !b) return fail(SSFM_ERR_INVALID, "NULL handle or batch");
  if (b->num_pairs < 0 || (b->num_pairs > 0 && (!b->offsets || !b->rays))) return fail(SSFM_ERR_INVALID, "bad batch");
  if (b->ray_format != SSFM_RAYS_F64 && b->ray_format != SSFM_RAYS_F32) return fail(SSFM_ERR_INVALID, "unknown ray_format");
  const bool f32 = b->ray_format == SSFM_RAYS_F32;
  if (b->rays_on_device && (reinterpret_cast<uintptr_t>(b->rays) & (f32 ? 7u : 15u)) != 0)
    return fail(SSFM_ERR_INVALID, "device rays must be 16-byte aligned (8-byte for SSFM_RAYS_F32): they are read as double2 / float2");
  SSFM_CK(cudaSetDevice(h->device));
  h->resident = false;
  h->have_results = false;
  h->P = b->num_pairs;
  h->h_offsets.assign(b->offsets, b->offsets + b->num_pairs + 1);
  if (b->num_pairs == 0) h->h_offsets.assign(1, 0);
  for (int p = 0; p < h->P; ++p)
    if (h->h_offsets[p + 1] < h->h_offsets[p] || h->h_offsets[p + 1] - h->h_offsets[p] > 0x7fffffffLL)
      return fail(SSFM_ERR_INVALID, "offsets must be non-decreasing (and each pair < 2^31 correspondences)");
  if (h->h_offsets[0] != 0) return fail(SSFM_ERR_INVALID, "offsets[0] must be 0");
  h->M = h->h_offsets[h->P];
  h->stats = SsfmRunStats();
  SSFM_CK(h->offsets.ensure(h->P + 1));
  SSFM_CK(cudaMemcpyAsync(h->offsets.p, h->h_offsets.data(), sizeof(long long) * (h->P + 1), cudaMemcpyHostToDevice, h->stream));
  const size_t m = (size_t)std::max<long long>(h->M, 1);
  SSFM_CK(h->uv4.ensure(m));
  SSFM_CK(h->xy64.ensure(4 * m));
  SSFM_CK(h->counts.ensure(8));
  SSFM_CK(cudaMemsetAsync(h->counts.p, 0, 8 * sizeof(int), h->stream));
  // Worker streams are non-blocking and start reading `offsets` before any chunk event in the pipelined plan:
  // they wait on this event first (run_range), so a reused engine never sees the previous batch's table.
  SSFM_CK(cudaEventRecord(h->ev_tables, h->stream));
  h->up_bounds.clear();
  if (pipelined && !b->rays_on_device && want_pipelined_upload(h)) {
    // Chunked, asynchronous upload: chunk k = pairs [k*16384, (k+1)*16384); H2D + pack on the copy
    // stream, an event per chunk.  ssfm_run's passes wait on the events, so the copy of chunk k+1
    // overlaps the kernels of chunk k.  (The caller's buffer is only read until ssfm_run returns.)
    SSFM_CK(h->rays_own.ensure(m * 6));
    if (f32) SSFM_CK(h->rays_f32.ensure(m * 6));
    h->d_rays = h->rays_own.p;
    if (int rc = plan_upload_chunks(h)) return rc;
    const int nchunks = (int)h->up_bounds.size() - 1;
    for (int k = 0; k < nchunks; ++k) {
      const long long c0 = h->h_offsets[h->up_bounds[k]], c1 = h->h_offsets[h->up_bounds[k + 1]];
      // the copies queue back to back on the copy stream; each chunk's pack kernel runs on the high-priority stream as
      // soon as the chunk has landed
      if (c1 > c0) {
        if (f32)
          SSFM_CK(cudaMemcpyAsync(h->rays_f32.p + 6 * c0, reinterpret_cast<const float*>(b->rays) + 6 * c0,
                                  sizeof(float) * 6 * (size_t)(c1 - c0), cudaMemcpyHostToDevice, h->stream));
        else
          SSFM_CK(cudaMemcpyAsync(h->rays_own.p + 6 * c0, b->rays + 6 * c0, sizeof(double) * 6 * (size_t)(c1 - c0),
                                  cudaMemcpyHostToDevice, h->stream));
      }
      SSFM_CK(cudaEventRecord(h->copy_ev[k], h->stream));
      SSFM_CK(cudaStreamWaitEvent(h->stream_pack, h->copy_ev[k], 0));
      if (c1 > c0) {
        const unsigned nb = (unsigned)((c1 - c0 + 255) / 256);
        if (f32)
          k_pack_f32<<<nb, 256, 0, h->stream_pack>>>(h->rays_f32.p + 6 * c0, c1 - c0, h->rays_own.p + 6 * c0, h->uv4.p + c0,
                                                     h->xy64.p + 4 * c0, h->up_flags.p + k);
        else
          k_pack<<<nb, 256, 0, h->stream_pack>>>(h->rays_own.p + 6 * c0, c1 - c0, h->uv4.p + c0, h->xy64.p + 4 * c0,
                                                 h->up_flags.p + k);
        SSFM_CK(cudaGetLastError());
      }
      SSFM_CK(cudaMemcpyAsync(h->h_up_flags + k, h->up_flags.p + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream_pack));
      SSFM_CK(cudaEventRecord(h->up_ev[k], h->stream_pack));
    }
    if (nchunks > 0) SSFM_CK(cudaStreamWaitEvent(h->stream, h->up_ev[nchunks - 1], 0));  // a sync of `stream` covers the packs
    h->stats.h2d_bytes = (long long)((f32 ? sizeof(float) : sizeof(double)) * 6 * (size_t)h->M + sizeof(long long) * (h->P + 1));
    h->unit_z = false;  // decided per chunk
    h->resident = true;
    return SSFM_OK;
  }
  SSFM_CK(cudaEventRecord(h->ev[4], h->stream));
  const float* src32 = nullptr;
  if (f32) {  // the float64 records the exact passes read are always the engine's own copy
    SSFM_CK(h->rays_own.ensure(m * 6));
    if (b->rays_on_device) {
      src32 = reinterpret_cast<const float*>(b->rays);
    } else {
      SSFM_CK(h->rays_f32.ensure(m * 6));
      if (h->M > 0)
        SSFM_CK(cudaMemcpyAsync(h->rays_f32.p, b->rays, sizeof(float) * 6 * (size_t)h->M, cudaMemcpyHostToDevice, h->stream));
      src32 = h->rays_f32.p;
      h->stats.h2d_bytes = (long long)(sizeof(float) * 6 * (size_t)h->M + sizeof(long long) * (h->P + 1));
    }
    h->d_rays = h->rays_own.p;
  } else if (b->rays_on_device) {
    h->d_rays = b->rays;
  } else {
    SSFM_CK(h->rays_own.ensure(m * 6));
    if (h->M > 0)
      SSFM_CK(cudaMemcpyAsync(h->rays_own.p, b->rays, sizeof(double) * 6 * (size_t)h->M, cudaMemcpyHostToDevice, h->stream));
    h->d_rays = h->rays_own.p;
    h->stats.h2d_bytes = (long long)(sizeof(double) * 6 * (size_t)h->M + sizeof(long long) * (h->P + 1));
  }
  if (h->M > 0) {
    const int threads = 256;
    const long long blocks = (h->M + threads - 1) / threads;
    if (f32)
      k_pack_f32<<<(unsigned)blocks, threads, 0, h->stream>>>(src32, h->M, h->rays_own.p, h->uv4.p, h->xy64.p, h->counts.p + 2);
    else
      k_pack<<<(unsigned)blocks, threads, 0, h->stream>>>(h->d_rays, h->M, h->uv4.p, h->xy64.p, h->counts.p + 2);
    SSFM_CK(cudaGetLastError());
  }
  SSFM_CK(cudaEventRecord(h->ev[5], h->stream));
  SSFM_CK(cudaMemcpyAsync(h->h_count + 2, h->counts.p + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));  // the caller's buffer may go away after we return
  h->unit_z = h->h_count[2] == 0 && !h->knobs.no_unitz;
  if (!h->unit_z) {
    std::string err;
    if (int rc = ensure_general_planes(h, h->stream, 0, h->M, &err)) return fail(rc, err);
    SSFM_CK(cudaStreamSynchronize(h->stream));
  }
  float ms = 0.f;
  SSFM_CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]));
  h->stats.pack_ms = ms;
  h->resident = true;
  return SSFM_OK;
}

int ssfm_upload(ssfm_handle h, const SsfmBatch* b) {
  refresh_knobs(h);
  return upload_impl(h, b, false);
}

static int upload_matches_impl(ssfm_handle h, const SsfmMatchBatch* b, bool pipelined) {
  if (!h || !b) return fail(SSFM_ERR_INVALID, "NULL handle or batch");
  if (b->num_pairs < 0 || b->num_images < 0) return fail(SSFM_ERR_INVALID, "bad batch");
  if (b->num_pairs > 0 && (!b->keypoint_offsets || !b->keypoints_xy || !b->pair_images || !b->match_offsets || !b->matches))
    return fail(SSFM_ERR_INVALID, "NULL array in match batch");
  SSFM_CK(cudaSetDevice(h->device));
  const int P = b->num_pairs;
  const long long M = P > 0 ? b->match_offsets[P] : 0;
  const long long NK = b->num_images > 0 ? b->keypoint_offsets[b->num_images] : 0;
  if (P > 0 && b->match_offsets[0] != 0) return fail(SSFM_ERR_INVALID, "match_offsets[0] must be 0");
  for (int p = 0; p < P; ++p) {
    const int i0 = b->pair_images[2 * p], i1 = b->pair_images[2 * p + 1];
    if (i0 < 0 || i1 < 0 || i0 >= b->num_images || i1 >= b->num_images) return fail(SSFM_ERR_INVALID, "pair image index out of range");
    if (b->match_offsets[p + 1] < b->match_offsets[p] || b->match_offsets[p + 1] - b->match_offsets[p] > 0x7fffffffLL)
      return fail(SSFM_ERR_INVALID, "match_offsets must be non-decreasing (and each pair < 2^31 matches)");
  }  // (keypoint indices of the individual matches are range-checked by the kernel)
  h->resident = false;
  h->have_results = false;
  h->matches_pending_check = false;
  h->P = P;
  h->M = M;
  if (P > 0) h->h_offsets.assign(b->match_offsets, b->match_offsets + P + 1);
  else h->h_offsets.assign(1, 0);
  h->stats = SsfmRunStats();
  h->up_bounds.clear();
  // persistent (grow-only) device buffers: no allocation after the first call of a given size
  SSFM_CK(h->m_kp.ensure((size_t)std::max<long long>(2 * NK, 2)));
  SSFM_CK(h->m_kpoff.ensure(b->num_images + 1));
  SSFM_CK(h->m_pairs.ensure((size_t)std::max(2 * P, 2)));
  SSFM_CK(h->m_matches.ensure((size_t)std::max<long long>(2 * M, 2)));
  SSFM_CK(h->m_kinv.ensure(9));
  SSFM_CK(h->rays_own.ensure((size_t)std::max<long long>(6 * M, 6)));
  SSFM_CK(h->uv4.ensure((size_t)std::max<long long>(M, 1)));
  SSFM_CK(h->xy64.ensure((size_t)std::max<long long>(4 * M, 4)));
  SSFM_CK(h->offsets.ensure(P + 1));
  SSFM_CK(h->counts.ensure(8));
  h->d_rays = h->rays_own.p;
  cudaStream_t st = h->stream;
  SSFM_CK(cudaMemsetAsync(h->counts.p, 0, 8 * sizeof(int), st));
  const bool pipe = pipelined && want_pipelined_upload(h);
  if (NK > 0 && !pipe) SSFM_CK(cudaMemcpyAsync(h->m_kp.p, b->keypoints_xy, sizeof(float) * 2 * (size_t)NK, cudaMemcpyHostToDevice, st));
  if (b->num_images > 0)
    SSFM_CK(cudaMemcpyAsync(h->m_kpoff.p, b->keypoint_offsets, sizeof(long long) * (b->num_images + 1), cudaMemcpyHostToDevice, st));
  SSFM_CK(cudaMemcpyAsync(h->offsets.p, h->h_offsets.data(), sizeof(long long) * (P + 1), cudaMemcpyHostToDevice, st));
  if (P > 0) SSFM_CK(cudaMemcpyAsync(h->m_pairs.p, b->pair_images, sizeof(int) * 2 * (size_t)P, cudaMemcpyHostToDevice, st));
  SSFM_CK(cudaMemcpyAsync(h->m_kinv.p, b->Kinv, sizeof(double) * 9, cudaMemcpyHostToDevice, st));
  SSFM_CK(cudaEventRecord(h->ev_tables, st));
  h->stats.h2d_bytes = (long long)(8 * NK + 8 * (long long)P + 8 * M + 8 * (P + 1) + 8 * (b->num_images + 1) + 72);
  auto build = [&](long long c0, long long c1, int* flag, cudaStream_t kst, cudaEvent_t landed) -> cudaError_t {  // matches [c0, c1): H2D, rays + FP32 plane in one kernel
    if (c1 <= c0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(h->m_matches.p + 2 * c0, b->matches + 2 * c0, sizeof(int) * 2 * (size_t)(c1 - c0), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if (kst != st) {  // pipelined: the kernel runs on the high-priority stream once the chunk has landed
      if ((e = cudaEventRecord(landed, st)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(kst, landed, 0)) != cudaSuccess) return e;
    }
    k_build_rays<<<(unsigned)((c1 - c0 + 255) / 256), 256, 0, kst>>>(reinterpret_cast<const float2*>(h->m_kp.p), h->m_kpoff.p, h->m_pairs.p,
                                                                    h->offsets.p, P, reinterpret_cast<const int2*>(h->m_matches.p), c0,
                                                                    c1 - c0, h->m_kinv.p, h->rays_own.p, h->uv4.p, h->xy64.p, flag, h->counts.p + 3);
    return cudaGetLastError();
  };
  if (pipe) {
    // chunk k = a range of pairs: its matches are copied and turned into rays while earlier chunks already run their
    // first look-ahead round (same plan as the ray upload; ssfm_run waits on the chunk events).  Keypoints travel
    // just in time: before a chunk's matches, the images it references that are not in HBM yet (image tables sorted
    // by first use copy progressively; an exhaustive pair list references every image at once and copies them all first).
    if (int rc = plan_upload_chunks(h)) return rc;
    const int nchunks = (int)h->up_bounds.size() - 1;
    int images_copied = 0;
    for (int k = 0; k < nchunks; ++k) {
      int need = images_copied;
      for (int p = h->up_bounds[k]; p < h->up_bounds[k + 1]; ++p)
        need = std::max(need, std::max(b->pair_images[2 * p], b->pair_images[2 * p + 1]) + 1);
      if (need > images_copied) {
        const long long k0 = b->keypoint_offsets[images_copied], k1 = b->keypoint_offsets[need];
        if (k1 > k0)
          SSFM_CK(cudaMemcpyAsync(h->m_kp.p + 2 * k0, b->keypoints_xy + 2 * k0, sizeof(float) * 2 * (size_t)(k1 - k0), cudaMemcpyHostToDevice, st));
        images_copied = need;
      }
      const long long c0 = h->h_offsets[h->up_bounds[k]], c1 = h->h_offsets[h->up_bounds[k + 1]];
      if (c1 <= c0) {  // empty chunk: still order the flag read after whatever precedes it
        SSFM_CK(cudaEventRecord(h->copy_ev[k], st));
        SSFM_CK(cudaStreamWaitEvent(h->stream_pack, h->copy_ev[k], 0));
      }
      SSFM_CK(build(c0, c1, h->up_flags.p + k, h->stream_pack, h->copy_ev[k]));
      SSFM_CK(cudaMemcpyAsync(h->h_up_flags + k, h->up_flags.p + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream_pack));
      SSFM_CK(cudaEventRecord(h->up_ev[k], h->stream_pack));
    }
    if (nchunks > 0) SSFM_CK(cudaStreamWaitEvent(st, h->up_ev[nchunks - 1], 0));  // a sync of `stream` covers the kernels
    h->unit_z = false;  // decided per chunk
    h->matches_pending_check = true;
    h->resident = true;
    return SSFM_OK;
  }
  SSFM_CK(cudaEventRecord(h->ev[4], st));
  SSFM_CK(build(0, M, h->counts.p + 2, st, nullptr));
  SSFM_CK(cudaEventRecord(h->ev[5], st));
  SSFM_CK(cudaMemcpyAsync(h->h_count + 2, h->counts.p + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  SSFM_CK(cudaStreamSynchronize(st));  // the caller's buffers may go away after we return
  if (h->h_count[3] != 0) return fail(SSFM_ERR_INVALID, "match keypoint index out of range");
  h->unit_z = h->h_count[2] == 0 && !h->knobs.no_unitz;
  if (!h->unit_z) {
    std::string err;
    if (int rc = ensure_general_planes(h, st, 0, M, &err)) return fail(rc, err);
    SSFM_CK(cudaStreamSynchronize(st));
  }
  float ms = 0.f;
  SSFM_CK(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]));
  h->stats.pack_ms = ms;
  h->resident = true;
  return SSFM_OK;
}

int ssfm_upload_matches(ssfm_handle h, const SsfmMatchBatch* b) {
  refresh_knobs(h);
  return upload_matches_impl(h, b, false);
}

int ssfm_estimate_pairs_from_matches(ssfm_handle h, const SsfmMatchBatch* batch, const SsfmOptions* opt,
                                     SsfmPairResult* results, uint8_t* inlier_flags) {
  refresh_knobs(h);
  if (!results) return fail(SSFM_ERR_INVALID, "results is NULL");
  if (int rc = check_options(opt)) return rc;
  auto abandon = [h](int rc) {  // never leave copies from the caller's buffers in flight, nor a half-uploaded batch resident
    cudaStreamSynchronize(h->stream);
    h->up_bounds.clear();
    h->resident = false;
    h->have_results = false;
    h->matches_pending_check = false;
    return rc;
  };
  if (int rc = upload_matches_impl(h, batch, !h->knobs.no_pipeline)) return abandon(rc);
  if (int rc = ssfm_run(h, opt)) return abandon(rc);
  SSFM_CK(cudaStreamSynchronize(h->stream));
  if (h->matches_pending_check) {  // the index-range flag of a pipelined upload is complete only now
    SSFM_CK(cudaMemcpy(h->h_count + 3, h->counts.p + 3, sizeof(int), cudaMemcpyDeviceToHost));
    h->matches_pending_check = false;
    if (h->h_count[3] != 0) return abandon(fail(SSFM_ERR_INVALID, "match keypoint index out of range"));
  }
  if (!h->up_bounds.empty()) {  // the batch is fully resident now; later ssfm_run calls use the plain plan
    bool all_unit = true;
    for (size_t k = 0; k + 1 < h->up_bounds.size(); ++k) all_unit = all_unit && h->h_up_flags[k] == 0;
    h->unit_z = all_unit && !h->knobs.no_unitz;
    h->up_bounds.clear();
  }
  return ssfm_download(h, results, inlier_flags);
}

int ssfm_run(ssfm_handle h, const SsfmOptions* opt) {
  refresh_knobs(h);
  if (!h) return fail(SSFM_ERR_INVALID, "NULL handle");
  if (int rc = check_options(opt)) return rc;
  if (!h->resident) return fail(SSFM_ERR_INVALID, "ssfm_run without a resident batch (call ssfm_upload first)");
  SSFM_CK(cudaSetDevice(h->device));
  Params P = make_params(*opt, h->knobs);
  const double pack_ms = h->stats.pack_ms;
  const long long h2d = h->stats.h2d_bytes;
  h->stats = SsfmRunStats();
  h->stats.pack_ms = pack_ms;
  h->stats.h2d_bytes = h2d;
  h->have_results = false;
  SSFM_CK(h->results.ensure(std::max(h->P, 1)));
  SSFM_CK(h->flags.ensure((size_t)std::max<long long>(h->M, 1)));

  RunCfg cfg;
  if (P.driver == SSFM_DRIVER_MSAC_FIXED) {
    cfg.first_cap = 32;
    cfg.round_cap = 64;
  } else {
    cfg.first_cap = (int)std::min<uint32_t>((std::max<uint32_t>(P.min_iters, 32u) + 31u) & ~31u, 1024u);
    cfg.round_cap = kRoundCap;
  }
  if (h->knobs.round_cap > 0) cfg.round_cap = (std::max(32, h->knobs.round_cap) + 31) & ~31;
  if (h->knobs.first_cap > 0) cfg.first_cap = (std::max(32, h->knobs.first_cap) + 31) & ~31;
  cfg.R = std::max(cfg.first_cap, cfg.round_cap);
  cfg.defer = P.driver == SSFM_DRIVER_LO_MSAC && P.num_lo_steps <= 0 && !h->knobs.no_defer;
  cfg.handover = !h->knobs.no_handover;
  cfg.handover_at = kHandover;
  if (h->knobs.handover_at > 0) cfg.handover_at = h->knobs.handover_at;
  if (cfg.defer) {
    // Parking refits pays when a wave holds thousands of them; below that a wave is a host round trip for a handful of
    // threads (one C1-sized pair: five waves = 1.0 of its 1.27 ms).  Small batches run the same refits inline, with the
    // deferred path's arithmetic (Params::inline_small_max), so the table does not depend on the batch size.
    int defer_min = kDeferMinPairs;
    if (h->knobs.defer_min_pairs >= 0) defer_min = h->knobs.defer_min_pairs;
    if (h->P < defer_min) {
      cfg.defer = false;
      P.inline_small_max = kSmallRefit;
      P.inline_handover = cfg.handover ? cfg.handover_at : 0;
    }
  }
  cfg.thr32 = (float)P.thr2;
  // 0: every small refit is solved one thread per problem.  (A warp-per-refit path for thin waves has lower
  // latency but sums in a different order, which would make results depend on how many pairs share a wave.)
  cfg.small_refit_threads_min = 0;
  cfg.small_refit_threads_min = h->knobs.refit_threads_min;

  // Pass lists per worker: contiguous ranges balanced by correspondence count.
  // Resident batch: one stream (a second one gains ~1 %).  While the upload is still streaming in
  // (ssfm_estimate_pairs), three streams over six interleaved parts (measured: 456 -> 440 ms end to end).
  int nw = (!h->up_bounds.empty() && h->P >= 4096) ? 3 : 1;
  if (h->knobs.workers > 0) nw = std::min(kMaxWorkers, h->knobs.workers);
  if (!h->up_bounds.empty())
    if (h->knobs.e2e_workers > 0) nw = std::min(kMaxWorkers, h->knobs.e2e_workers);
  nw = std::max(1, std::min(nw, std::max(h->P, 1)));
  std::vector<std::vector<PassDesc>> plan(nw);
  {
    const int pipelined = h->up_bounds.empty() ? 0 : 1;
    // Resident batch: one contiguous range per worker.  Upload in flight: the pair list is cut into more parts
    // than workers, dealt round-robin in upload order, so that every worker owns pairs that arrive early and the
    // later rounds of one part overlap the arrival (and first round) of the next.
    int nparts = nw;
    if (pipelined && nw > 1) {
      nparts = 3 * nw;
      if (h->knobs.e2e_parts > 0) nparts = std::max(nw, h->knobs.e2e_parts);
    }
    std::vector<int> bounds(nparts + 1, 0);
    bounds[nparts] = h->P;
    for (int k = 1; k < nparts; ++k) {
      const long long target = h->M * k / nparts;
      int b = (int)(std::lower_bound(h->h_offsets.begin(), h->h_offsets.end(), target) - h->h_offsets.begin());
      bounds[k] = std::max(bounds[k - 1], std::min(b, h->P));
    }
    for (int k = 0; k < nparts; ++k)
      for (int p0 = bounds[k]; p0 < bounds[k + 1]; p0 += kMaxPassPairs)
        plan[k % nw].push_back(PassDesc{p0, std::min(kMaxPassPairs, bounds[k + 1] - p0), pipelined});
    if (!pipelined) SSFM_CK(cudaStreamSynchronize(h->stream));
  }
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> threads;
  for (int k = 1; k < nw; ++k) {
    Worker* w = &h->workers[k];
    const std::vector<PassDesc>* pl = &plan[k];
    threads.emplace_back([h, w, &P, &cfg, pl]() { w->rc = run_range(h, *w, P, cfg, *pl); });
  }
  h->workers[0].rc = run_range(h, h->workers[0], P, cfg, plan[0]);
  for (auto& t : threads) t.join();
  const auto t1 = std::chrono::steady_clock::now();
  unsigned long long hc[32] = {};
  for (int k = 0; k < nw; ++k) {
    Worker& w = h->workers[k];
    if (w.rc != SSFM_OK) return fail(w.rc, w.err);
    h->stats.solve_ms += w.solve_ms;
    h->stats.score_ms += w.score_ms;
    h->stats.chain_ms += w.chain_ms;
    h->stats.rounds = std::max(h->stats.rounds, w.rounds);
    h->stats.kernel_launches += w.launches;
    h->stats.score_launches += w.score_launches;
    h->stats.refit_waves += w.refit_waves;
    for (int i = 0; i < 32; ++i) hc[i] += w.hc[i];
  }
#if defined(SSFM_PROFILE_CHAIN)
  {
    const char* names[] = {"scan", "rescore", "lo_collect", "lo_shuffle", "lo_lm", "lo_score", "final_lm", "final_rest", "total"};
    fprintf(stderr, "[chain profile, warp-cycles]");
    for (int k = 0; k < 9; ++k) fprintf(stderr, " %s=%.3e", names[k], (double)hc[8 + k]);
    fprintf(stderr, "\n");
  }
#endif
  h->stats.total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  h->stats.workers = nw;
  h->stats.evals_executed = (long long)hc[0];
  h->stats.evals_exact = (long long)hc[1];
  h->have_results = true;
  return SSFM_OK;
}

int ssfm_download(ssfm_handle h, SsfmPairResult* results, uint8_t* inlier_flags) {
  if (!h) return fail(SSFM_ERR_INVALID, "NULL handle");
  if (!h->have_results) return fail(SSFM_ERR_INVALID, "no results to download (call ssfm_run first)");
  SSFM_CK(cudaSetDevice(h->device));
  long long bytes = 0;
  if (results && h->P > 0) {
    SSFM_CK(cudaMemcpyAsync(results, h->results.p, sizeof(SsfmPairResult) * (size_t)h->P, cudaMemcpyDeviceToHost, h->stream));
    bytes += (long long)sizeof(SsfmPairResult) * h->P;
  }
  if (inlier_flags && h->M > 0) {
    SSFM_CK(cudaMemcpyAsync(inlier_flags, h->flags.p, (size_t)h->M, cudaMemcpyDeviceToHost, h->stream));
    bytes += h->M;
  }
  SSFM_CK(cudaStreamSynchronize(h->stream));
  h->stats.d2h_bytes = bytes;
  if (results) {
    long long useful = 0;
    for (int p = 0; p < h->P; ++p) useful += results[p].evals;
    h->stats.evals_useful = useful;
  }
  return SSFM_OK;
}

int ssfm_get_stats(ssfm_handle h, SsfmRunStats* s) {
  if (!h || !s) return fail(SSFM_ERR_INVALID, "NULL argument");
  *s = h->stats;
  return SSFM_OK;
}

int ssfm_device_results(ssfm_handle h, void** dev_ptr, int32_t* num_pairs) {
  if (!h || !dev_ptr) return fail(SSFM_ERR_INVALID, "NULL argument");
  if (!h->have_results) return fail(SSFM_ERR_INVALID, "no results (call ssfm_run first)");
  *dev_ptr = h->results.p;
  if (num_pairs) *num_pairs = h->P;
  return SSFM_OK;
}

int ssfm_estimate_pairs(ssfm_handle h, const SsfmBatch* batch, const SsfmOptions* opt, SsfmPairResult* results,
                        uint8_t* inlier_flags) {
  refresh_knobs(h);
  if (!results) return fail(SSFM_ERR_INVALID, "results is NULL");
  if (int rc = check_options(opt)) return rc;
  auto abandon = [h](int rc) {  // never leave copies from the caller's buffer in flight, nor a half-uploaded batch resident
    cudaStreamSynchronize(h->stream);
    h->up_bounds.clear();
    h->resident = false;
    h->have_results = false;
    return rc;
  };
  if (int rc = upload_impl(h, batch, !h->knobs.no_pipeline)) return abandon(rc);
  if (int rc = ssfm_run(h, opt)) return abandon(rc);
  SSFM_CK(cudaStreamSynchronize(h->stream));
  if (!h->up_bounds.empty()) {  // the batch is fully resident now; later ssfm_run calls use the plain plan
    bool all_unit = true;
    for (size_t k = 0; k + 1 < h->up_bounds.size(); ++k) all_unit = all_unit && h->h_up_flags[k] == 0;
    h->unit_z = all_unit && !h->knobs.no_unitz;
    h->up_bounds.clear();
  }
  if (int rc = ssfm_download(h, results, inlier_flags)) return rc;
  if (batch->rays_on_device) { h->resident = false; h->d_rays = nullptr; }  // never keep a caller pointer
  return SSFM_OK;
}

// ----------------------------------------------------------------------------------------------
// replay hooks
// ----------------------------------------------------------------------------------------------
int ssfm_sample(uint32_t seed, uint32_t pair, uint32_t iter, int32_t k, int32_t n, int32_t* idx) {
  if (!idx || k <= 0 || k > 8 || n < k) return fail(SSFM_ERR_INVALID, "ssfm_sample: need 0 < k <= 8 and n >= k");
  philox_sample<8>(seed, pair, iter, k, n, idx);
  return SSFM_OK;
}

int ssfm_retriangulate(ssfm_handle h, const SsfmTrackBatch* b, const SsfmOptions* opt, double* points_xyz, int32_t* num_inliers,
                       int32_t* status, uint32_t* num_iterations) {
  if (!h || !b || !opt || !points_xyz || !num_inliers || !status) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: NULL argument");
  if (int rc = check_options(opt)) return rc;
  if (b->num_points < 0 || b->num_cameras < 0 || !(b->focal > 0.0)) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: bad sizes or focal");
  if (b->num_points == 0) return SSFM_OK;
  if (!b->obs_offsets || !b->camera_tr) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: NULL table");
  const long long M = b->obs_offsets[b->num_points];
  for (int p = 0; p < b->num_points; ++p)
    if (b->obs_offsets[p + 1] < b->obs_offsets[p]) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: offsets not monotone");
  if (b->obs_offsets[0] != 0 || (M > 0 && (!b->obs_camera || !b->obs_xy))) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: bad observation tables");
  for (long long i = 0; i < M; ++i)
    if (b->obs_camera[i] < 0 || b->obs_camera[i] >= b->num_cameras) return fail(SSFM_ERR_INVALID, "ssfm_retriangulate: camera index out of range");
  SSFM_CK(cudaSetDevice(h->device));
  Params P = make_params(*opt, h->knobs);
  DevBuf<tri::Cam> cams;
  DevBuf<double> d_tr, d_xy, d_pts;
  DevBuf<long long> d_off;
  DevBuf<int> d_cam, d_scratch, d_ninl, d_status, d_order, d_list0, d_list1, d_cnt;
  DevBuf<tri::LoState> d_states;
  DevBuf<unsigned int> d_iters;
  DevBuf<uint32_t> d_mt;
  const int NP = b->num_points;
  const int kPass = 262144;
  auto release = [&]() {
    cams.release(); d_tr.release(); d_xy.release(); d_pts.release(); d_off.release(); d_cam.release(); d_scratch.release();
    d_ninl.release(); d_status.release(); d_iters.release(); d_mt.release(); d_order.release();
    d_list0.release(); d_list1.release(); d_cnt.release(); d_states.release();
  };
#define SSFM_RT(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      release();                                                                                        \
      return fail(e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                                   \
  } while (0)
  SSFM_RT(cams.ensure(std::max(b->num_cameras, 1)));
  SSFM_RT(d_tr.ensure((size_t)std::max(b->num_cameras, 1) * 6));
  SSFM_RT(d_xy.ensure((size_t)std::max<long long>(M, 1) * 2));
  SSFM_RT(d_cam.ensure((size_t)std::max<long long>(M, 1)));
  SSFM_RT(d_off.ensure((size_t)NP + 1));
  SSFM_RT(d_pts.ensure((size_t)NP * 3));
  SSFM_RT(d_ninl.ensure(NP));
  SSFM_RT(d_status.ensure(NP));
  SSFM_RT(d_iters.ensure(NP));
  SSFM_RT(d_mt.ensure((size_t)std::min(NP, kPass) * 625));
  cudaStream_t st = h->stream;
  if (b->num_cameras > 0) SSFM_RT(cudaMemcpyAsync(d_tr.p, b->camera_tr, sizeof(double) * 6 * b->num_cameras, cudaMemcpyHostToDevice, st));
  if (M > 0) {
    SSFM_RT(cudaMemcpyAsync(d_xy.p, b->obs_xy, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, st));
    SSFM_RT(cudaMemcpyAsync(d_cam.p, b->obs_camera, sizeof(int) * M, cudaMemcpyHostToDevice, st));
  }
  SSFM_RT(cudaMemcpyAsync(d_off.p, b->obs_offsets, sizeof(long long) * ((size_t)NP + 1), cudaMemcpyHostToDevice, st));
  if (b->num_cameras > 0) k_tri_cameras<<<(b->num_cameras + 127) / 128, 128, 0, st>>>(d_tr.p, b->num_cameras, cams.p);
  // process the points in order of track length (stable counting sort on the host)
  std::vector<int> order(NP);
  {
    std::vector<std::pair<int, int>> key(NP);
    for (int p = 0; p < NP; ++p) key[p] = std::make_pair((int)(b->obs_offsets[p + 1] - b->obs_offsets[p]), p);
    std::stable_sort(key.begin(), key.end());
    for (int p = 0; p < NP; ++p) order[p] = key[p].second;
  }
  SSFM_RT(d_order.ensure(NP));
  SSFM_RT(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int) * NP, cudaMemcpyHostToDevice, st));
  SSFM_RT(d_scratch.ensure((size_t)std::max<long long>(4 * M, 1)));
  SSFM_RT(d_list0.ensure(std::min(NP, kPass)));
  SSFM_RT(d_list1.ensure(std::min(NP, kPass)));
  SSFM_RT(d_states.ensure(std::min(NP, kPass)));
  SSFM_RT(d_cnt.ensure(1));
  for (int p0 = 0; p0 < NP; p0 += kPass) {
    const int np = std::min(kPass, NP - p0);
    // phases: every point runs up to stop_at iterations; the unfinished ones are compacted and continued together
    // (iteration counts range from min_num_iterations to max_num_iterations: one launch would leave most lanes idle)
    int count = np;
    int* cur = nullptr;
    int* nxt = d_list0.p;
    unsigned int stop = 256;
    for (int phase = 0; count > 0; ++phase) {
      SSFM_RT(cudaMemsetAsync(d_cnt.p, 0, sizeof(int), st));
      k_retriangulate<<<(count + 63) / 64, 64, 0, st>>>(P, cams.p, d_off.p, d_cam.p, d_xy.p, b->focal, p0, count, d_order.p, cur, nxt,
                                                        d_cnt.p, stop, d_states.p, d_scratch.p, d_mt.p, d_pts.p, d_ninl.p,
                                                        d_status.p, d_iters.p);
      SSFM_RT(cudaGetLastError());
      int left = 0;
      SSFM_RT(cudaMemcpyAsync(&left, d_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      SSFM_RT(cudaStreamSynchronize(st));
      count = left;
      cur = nxt;
      nxt = (nxt == d_list0.p) ? d_list1.p : d_list0.p;
      stop = stop >= 0x40000000u ? 0xFFFFFFFFu : stop * 4;
    }
  }
  SSFM_RT(cudaMemcpyAsync(points_xyz, d_pts.p, sizeof(double) * 3 * NP, cudaMemcpyDeviceToHost, st));
  SSFM_RT(cudaMemcpyAsync(num_inliers, d_ninl.p, sizeof(int) * NP, cudaMemcpyDeviceToHost, st));
  SSFM_RT(cudaMemcpyAsync(status, d_status.p, sizeof(int) * NP, cudaMemcpyDeviceToHost, st));
  if (num_iterations) SSFM_RT(cudaMemcpyAsync(num_iterations, d_iters.p, sizeof(unsigned int) * NP, cudaMemcpyDeviceToHost, st));
  SSFM_RT(cudaStreamSynchronize(st));
#undef SSFM_RT
  release();
  return SSFM_OK;
}

int ssfm_sixpt_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples6, int32_t num_samples,
                     double* models, int32_t* num_models) {
  if (!h || !rays || !samples6 || !models || !num_models || n < 6 || num_samples < 0)
    return fail(SSFM_ERR_INVALID, "ssfm_sixpt_solve: bad argument");
  for (int i = 0; i < 6 * num_samples; ++i)
    if (samples6[i] < 0 || samples6[i] >= n) return fail(SSFM_ERR_INVALID, "ssfm_sixpt_solve: sample index out of range");
  if (num_samples == 0) return SSFM_OK;
  SSFM_CK(cudaSetDevice(h->device));
  DevBuf<double> d_rays, d_models, d_M;
  DevBuf<int> d_samples, d_nm;
  SSFM_CK(d_rays.ensure((size_t)n * 6));
  SSFM_CK(d_M.ensure(((size_t)num_samples + kSixSamplesPerBlock) * sixc::kMSize));
  SSFM_CK(d_models.ensure((size_t)num_samples * kSixMaxModels * 7));
  SSFM_CK(d_samples.ensure((size_t)num_samples * 6));
  SSFM_CK(d_nm.ensure(num_samples));
  SSFM_CK(cudaMemcpyAsync(d_rays.p, rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(d_samples.p, samples6, sizeof(int) * 6 * num_samples, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemsetAsync(d_models.p, 0, sizeof(double) * num_samples * kSixMaxModels * 7, h->stream));
  SSFM_CK(cudaFuncSetAttribute(k_sixpt_solve_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSixSolveSmem));
  k_sixpt_solve_samples<<<(num_samples + kSixSamplesPerBlock - 1) / kSixSamplesPerBlock, kSixSolveThreads, kSixSolveSmem, h->stream>>>(
      d_rays.p, d_samples.p, num_samples, d_models.p, d_nm.p, d_M.p);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(models, d_models.p, sizeof(double) * num_samples * kSixMaxModels * 7, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(num_models, d_nm.p, sizeof(int) * num_samples, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  d_rays.release(); d_models.release(); d_samples.release(); d_nm.release(); d_M.release();
  return SSFM_OK;
}

int ssfm_selection_sample(uint32_t seed, uint32_t pair, uint32_t hypothesis, int32_t n_total, int32_t k, int32_t* idx) {
  if (!idx || k <= 0 || n_total < k) return fail(SSFM_ERR_INVALID, "ssfm_selection_sample: need 0 < k <= n_total");
  knuth_sample(seed, pair, hypothesis, n_total, k, idx);
  return SSFM_OK;
}

#define SSFM_TMP(T, name, count)                                  \
  DevBuf<T> name;                                                 \
  {                                                               \
    cudaError_t e__ = name.ensure(std::max<size_t>((count), 1));  \
    if (e__ != cudaSuccess) return fail(SSFM_ERR_OOM, "cudaMalloc failed in hook"); \
  }
struct TmpGuard {
  std::vector<void*> ptrs;
  ~TmpGuard() {
    for (void* p : ptrs) cudaFree(p);
  }
};
#define SSFM_KEEP(guard, buf) \
  guard.ptrs.push_back(buf.p); \
  buf.cap = 0;

static int minimal_solve_impl(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t ns, int32_t solver,
                              int skip, double* models, int32_t* num_models);
int ssfm_minimal_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t ns, int32_t solver,
                       double* models, int32_t* num_models) {
  return minimal_solve_impl(h, rays, n, samples, ns, solver, 0, models, num_models);
}
int ssfm_minimal_solve_opt(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t ns, const SsfmOptions* opt,
                           double* models, int32_t* num_models) {
  if (int rc = check_options(opt)) return rc;
  return minimal_solve_impl(h, rays, n, samples, ns, opt->solver, opt->complex_root_models == SSFM_COMPLEX_SKIP, models, num_models);
}
static int minimal_solve_impl(ssfm_handle h, const double* rays, int32_t n, const int32_t* samples, int32_t ns, int32_t solver,
                              int skip, double* models, int32_t* num_models) {
  if (!h || !rays || !samples || !models || !num_models || n < 3 || ns < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (solver < 0 || solver > 2) return fail(SSFM_ERR_INVALID, "unknown solver kind");
  for (int i = 0; i < 3 * ns; ++i)
    if (samples[i] < 0 || samples[i] >= n) return fail(SSFM_ERR_INVALID, "sample index out of range");
  if (ns == 0) return SSFM_OK;
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, d_rays, (size_t)n * 6) SSFM_KEEP(g, d_rays)
  SSFM_TMP(int, d_s, (size_t)ns * 3) SSFM_KEEP(g, d_s)
  SSFM_TMP(double, d_m, (size_t)ns * 24) SSFM_KEEP(g, d_m)
  SSFM_TMP(int, d_nm, (size_t)ns) SSFM_KEEP(g, d_nm)
  double* dr = (double*)g.ptrs[0]; int* ds = (int*)g.ptrs[1]; double* dm = (double*)g.ptrs[2]; int* dn = (int*)g.ptrs[3];
  SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(ds, samples, sizeof(int) * 3 * ns, cudaMemcpyHostToDevice, h->stream));
  const int blocks = (ns + 63) / 64;
  if (solver == 0) k_solve_samples<0><<<blocks, 64, 0, h->stream>>>(dr, ds, ns, skip, dm, dn);
  else if (solver == 1) k_solve_samples<1><<<blocks, 64, 0, h->stream>>>(dr, ds, ns, skip, dm, dn);
  else k_solve_samples<2><<<blocks, 64, 0, h->stream>>>(dr, ds, ns, skip, dm, dn);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(models, dm, sizeof(double) * 24 * ns, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(num_models, dn, sizeof(int) * ns, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_score(ssfm_handle h, const double* models6, int32_t M, const double* rays, int32_t n, double thr2, float* scores,
               int32_t* counts, float* kernel_ms) {
  if (n < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  const int64_t offs[2] = {0, n};
  return ssfm_score_pairs(h, models6, M, rays, offs, 1, thr2, scores, counts, kernel_ms);
}

int ssfm_score_pairs(ssfm_handle h, const double* models6, int32_t M, const double* rays, const int64_t* offsets, int32_t num_pairs,
                     double thr2, float* scores, int32_t* counts, float* kernel_ms) {
  if (!h || !models6 || !rays || !offsets || !scores || !counts || M < 0 || num_pairs < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  refresh_knobs(h);
  if (M == 0 || num_pairs == 0) return SSFM_OK;
  if (offsets[0] != 0) return fail(SSFM_ERR_INVALID, "offsets[0] must be 0");
  long long nmax = 0;
  for (int p = 0; p < num_pairs; ++p) {
    if (offsets[p + 1] < offsets[p]) return fail(SSFM_ERR_INVALID, "offsets must be non-decreasing");
    nmax = std::max<long long>(nmax, offsets[p + 1] - offsets[p]);
  }
  const long long n = offsets[num_pairs];
  if (nmax > 0x7fffffffLL) return fail(SSFM_ERR_INVALID, "pair too large");
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_rays, (size_t)n * 6) SSFM_KEEP(g, b_rays)
  SSFM_TMP(float4, b_u, (size_t)n) SSFM_KEEP(g, b_u)
  SSFM_TMP(float4, b_v, (size_t)n) SSFM_KEEP(g, b_v)
  SSFM_TMP(double, b_m, (size_t)num_pairs * M * 6) SSFM_KEEP(g, b_m)
  double* dr = (double*)g.ptrs[0]; float4* du = (float4*)g.ptrs[1]; float4* dv = (float4*)g.ptrs[2]; double* dm = (double*)g.ptrs[3];
  SSFM_TMP(float4, b_uv, (size_t)n) SSFM_KEEP(g, b_uv)
  SSFM_TMP(int, b_flag, 1) SSFM_KEEP(g, b_flag)
  SSFM_TMP(long long, b_off, (size_t)num_pairs + 1) SSFM_KEEP(g, b_off)
  float4* duv = (float4*)g.ptrs[4]; int* dflag = (int*)g.ptrs[5]; long long* doff = (long long*)g.ptrs[6];
  if (n > 0) SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dm, models6, sizeof(double) * 6 * (size_t)M * num_pairs, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(doff, offsets, sizeof(long long) * ((size_t)num_pairs + 1), cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemsetAsync(dflag, 0, sizeof(int), h->stream));
  if (n > 0) k_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dr, n, duv, (double*)nullptr, dflag);
  SSFM_CK(cudaMemcpyAsync(h->h_count + 3, dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  const bool unit_z = h->h_count[3] == 0 && !h->knobs.no_unitz;
  if (!unit_z && n > 0) k_pack_general<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dr, n, du, dv);
  // Correspondences of every pair are split into chunks (multiples of the tile) so that the grid is ONE full wave of CTAs:
  // gx * num_pairs CTAs per chunk layer, as many layers as fit the resident-CTA slots (rounded DOWN: a 19th layer on 18.5
  // layers' worth of slots would run as a second, almost empty wave and double the launch's duration).
  const int gx = (M + 4 * kScoreThreads - 1) / (4 * kScoreThreads);
  const int max_chunks = (int)std::max<long long>(1, (nmax + kTile - 1) / kTile);
  int per_sm = 8;
  if (unit_z) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_models<true>, kScoreThreads, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_models<false>, kScoreThreads, 0);
  const long long slots = (long long)std::max(per_sm, 1) * h->num_sms, layer = (long long)gx * num_pairs;
  int nchunks = (int)std::min<long long>(max_chunks, std::max<long long>(1, slots / layer));
  // chunk starts only need the 16-byte alignment of the bulk copies (one float4 per correspondence): a multiple of 64 keeps
  // the split even for small pairs, where whole 512-correspondence tiles would leave half the machine idle
  int chunk = (int)(((nmax + nchunks - 1) / nchunks + 63) / 64 * 64);
  if (chunk <= 0) chunk = 64;
  nchunks = (int)std::max<long long>(1, (nmax + chunk - 1) / chunk);
  SSFM_TMP(float, b_ps, (size_t)num_pairs * nchunks * M) SSFM_KEEP(g, b_ps)
  SSFM_TMP(int, b_pc, (size_t)num_pairs * nchunks * M) SSFM_KEEP(g, b_pc)
  SSFM_TMP(float, b_s, (size_t)num_pairs * M) SSFM_KEEP(g, b_s)
  SSFM_TMP(int, b_c, (size_t)num_pairs * M) SSFM_KEEP(g, b_c)
  float* ps = (float*)g.ptrs[7]; int* pc = (int*)g.ptrs[8]; float* ds = (float*)g.ptrs[9]; int* dc = (int*)g.ptrs[10];
  // warm-up launch (untimed), then the timed one
  for (int rep = 0; rep < 2; ++rep) {
    if (rep == 1) SSFM_CK(cudaEventRecord(h->ev[0], h->stream));
    dim3 grid(gx, nchunks, num_pairs);
    if (unit_z) k_score_models<true><<<grid, kScoreThreads, 0, h->stream>>>(duv, nullptr, doff, chunk, dm, M, (float)thr2, ps, pc);
    else k_score_models<false><<<grid, kScoreThreads, 0, h->stream>>>(du, dv, doff, chunk, dm, M, (float)thr2, ps, pc);
    k_reduce_parts<<<dim3((M + 255) / 256, num_pairs), 256, 0, h->stream>>>(ps, pc, nchunks, M, ds, dc);
    if (rep == 1) SSFM_CK(cudaEventRecord(h->ev[1], h->stream));
  }
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(scores, ds, sizeof(float) * (size_t)M * num_pairs, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(counts, dc, sizeof(int) * (size_t)M * num_pairs, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  if (kernel_ms) SSFM_CK(cudaEventElapsedTime(kernel_ms, h->ev[0], h->ev[1]));
  return SSFM_OK;
}

int ssfm_score_exact(ssfm_handle h, const double* E9, int32_t M, const double* rays, int32_t n, double thr2, double* scores,
                     int32_t* counts) {
  if (!h || !E9 || !rays || !scores || !counts || M < 0 || n < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (M == 0) return SSFM_OK;
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_rays, (size_t)n * 6) SSFM_KEEP(g, b_rays)
  SSFM_TMP(double, b_e, (size_t)M * 9) SSFM_KEEP(g, b_e)
  SSFM_TMP(double, b_s, (size_t)M) SSFM_KEEP(g, b_s)
  SSFM_TMP(int, b_c, (size_t)M) SSFM_KEEP(g, b_c)
  double* dr = (double*)g.ptrs[0]; double* de = (double*)g.ptrs[1]; double* ds = (double*)g.ptrs[2]; int* dc = (int*)g.ptrs[3];
  if (n > 0) SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(de, E9, sizeof(double) * 9 * (size_t)M, cudaMemcpyHostToDevice, h->stream));
  k_score_exact<<<(M + 3) / 4, 128, 0, h->stream>>>(de, M, dr, n, thr2, ds, dc);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(scores, ds, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(counts, dc, sizeof(int) * M, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_least_squares(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx, const int32_t* sample_offsets,
                       int32_t nprob, int32_t inward, double* E9) {
  if (!h || !rays || !sample_idx || !sample_offsets || !E9 || nprob < 0 || n < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (nprob == 0) return SSFM_OK;
  const int total = sample_offsets[nprob];
  for (int i = 0; i < total; ++i)
    if (sample_idx[i] < 0 || sample_idx[i] >= n) return fail(SSFM_ERR_INVALID, "sample index out of range");
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_rays, (size_t)n * 6) SSFM_KEEP(g, b_rays)
  SSFM_TMP(int, b_i, (size_t)total) SSFM_KEEP(g, b_i)
  SSFM_TMP(int, b_o, (size_t)nprob + 1) SSFM_KEEP(g, b_o)
  SSFM_TMP(double, b_e, (size_t)nprob * 9) SSFM_KEEP(g, b_e)
  double* dr = (double*)g.ptrs[0]; int* di = (int*)g.ptrs[1]; int* dof = (int*)g.ptrs[2]; double* de = (double*)g.ptrs[3];
  if (n > 0) SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  if (total > 0) SSFM_CK(cudaMemcpyAsync(di, sample_idx, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dof, sample_offsets, sizeof(int) * ((size_t)nprob + 1), cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(de, E9, sizeof(double) * 9 * (size_t)nprob, cudaMemcpyHostToDevice, h->stream));
  k_least_squares<<<(nprob + 3) / 4, 128, 0, h->stream>>>(dr, di, dof, nprob, inward, de);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(E9, de, sizeof(double) * 9 * (size_t)nprob, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_sixpt_least_squares(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx,
                             const int32_t* sample_offsets, int32_t nprob, double* models7) {
  if (!h || !rays || !sample_idx || !sample_offsets || !models7 || nprob < 0 || n < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (nprob == 0) return SSFM_OK;
  const int total = sample_offsets[nprob];
  for (int i = 0; i < total; ++i)
    if (sample_idx[i] < 0 || sample_idx[i] >= n) return fail(SSFM_ERR_INVALID, "sample index out of range");
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_rays, (size_t)n * 6) SSFM_KEEP(g, b_rays)
  SSFM_TMP(int, b_i, (size_t)total) SSFM_KEEP(g, b_i)
  SSFM_TMP(int, b_o, (size_t)nprob + 1) SSFM_KEEP(g, b_o)
  SSFM_TMP(double, b_m, (size_t)nprob * 7) SSFM_KEEP(g, b_m)
  double* dr = (double*)g.ptrs[0]; int* di = (int*)g.ptrs[1]; int* dof = (int*)g.ptrs[2]; double* dm = (double*)g.ptrs[3];
  if (n > 0) SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  if (total > 0) SSFM_CK(cudaMemcpyAsync(di, sample_idx, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dof, sample_offsets, sizeof(int) * ((size_t)nprob + 1), cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dm, models7, sizeof(double) * 7 * (size_t)nprob, cudaMemcpyHostToDevice, h->stream));
  k_sixpt_least_squares<<<(nprob + 63) / 64, 64, 0, h->stream>>>(dr, di, dof, nprob, dm);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(models7, dm, sizeof(double) * 7 * (size_t)nprob, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_non_minimal_solve(ssfm_handle h, const double* rays, int32_t n, const int32_t* sample_idx,
                           const int32_t* sample_offsets, int32_t nprob, double* E9, int32_t* ok) {
  if (!h || !rays || !sample_idx || !sample_offsets || !E9 || !ok || nprob < 0 || n < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (nprob == 0) return SSFM_OK;
  const int total = sample_offsets[nprob];
  for (int i = 0; i < total; ++i)
    if (sample_idx[i] < 0 || sample_idx[i] >= n) return fail(SSFM_ERR_INVALID, "sample index out of range");
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_rays, (size_t)n * 6) SSFM_KEEP(g, b_rays)
  SSFM_TMP(int, b_i, (size_t)total) SSFM_KEEP(g, b_i)
  SSFM_TMP(int, b_o, (size_t)nprob + 1) SSFM_KEEP(g, b_o)
  SSFM_TMP(double, b_e, (size_t)nprob * 9) SSFM_KEEP(g, b_e)
  SSFM_TMP(int, b_k, (size_t)nprob) SSFM_KEEP(g, b_k)
  double* dr = (double*)g.ptrs[0]; int* di = (int*)g.ptrs[1]; int* dof = (int*)g.ptrs[2]; double* de = (double*)g.ptrs[3]; int* dk = (int*)g.ptrs[4];
  if (n > 0) SSFM_CK(cudaMemcpyAsync(dr, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  if (total > 0) SSFM_CK(cudaMemcpyAsync(di, sample_idx, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dof, sample_offsets, sizeof(int) * ((size_t)nprob + 1), cudaMemcpyHostToDevice, h->stream));
  k_non_minimal<<<(nprob + 3) / 4, 128, 0, h->stream>>>(dr, n, di, dof, nprob, de, dk);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(E9, de, sizeof(double) * 9 * (size_t)nprob, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(ok, dk, sizeof(int) * (size_t)nprob, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_decompose(ssfm_handle h, const double* E9, int32_t num, int32_t inward, double* r3, double* t3) {
  if (!h || !E9 || !r3 || !t3 || num < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (num == 0) return SSFM_OK;
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_e, (size_t)num * 9) SSFM_KEEP(g, b_e)
  SSFM_TMP(double, b_r, (size_t)num * 3) SSFM_KEEP(g, b_r)
  SSFM_TMP(double, b_t, (size_t)num * 3) SSFM_KEEP(g, b_t)
  double* de = (double*)g.ptrs[0]; double* dr = (double*)g.ptrs[1]; double* dt = (double*)g.ptrs[2];
  SSFM_CK(cudaMemcpyAsync(de, E9, sizeof(double) * 9 * (size_t)num, cudaMemcpyHostToDevice, h->stream));
  k_decompose<<<(num + 127) / 128, 128, 0, h->stream>>>(de, num, inward, dr, dt);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(r3, dr, sizeof(double) * 3 * (size_t)num, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaMemcpyAsync(t3, dt, sizeof(double) * 3 * (size_t)num, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_decompose_rescaled(ssfm_handle h, const double* E9, int32_t num, const double* scales, int32_t nscales, int32_t inward,
                            double* r3) {
  if (!h || !E9 || !scales || !r3 || num < 0 || nscales < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (num == 0 || nscales == 0) return SSFM_OK;
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(double, b_e, (size_t)num * 9) SSFM_KEEP(g, b_e)
  SSFM_TMP(double, b_s, (size_t)nscales) SSFM_KEEP(g, b_s)
  SSFM_TMP(double, b_r, (size_t)num * nscales * 3) SSFM_KEEP(g, b_r)
  double* de = (double*)g.ptrs[0]; double* dsc = (double*)g.ptrs[1]; double* dr = (double*)g.ptrs[2];
  SSFM_CK(cudaMemcpyAsync(de, E9, sizeof(double) * 9 * (size_t)num, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dsc, scales, sizeof(double) * (size_t)nscales, cudaMemcpyHostToDevice, h->stream));
  const long long total = (long long)num * nscales;
  k_decompose_rescaled<<<(unsigned)((total + 127) / 128), 128, 0, h->stream>>>(de, num, dsc, nscales, inward, dr);
  SSFM_CK(cudaGetLastError());
  SSFM_CK(cudaMemcpyAsync(r3, dr, sizeof(double) * 3 * (size_t)total, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_lo_shuffle(ssfm_handle h, uint32_t seed, int32_t ncalls, const int32_t* sizes, const int32_t* targets, int32_t* out) {
  if (!h || !sizes || !targets || !out || ncalls < 0) return fail(SSFM_ERR_INVALID, "bad argument");
  if (ncalls == 0) return SSFM_OK;
  int maxsz = 1, total = 0;
  for (int i = 0; i < ncalls; ++i) {
    if (sizes[i] < 0 || targets[i] < 0 || targets[i] > sizes[i]) return fail(SSFM_ERR_INVALID, "need 0 <= target <= size");
    maxsz = std::max(maxsz, sizes[i]);
    total += targets[i];
  }
  SSFM_CK(cudaSetDevice(h->device));
  TmpGuard g;
  SSFM_TMP(int, b_sz, (size_t)ncalls) SSFM_KEEP(g, b_sz)
  SSFM_TMP(int, b_tg, (size_t)ncalls) SSFM_KEEP(g, b_tg)
  SSFM_TMP(uint32_t, b_mt, 625) SSFM_KEEP(g, b_mt)
  SSFM_TMP(int, b_w, (size_t)maxsz) SSFM_KEEP(g, b_w)
  SSFM_TMP(int, b_o, (size_t)total) SSFM_KEEP(g, b_o)
  int* dsz = (int*)g.ptrs[0]; int* dtg = (int*)g.ptrs[1]; uint32_t* dmt = (uint32_t*)g.ptrs[2]; int* dw = (int*)g.ptrs[3]; int* dout = (int*)g.ptrs[4];
  SSFM_CK(cudaMemcpyAsync(dsz, sizes, sizeof(int) * ncalls, cudaMemcpyHostToDevice, h->stream));
  SSFM_CK(cudaMemcpyAsync(dtg, targets, sizeof(int) * ncalls, cudaMemcpyHostToDevice, h->stream));
  k_lo_shuffle<<<1, 32, 0, h->stream>>>(seed, ncalls, dsz, dtg, dmt, dw, dout);
  SSFM_CK(cudaGetLastError());
  if (total > 0) SSFM_CK(cudaMemcpyAsync(out, dout, sizeof(int) * total, cudaMemcpyDeviceToHost, h->stream));
  SSFM_CK(cudaStreamSynchronize(h->stream));
  return SSFM_OK;
}

int ssfm_measure_fp32_peaks(ssfm_handle h, double* scalar_tflops, double* packed_tflops) {
  if (!h || !scalar_tflops || !packed_tflops) return fail(SSFM_ERR_INVALID, "bad argument");
  SSFM_CK(cudaSetDevice(h->device));
  const int threads = 256, blocks = h->num_sms * 8, iters = 4096;
  TmpGuard g;
  SSFM_TMP(float, b_o, (size_t)threads * blocks) SSFM_KEEP(g, b_o)
  float* d = (float*)g.ptrs[0];
  double best[2] = {0.0, 0.0};
  for (int packed = 0; packed < 2; ++packed)
    for (int rep = 0; rep < 5; ++rep) {
      SSFM_CK(cudaEventRecord(h->ev[0], h->stream));
      if (packed) k_fma2_peak<<<blocks, threads, 0, h->stream>>>(d, iters, 1.000001f, 1e-7f);
      else k_fma_peak<<<blocks, threads, 0, h->stream>>>(d, iters, 1.000001f, 1e-7f);
      SSFM_CK(cudaEventRecord(h->ev[1], h->stream));
      SSFM_CK(cudaStreamSynchronize(h->stream));
      float ms = 0.f;
      SSFM_CK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
      // 64 FMA per iteration and thread (scalar) or 64 packed FMA = 128 FMA (packed), 2 flop each
      const double flops = 2.0 * 64.0 * (packed ? 2.0 : 1.0) * (double)iters * threads * blocks;
      if (rep > 0) best[packed] = std::max(best[packed], flops / (ms * 1e-3) / 1e12);
    }
  SSFM_CK(cudaGetLastError());
  *scalar_tflops = best[0];
  *packed_tflops = best[1];
  return SSFM_OK;
}

int ssfm_measure_fp32_peak(ssfm_handle h, double* tflops) {  // the roofline denominator: the higher of the two
  if (!tflops) return fail(SSFM_ERR_INVALID, "bad argument");
  double a = 0.0, b = 0.0;
  if (int rc = ssfm_measure_fp32_peaks(h, &a, &b)) return rc;
  *tflops = std::max(a, b);
  return SSFM_OK;
}

}  // extern "C"
