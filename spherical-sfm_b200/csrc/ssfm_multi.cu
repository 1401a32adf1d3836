// ssfm_multi.cu -- one batch, N devices, ONE process: the multi-GPU entry of libssfm_b200.so.
//
// The reference caller is a single C++ process that fans the pair list out over host threads
// (`#pragma omp parallel for` over all i<j pairs, examples/spherical_sfm_tools.cpp:321-332) and has no
// cross-pair state (:332-420).  Here the same call fans out over GPUs: the pair list is cut into one
// contiguous shard per device, balanced by correspondence count; one host thread per device drives that
// device's engine through the ordinary single-device entry (chunk-pipelined H2D from the caller's buffer
// over that device's own PCIe link, all kernels, D2H straight into the caller's result table and flags).
// No data-path collective exists: pair p draws from Philox key (seed, first_pair_id + p) wherever it runs,
// so the table is byte-identical to the single-device one.
//
// Only when device-side consumers want the whole table in every GPU's HBM is there an exchange:
// ssfm_multi_allgather_results() is an all-gather-v of the 168-byte per-pair records over NVLink/NVSwitch
// (one ncclBroadcast per shard inside a group; NCCL is dlopen'ed on first use, the library has no link-time
// dependency on it).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ssfm.h"

extern "C" void ssfm_internal_set_error(const char* msg);  // ssfm_engine.cu

namespace {

int mfail(int code, const std::string& msg) {
  ssfm_internal_set_error(msg.c_str());  // shows up in ssfm_last_error() of the calling thread
  return code;
}

// ---- the slice of NCCL's C API that is used, resolved at run time ----
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string* why) {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      *why = std::string("NCCL is not loadable (") + dlerror() + ")";
      return false;
    }
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Broadcast) {
      *why = "NCCL symbols missing";
      return false;
    }
    return true;
  }
};

}  // namespace

struct ssfm_multi {
  std::vector<int> devices;
  std::vector<ssfm_handle> engines;
  std::vector<int> bounds;  // pair bounds of the last call's shards (devices.size() + 1)
  int num_pairs = 0;
  std::vector<int> rc;
  std::vector<std::string> err;
  // all-gather state
  NcclApi nccl;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  std::vector<void*> tables;  // per device: num_pairs x SsfmPairResult
  size_t table_cap = 0;
};

extern "C" {

int ssfm_partition_pairs(const int64_t* offsets, int32_t num_pairs, int32_t num_shards, int32_t* bounds) {
  if (!offsets || !bounds || num_pairs < 0 || num_shards <= 0) return mfail(SSFM_ERR_INVALID, "ssfm_partition_pairs: bad argument");
  const int64_t total = offsets[num_pairs] - offsets[0];
  bounds[0] = 0;
  for (int r = 1; r < num_shards; ++r) {
    // first pair whose start offset reaches r/num_shards of the correspondences (never before the previous bound)
    const double target = (double)offsets[0] + (double)total * (double)r / (double)num_shards;
    const int64_t* it = std::lower_bound(offsets, offsets + num_pairs + 1, target, [](int64_t o, double t) { return (double)o < t; });
    int b = (int)(it - offsets);
    b = std::min(std::max(b, bounds[r - 1]), num_pairs);
    bounds[r] = b;
  }
  bounds[num_shards] = num_pairs;
  return SSFM_OK;
}

int ssfm_multi_create(const int32_t* devices, int32_t num_devices, ssfm_multi_handle* out) {
  if (!out) return mfail(SSFM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!devices || num_devices <= 0) return mfail(SSFM_ERR_INVALID, "ssfm_multi_create: need at least one device");
  for (int i = 0; i < num_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return mfail(SSFM_ERR_INVALID, "ssfm_multi_create: duplicate device");
  ssfm_multi* m = new ssfm_multi();
  m->devices.assign(devices, devices + num_devices);
  m->engines.assign(num_devices, nullptr);
  m->rc.assign(num_devices, SSFM_OK);
  m->err.assign(num_devices, std::string());
  for (int i = 0; i < num_devices; ++i) {
    const int rc = ssfm_create(devices[i], &m->engines[i]);
    if (rc != SSFM_OK) {
      const std::string why = ssfm_last_error();
      ssfm_multi_destroy(m);
      return mfail(rc, "device " + std::to_string(devices[i]) + ": " + why);
    }
  }
  *out = m;
  return SSFM_OK;
}

void ssfm_multi_destroy(ssfm_multi_handle m) {
  if (!m) return;
  for (size_t i = 0; i < m->comms.size(); ++i)
    if (m->comms[i]) m->nccl.CommDestroy(m->comms[i]);
  for (size_t i = 0; i < m->tables.size(); ++i) {
    cudaSetDevice(m->devices[i]);
    if (m->tables[i]) cudaFree(m->tables[i]);
    if (i < m->streams.size() && m->streams[i]) cudaStreamDestroy(m->streams[i]);
  }
  for (ssfm_handle h : m->engines)
    if (h) ssfm_destroy(h);
  delete m;
}

int32_t ssfm_multi_num_devices(ssfm_multi_handle m) { return m ? (int32_t)m->devices.size() : 0; }

int ssfm_multi_get_stats(ssfm_multi_handle m, int32_t index, SsfmRunStats* stats, int32_t* first_pair, int32_t* num_pairs) {
  if (!m || !stats || index < 0 || index >= (int)m->engines.size()) return mfail(SSFM_ERR_INVALID, "ssfm_multi_get_stats: bad argument");
  if (first_pair) *first_pair = m->bounds.empty() ? 0 : m->bounds[index];
  if (num_pairs) *num_pairs = m->bounds.empty() ? 0 : m->bounds[index + 1] - m->bounds[index];
  return ssfm_get_stats(m->engines[index], stats);
}

int ssfm_estimate_pairs_multi(ssfm_multi_handle m, const SsfmBatch* batch, const SsfmOptions* opt, SsfmPairResult* results,
                              uint8_t* inlier_flags) {
  if (!m || !batch || !opt || !results) return mfail(SSFM_ERR_INVALID, "ssfm_estimate_pairs_multi: NULL argument");
  if (batch->rays_on_device) return mfail(SSFM_ERR_INVALID, "ssfm_estimate_pairs_multi takes host rays (one buffer, N devices)");
  if (batch->num_pairs < 0 || (batch->num_pairs > 0 && (!batch->offsets || !batch->rays)))
    return mfail(SSFM_ERR_INVALID, "ssfm_estimate_pairs_multi: bad batch");
  const int nd = (int)m->devices.size();
  const int P = batch->num_pairs;
  m->num_pairs = P;
  m->bounds.assign(nd + 1, 0);
  static const int64_t zero = 0;
  if (int rc = ssfm_partition_pairs(P > 0 ? batch->offsets : &zero, P, nd, m->bounds.data())) return rc;
  auto work = [&](int i) {
    const int p0 = m->bounds[i], p1 = m->bounds[i + 1];
    const int64_t c0 = P > 0 ? batch->offsets[p0] : 0;
    std::vector<int64_t> offs((size_t)(p1 - p0) + 1);
    for (int p = p0; p <= p1; ++p) offs[p - p0] = (P > 0 ? batch->offsets[p] : 0) - c0;
    SsfmBatch b;
    b.num_pairs = p1 - p0;
    b.offsets = offs.data();
    const size_t elem = batch->ray_format == SSFM_RAYS_F32 ? sizeof(float) : sizeof(double);
    b.rays = batch->rays ? reinterpret_cast<const double*>(reinterpret_cast<const char*>(batch->rays) + elem * 6 * (size_t)c0) : nullptr;
    b.rays_on_device = 0;
    b.ray_format = batch->ray_format;
    SsfmOptions o = *opt;
    o.first_pair_id = opt->first_pair_id + (uint32_t)p0;  // pair p keeps its Philox key wherever it runs
    m->rc[i] = ssfm_estimate_pairs(m->engines[i], &b, &o, results + p0, inlier_flags ? inlier_flags + c0 : nullptr);
    if (m->rc[i] != SSFM_OK) m->err[i] = ssfm_last_error();  // thread-local in that thread: copy it out
  };
  std::vector<std::thread> threads;
  for (int i = 1; i < nd; ++i) threads.emplace_back(work, i);
  work(0);
  for (auto& t : threads) t.join();
  for (int i = 0; i < nd; ++i)
    if (m->rc[i] != SSFM_OK) return mfail(m->rc[i], "device " + std::to_string(m->devices[i]) + ": " + m->err[i]);
  return SSFM_OK;
}

int ssfm_multi_allgather_results(ssfm_multi_handle m, void** dev_tables, int32_t* num_pairs) {
  if (!m || !dev_tables) return mfail(SSFM_ERR_INVALID, "ssfm_multi_allgather_results: NULL argument");
  if (m->bounds.empty()) return mfail(SSFM_ERR_INVALID, "ssfm_multi_allgather_results: no results (call ssfm_estimate_pairs_multi first)");
  const int nd = (int)m->devices.size();
  const size_t P = (size_t)std::max(m->num_pairs, 1);
#define SSFM_MCK(call)                                                                                       \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) return mfail(e__ == cudaErrorMemoryAllocation ? SSFM_ERR_OOM : SSFM_ERR_CUDA,    \
                                         std::string(#call) + ": " + cudaGetErrorString(e__));               \
  } while (0)
  if (m->tables.empty()) {
    m->tables.assign(nd, nullptr);
    m->streams.assign(nd, nullptr);
    for (int i = 0; i < nd; ++i) {
      SSFM_MCK(cudaSetDevice(m->devices[i]));
      SSFM_MCK(cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking));
    }
  }
  if (P > m->table_cap) {
    for (int i = 0; i < nd; ++i) {
      SSFM_MCK(cudaSetDevice(m->devices[i]));
      if (m->tables[i]) cudaFree(m->tables[i]);
      m->tables[i] = nullptr;
      SSFM_MCK(cudaMalloc(&m->tables[i], P * sizeof(SsfmPairResult)));
    }
    m->table_cap = P;
  }
  // each device puts its shard's rows at their global position
  for (int i = 0; i < nd; ++i) {
    void* src = nullptr;
    int32_t n = 0;
    if (int rc = ssfm_device_results(m->engines[i], &src, &n)) return mfail(rc, ssfm_last_error());
    if (n != m->bounds[i + 1] - m->bounds[i]) return mfail(SSFM_ERR_INVALID, "ssfm_multi_allgather_results: stale engine results");
    SSFM_MCK(cudaSetDevice(m->devices[i]));
    if (n > 0)
      SSFM_MCK(cudaMemcpyAsync((char*)m->tables[i] + (size_t)m->bounds[i] * sizeof(SsfmPairResult), src, (size_t)n * sizeof(SsfmPairResult),
                               cudaMemcpyDeviceToDevice, m->streams[i]));
  }
  if (nd > 1) {
    std::string why;
    if (!m->nccl.load(&why)) return mfail(SSFM_ERR_CUDA, why);
    if (m->comms.empty()) {
      m->comms.assign(nd, nullptr);
      const ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), nd, m->devices.data());
      if (r != 0) {
        m->comms.clear();
        return mfail(SSFM_ERR_CUDA, std::string("ncclCommInitAll: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(r) : "failed"));
      }
    }
    // all-gather-v: shard r is broadcast from device r to everyone, all shards in one group
    ncclResult_t r = m->nccl.GroupStart();
    for (int root = 0; root < nd && r == 0; ++root) {
      const size_t off = (size_t)m->bounds[root] * sizeof(SsfmPairResult);
      const size_t bytes = (size_t)(m->bounds[root + 1] - m->bounds[root]) * sizeof(SsfmPairResult);
      if (bytes == 0) continue;
      for (int i = 0; i < nd && r == 0; ++i)
        r = m->nccl.Broadcast((char*)m->tables[i] + off, (char*)m->tables[i] + off, bytes, 0 /* ncclInt8/ncclChar */, root, m->comms[i],
                              m->streams[i]);
    }
    const ncclResult_t r2 = m->nccl.GroupEnd();
    if (r != 0 || r2 != 0)
      return mfail(SSFM_ERR_CUDA, std::string("NCCL all-gather: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(r != 0 ? r : r2) : "failed"));
  }
  for (int i = 0; i < nd; ++i) {
    SSFM_MCK(cudaSetDevice(m->devices[i]));
    SSFM_MCK(cudaStreamSynchronize(m->streams[i]));
    dev_tables[i] = m->tables[i];
  }
#undef SSFM_MCK
  if (num_pairs) *num_pairs = m->num_pairs;
  return SSFM_OK;
}

}  // extern "C"
