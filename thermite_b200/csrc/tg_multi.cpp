// tg_multi.cpp -- several GPUs of one host behind one C-ABI call (include/thermite_gpu.h: tg_multi_*).
//
// The reference's driver is one process that aligns read after read and writes its output in read order
// (src/main.rs:45-83 -> src/aligner.rs:22-120, serial loop :54-115).  Reads are independent, so a batch is cut
// into contiguous shards [g*N/G, (g+1)*N/G), one per GPU; the index is replicated once (ONE ncclBroadcast over NVLink --
// the only collective on the path) and there is no exchange step per batch.
//
// One host thread and one context per GPU.  The result is ONE pinned buffer set: every shard owns a segment of the
// record / operation pools and its slice of the per-read first / count arrays, the device writes first indices and
// operation offsets already rebased to the whole result, and every GPU's (early and late) D2H copies land at their
// final place.  Nothing is merged or touched by the host afterwards -- a merge pass over 8 x 250 MB of records costs more
// host memory bandwidth than the copies themselves.  Segment sizes follow the records-per-read rate of the previous batch
// (+12 %); a shard that needs more reports it and the batch is re-run with exact sizes (first batch of a new workload).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "tg_internal.h"

namespace {

// ---- NCCL, loaded at run time (the library has no link-time dependency on it) -------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;  // ncclSuccess = 0
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (getenv("TG_MULTI_NO_NCCL")) return false;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return false;
    CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
    Broadcast = (decltype(Broadcast))dlsym(handle, "ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast;
  }
};
constexpr int kNcclUint8 = 1;

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Worker {
  int slot = 0, device = 0;
  tg_ctx* ctx = nullptr;
  std::thread th;
  // job of the current generation
  uint32_t lo = 0, hi = 0;
  TgHostSegment seg{};
  TgShardStat stat{};
  tg_status status = TG_OK;
  std::string error;
};

}  // namespace

struct tg_multi {
  std::vector<int> devices;            // per slot
  std::vector<int> distinct;           // distinct devices in first-seen order
  std::vector<tg_index*> replica;      // per distinct device
  std::vector<void*> owned_blob;       // device memory of the replicas made here (distinct[1..])
  std::vector<Worker> w;
  std::string replication = "single";
  float bcast_ms = 0.f;
  tg_opts opts{};
  // job hand-off
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  int job_kind = 0;                    // 0: align shard, 1: create context
  const uint8_t* bases = nullptr;
  const uint64_t* offs = nullptr;
  // result buffers (pinned, portable: every device copies into them); two sets when tg_multi_set_result_buffers(m, 2)
  struct ResultSet {
    uint32_t* first = nullptr;
    uint32_t* count = nullptr;
    tg_aln_c* alns = nullptr;
    uint32_t* ops = nullptr;
    size_t cap_first = 0, cap_count = 0, cap_alns = 0, cap_ops = 0;
  } rs[2];
  int result_sets = 1, cur = 0;
  double rate_alns = 1.25, rate_ops = 8.0;  // records / operation words per read expected in the next batch
};

namespace {

tg_status cuda_fail(const char* what, cudaError_t e) { return tg_fail(TG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define MU_CHECK(call)                                    \
  do {                                                    \
    cudaError_t e_ = (call);                              \
    if (e_ != cudaSuccess) return cuda_fail(#call, e_);   \
  } while (0)

template <class T>
tg_status grow_pinned(T*& p, size_t& cap, size_t want) {
  if (want <= cap) return TG_OK;
  if (p) cudaFreeHost(p);
  p = nullptr; cap = 0;
  const size_t n = want + want / 4 + 1024;
  void* q = nullptr;
  cudaError_t e = cudaHostAlloc(&q, n * sizeof(T), cudaHostAllocPortable);
  if (e != cudaSuccess) return cuda_fail("cudaHostAlloc(result)", e);
  p = (T*)q; cap = n;
  return TG_OK;
}

void worker_main(tg_multi* m, int slot) {
  Worker& me = m->w[slot];
  cudaSetDevice(me.device);
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_go.wait(lk, [&] { return m->stop || m->generation != seen; });
      if (m->stop) return;
      seen = m->generation;
    }
    tg_status st = TG_OK;
    if (m->job_kind == 1) {
      tg_index* ix = nullptr;
      for (size_t d = 0; d < m->distinct.size(); d++)
        if (m->distinct[d] == me.device) ix = m->replica[d];
      st = tg_ctx_create(ix, &m->opts, &me.ctx);
    } else {
      st = tg_ctx_align_segment(me.ctx, m->bases, m->offs + me.lo, me.hi - me.lo, me.seg, &me.stat);
    }
    me.status = st;
    if (st != TG_OK) me.error = tg_last_error();  // (thread-local message: hand it to the caller's thread)
    {
      std::lock_guard<std::mutex> lk(m->mu);
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}

// runs one generation of jobs on all workers and returns the first error
tg_status run_jobs(tg_multi* m, int kind) {
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->job_kind = kind;
    m->pending = (int)m->w.size();
    m->generation++;
  }
  m->cv_go.notify_all();
  {
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->pending == 0; });
  }
  for (Worker& w : m->w)
    if (w.status != TG_OK) return tg_fail(w.status, "device " + std::to_string(w.device) + ": " + w.error);
  return TG_OK;
}

// index replicas on distinct[1..] from the one on distinct[0]
tg_status replicate(tg_multi* m, const tg_index_host* hix) {
  const size_t nb = hix->hdr()->device_bytes;
  const int nd = (int)m->distinct.size();
  tg_status st;
  if (nd == 1) return tg_index_create(hix, m->distinct[0], &m->replica[0]);
  // every replica lives in a device buffer owned here: the root's is uploaded, the others receive it
  m->owned_blob.assign(nd, nullptr);
  for (int d = 0; d < nd; d++) {
    MU_CHECK(cudaSetDevice(m->distinct[d]));
    MU_CHECK(cudaMalloc(&m->owned_blob[d], nb));
  }
  MU_CHECK(cudaSetDevice(m->distinct[0]));
  MU_CHECK(cudaMemcpy(m->owned_blob[0], hix->blob.data(), nb, cudaMemcpyHostToDevice));
  bool done = false;
  NcclApi nccl;
  if (nccl.load()) {
    std::vector<ncclComm_t> comms(nd, nullptr);
    std::vector<cudaStream_t> streams(nd, nullptr);
    std::vector<cudaEvent_t> ev(2, nullptr);
    ncclResult_t r = nccl.CommInitAll(comms.data(), nd, m->distinct.data());
    if (r == 0) {
      for (int d = 0; d < nd; d++) {
        MU_CHECK(cudaSetDevice(m->distinct[d]));
        MU_CHECK(cudaStreamCreateWithFlags(&streams[d], cudaStreamNonBlocking));
      }
      MU_CHECK(cudaSetDevice(m->distinct[0]));
      MU_CHECK(cudaEventCreate(&ev[0]));
      MU_CHECK(cudaEventCreate(&ev[1]));
      MU_CHECK(cudaEventRecord(ev[0], streams[0]));
      r = nccl.GroupStart();
      for (int d = 0; d < nd && r == 0; d++) {
        MU_CHECK(cudaSetDevice(m->distinct[d]));
        r = nccl.Broadcast(m->owned_blob[0], m->owned_blob[d], nb, kNcclUint8, 0, comms[d], streams[d]);
      }
      if (r == 0) r = nccl.GroupEnd();
      MU_CHECK(cudaSetDevice(m->distinct[0]));
      MU_CHECK(cudaEventRecord(ev[1], streams[0]));
      for (int d = 0; d < nd; d++) {
        MU_CHECK(cudaSetDevice(m->distinct[d]));
        MU_CHECK(cudaStreamSynchronize(streams[d]));
      }
      if (r == 0) {
        MU_CHECK(cudaEventElapsedTime(&m->bcast_ms, ev[0], ev[1]));
        m->replication = "nccl";
        done = true;
      }
      for (int d = 0; d < nd; d++) {
        cudaSetDevice(m->distinct[d]);
        if (streams[d]) cudaStreamDestroy(streams[d]);
        if (comms[d]) nccl.CommDestroy(comms[d]);
      }
      cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    }
  }
  if (!done) {  // no NCCL: device-to-device copies from the root (peer-to-peer over NVLink where enabled, else staged by the driver)
    const double t0 = now_ms();
    for (int d = 1; d < nd; d++) MU_CHECK(cudaMemcpyPeer(m->owned_blob[d], m->distinct[d], m->owned_blob[0], m->distinct[0], nb));
    for (int d = 0; d < nd; d++) { MU_CHECK(cudaSetDevice(m->distinct[d])); MU_CHECK(cudaDeviceSynchronize()); }
    m->bcast_ms = (float)(now_ms() - t0);
    m->replication = "peer-copy";
  }
  for (int d = 0; d < nd; d++)
    if ((st = tg_index_create_from_device_blob(m->owned_blob[d], nb, m->distinct[d], &m->replica[d])) != TG_OK) return st;
  return TG_OK;
}

}  // namespace

extern "C" {

void tg_multi_destroy(tg_multi* m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->stop = true;
  }
  m->cv_go.notify_all();
  for (Worker& w : m->w)
    if (w.th.joinable()) w.th.join();
  for (Worker& w : m->w)
    if (w.ctx) tg_ctx_destroy(w.ctx);
  for (tg_index* ix : m->replica)
    if (ix) tg_index_destroy(ix);
  for (size_t d = 0; d < m->owned_blob.size(); d++)
    if (m->owned_blob[d]) { cudaSetDevice(m->distinct[d]); cudaFree(m->owned_blob[d]); }
  for (auto& r : m->rs) {
    if (r.first) cudaFreeHost(r.first);
    if (r.count) cudaFreeHost(r.count);
    if (r.alns) cudaFreeHost(r.alns);
    if (r.ops) cudaFreeHost(r.ops);
  }
  delete m;
}

tg_status tg_multi_create(const tg_index_host* ix, const int* devices, int n_devices, const tg_opts* opts, tg_multi** out) {
  TG_GUARD_BEGIN
  if (!ix || !devices || !opts || !out || n_devices < 1) return tg_fail(TG_ERR_INVALID, "null argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return tg_fail(TG_ERR_CUDA, "no CUDA device: libthermite_gpu has no CPU fallback");
  for (int i = 0; i < n_devices; i++)
    if (devices[i] < 0 || devices[i] >= n_dev) return tg_fail(TG_ERR_INVALID, "device index out of range");
  tg_multi* m = new tg_multi();
  m->opts = *opts;
  m->devices.assign(devices, devices + n_devices);
  for (int d : m->devices)  // a device may be listed more than once: its slots share one replica
    if (std::find(m->distinct.begin(), m->distinct.end(), d) == m->distinct.end()) m->distinct.push_back(d);
  m->replica.assign(m->distinct.size(), nullptr);
  tg_status st = replicate(m, ix);
  if (st != TG_OK) { tg_multi_destroy(m); return st; }
  m->w.resize(n_devices);
  for (int s = 0; s < n_devices; s++) { m->w[s].slot = s; m->w[s].device = m->devices[s]; }
  for (int s = 0; s < n_devices; s++) m->w[s].th = std::thread(worker_main, m, s);
  // contexts (k-mer table build on every GPU) in parallel, each on its worker thread
  if ((st = run_jobs(m, 1)) != TG_OK) { tg_multi_destroy(m); return st; }
  *out = m;
  return TG_OK;
  TG_GUARD_END
}

int tg_multi_n_devices(const tg_multi* m) { return m ? (int)m->w.size() : 0; }

tg_status tg_multi_set_result_buffers(tg_multi* m, int n) {
  if (!m || (n != 1 && n != 2)) return tg_fail(TG_ERR_INVALID, "result buffers: 1 or 2");
  m->result_sets = n;
  return TG_OK;
}

const char* tg_multi_replication(const tg_multi* m, float* ms) {
  if (ms) *ms = m ? m->bcast_ms : 0.f;
  return m ? m->replication.c_str() : "";
}

tg_ctx* tg_multi_ctx(tg_multi* m, int g) { return (m && g >= 0 && g < (int)m->w.size()) ? m->w[g].ctx : nullptr; }

tg_status tg_multi_last_timing(const tg_multi* m, int g, double* wall_ms, float* seed_ms, float* extend_ms, float* dp_ms) {
  if (!m || g < 0 || g >= (int)m->w.size()) return tg_fail(TG_ERR_INVALID, "device slot out of range");
  const TgShardStat& s = m->w[g].stat;
  if (wall_ms) *wall_ms = s.wall_ms;
  if (seed_ms) *seed_ms = s.seed_ms;
  if (extend_ms) *extend_ms = s.extend_ms;
  if (dp_ms) *dp_ms = s.dp_ms;
  return TG_OK;
}

tg_status tg_multi_align_batch(tg_multi* m, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result_c* out) {
  TG_GUARD_BEGIN
  if (!m || !offs || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  if (n_reads == 0) return TG_OK;
  if (!bases && offs[n_reads] > offs[0]) return tg_fail(TG_ERR_INVALID, "null argument");
  const int G = (int)m->w.size();
  tg_status st;
  if (m->result_sets == 2) m->cur ^= 1;  // the previous call's result stays valid during this one
  tg_multi::ResultSet& R = m->rs[m->cur];
  if ((st = grow_pinned(R.first, R.cap_first, (size_t)n_reads)) != TG_OK) return st;
  if ((st = grow_pinned(R.count, R.cap_count, (size_t)n_reads)) != TG_OK) return st;
  m->bases = bases; m->offs = offs;
  std::vector<uint64_t> need_a(G, 0), need_o(G, 0);  // exact needs reported by a failed attempt
  for (int attempt = 0;; attempt++) {
    // contiguous shards, one pool segment each
    uint64_t a_base = 0, o_base = 0;
    std::vector<uint64_t> ab(G), ob(G), ac(G), oc(G);
    for (int g = 0; g < G; g++) {
      Worker& w = m->w[g];
      w.lo = (uint32_t)((uint64_t)n_reads * g / G); w.hi = (uint32_t)((uint64_t)n_reads * (g + 1) / G);
      const uint64_t ng = w.hi - w.lo;
      ac[g] = std::max<uint64_t>((uint64_t)(ng * m->rate_alns * 1.12) + 4096, need_a[g] + need_a[g] / 64 + 64);
      oc[g] = std::max<uint64_t>((uint64_t)(ng * m->rate_ops * 1.12) + 65536, need_o[g] + need_o[g] / 64 + 64);
      ab[g] = a_base; ob[g] = o_base;
      a_base += ac[g]; o_base += oc[g];
    }
    if (a_base > 0xFFFFFFFFull || o_base > 0xFFFFFFFFull)
      return tg_fail(TG_ERR_CAPACITY, "batch too large for 32-bit record / operation offsets: split the batch");
    if ((st = grow_pinned(R.alns, R.cap_alns, (size_t)a_base)) != TG_OK) return st;
    if ((st = grow_pinned(R.ops, R.cap_ops, (size_t)o_base)) != TG_OK) return st;
    for (int g = 0; g < G; g++) {
      Worker& w = m->w[g];
      w.seg.first = R.first + w.lo; w.seg.count = R.count + w.lo;
      w.seg.alns = R.alns + ab[g]; w.seg.ops = R.ops + ob[g];
      w.seg.alns_cap = ac[g]; w.seg.ops_cap = oc[g];
      w.seg.first_base = ab[g]; w.seg.ops_base = ob[g];
    }
    if ((st = run_jobs(m, 0)) != TG_OK) return st;
    bool overflow = false;
    for (int g = 0; g < G; g++) {
      const TgShardStat& s = m->w[g].stat;
      if (s.overflow) { overflow = true; need_a[g] = s.need_alns; need_o[g] = s.need_ops; }
      else { need_a[g] = s.n_alns; need_o[g] = s.n_ops; }
    }
    if (!overflow) {
      out->n_reads = n_reads; out->n_segments = (uint32_t)G;
      for (int g = 0; g < G; g++) {
        const TgShardStat& s = m->w[g].stat;
        out->n_alns += s.n_alns; out->n_ops += s.n_ops;
        out->swg_cells += s.swg_cells; out->swg_extensions += s.swg_extensions;
        out->seed_hits += s.seed_hits; out->n_smems += s.n_smems;
        if (s.n_alns) out->alns_extent = ab[g] + s.n_alns;
        if (s.n_ops) out->ops_extent = ob[g] + s.n_ops;
      }
      // sizes of the next batch's segments: the densest shard of this one
      double ra = 0.0, ro = 0.0;
      for (int g = 0; g < G; g++) {
        const double ng = (double)std::max<uint32_t>(m->w[g].hi - m->w[g].lo, 1u);
        ra = std::max(ra, (double)m->w[g].stat.n_alns / ng);
        ro = std::max(ro, (double)m->w[g].stat.n_ops / ng);
      }
      m->rate_alns = std::max(ra, 0.05); m->rate_ops = std::max(ro, 0.25);
      break;
    }
    if (attempt >= 3) return tg_fail(TG_ERR_CAPACITY, "result segments kept overflowing");
  }
  out->read_aln_first = R.first; out->read_aln_count = R.count;
  out->alns = R.alns; out->ops = R.ops;
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"
