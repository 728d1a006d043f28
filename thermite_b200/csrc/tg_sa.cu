// tg_sa.cu -- suffix array of the packed both-strand text, built on the GPU.
//
// Replaces `divsufsort64` in Index::create_from_files (/root/reference/src/index.rs:103-105): the plain lexicographic
// suffix order of the `$`-joined text with `$` < A < C < G < N < T and a suffix that is a proper prefix of another one
// first.  That order is unique, so the array equals the one the host SA-IS (`tg_sais`) and the reference produce.
//
// Method: prefix doubling over *unresolved* suffixes only.
//   step 0   key(i) = the 16 symbols at i (one funnel shift out of text4, codes + 1 so that "beyond the end" = 0 sorts
//            first), radix sort of (key, i)                                                     -> 16-order
//   step h   only suffixes that still share their h-prefix with another one are touched: they sit in contiguous ranges
//            of SA ("groups"); the ascending list of their SA positions is `pos`.  key(q) = (rank[i], rank[i+h]) for
//            i = sa[pos[q]], radix sort of the m pairs, written back through the same `pos` (a group's members stay
//            in the group's range because rank[i] = first SA position of the group + 1 leads the key), new ranks from
//            a max-scan over group heads, singletons leave `pos`.                               -> 2h-order
// A genome is almost fully resolved after step 0 (4^16 >> text length); what remains are repeat copies and the long
// N runs, so every later step sorts a few percent of the text and the number of steps is log2(longest repeat / 16).
// Sorting, scanning and compaction use CUB device primitives (library code, not on the alignment hot path); the
// kernels around them are below.  Memory: ~46 B per text symbol (4.3 GB for chr21's 93 M symbols).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <string>

#include "tg_internal.h"

namespace {

struct MaxU32 {
  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// key of step 0: 16 symbols starting at i, most significant nibble first, codes shifted to 1..6, 0 beyond the end
__global__ void k_sa_key16(const uint64_t* __restrict__ text4, uint32_t n, uint64_t* __restrict__ key,
                           uint32_t* __restrict__ val, uint32_t* __restrict__ pos) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w = i >> 4;
    const unsigned s = (unsigned)(i & 15) * 4;
    uint64_t v = text4[w];
    if (s) v = (v << s) | (text4[w + 1] >> (64 - s));  // text4 has >= 3 words of padding behind the text
    v += 0x1111111111111111ull;                        // codes 0..5 -> 1..6, no carry between nibbles
    const uint64_t left = (uint64_t)n - i;             // symbols from i to the end, >= 1
    if (left < 16) v &= ~0ull << (4 * (16 - left));
    key[i] = v;
    val[i] = (uint32_t)i;
    pos[i] = (uint32_t)i;
  }
}

// key of step h for the unresolved suffix at SA position pos[q]
__global__ void k_sa_key_pair(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ pos,
                              const uint32_t* __restrict__ rank, uint32_t m, uint32_t n, uint32_t h, unsigned bits,
                              uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = sa[pos[q]];
    const uint64_t j = (uint64_t)i + h;
    const uint32_t r2 = j < n ? rank[j] : 0u;
    key[q] = ((uint64_t)rank[i] << bits) | r2;
    val[q] = i;
  }
}

// group heads of the sorted keys: head[q] = pos[q] + 1 where a new key starts, else 0; act[q] = the key is shared
__global__ void k_sa_heads(const uint64_t* __restrict__ skey, const uint32_t* __restrict__ pos, uint32_t m,
                           uint32_t* __restrict__ head, uint8_t* __restrict__ act) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = skey[q];
    const bool first = q == 0 || skey[q - 1] != k;
    const bool last = q + 1 == m || skey[q + 1] != k;
    head[q] = first ? pos[q] + 1 : 0u;
    act[q] = !(first && last);
  }
}

// write the sorted suffixes back to their SA positions and give them the rank of their group (first position + 1)
__global__ void k_sa_commit(const uint32_t* __restrict__ sval, const uint32_t* __restrict__ pos,
                            const uint32_t* __restrict__ grp, uint32_t m, uint32_t* __restrict__ sa,
                            uint32_t* __restrict__ rank) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = sval[q];
    sa[pos[q]] = i;
    rank[i] = grp[q];
  }
}

struct SaBufs {
  void* p[16] = {nullptr};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st = nullptr;
  ~SaBufs() {
    for (void* q : p)
      if (q) cudaFree(q);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
  }
};

#define SA_TRY(call)                                                                                      \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

tg_status sa_device(const uint64_t* text4, uint64_t text_len, int device, uint32_t* sa_out, float* ms, uint32_t* steps) {
  if (!text4 || !sa_out) return tg_fail(TG_ERR_INVALID, "null argument");
  if (text_len == 0) return TG_OK;
  if (text_len >= (1ull << 31)) return tg_fail(TG_ERR_CAPACITY, "text must be < 2^31 symbols");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev)
    return tg_fail(TG_ERR_CUDA, "no CUDA device: the suffix-array builder of libthermite_gpu has no CPU fallback");
  SA_TRY(cudaSetDevice(device));
  const uint32_t n = (uint32_t)text_len;
  const size_t words = text_len / 16 + 4;  // layout of build_index(): T/16 + 4 words, zero padded
  unsigned bits = 1;
  while ((1ull << bits) <= (uint64_t)n) bits++;  // ranks are 1..n

  SaBufs b;
  SA_TRY(cudaStreamCreateWithFlags(&b.st, cudaStreamNonBlocking));
  SA_TRY(cudaEventCreate(&b.e0));
  SA_TRY(cudaEventCreate(&b.e1));
  uint64_t *d_text, *d_key, *d_skey;
  uint32_t *d_val, *d_sval, *d_sa, *d_rank, *d_head, *d_pos, *d_pos2, *d_m;
  uint8_t* d_act;
  void* d_tmp;
  size_t tmp_sort = 0, tmp_scan = 0, tmp_sel = 0;
  SA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                         (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, 0, 64, b.st));
  SA_TRY(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, MaxU32(),
                                        (int64_t)n, b.st));
  SA_TRY(cub::DeviceSelect::Flagged(nullptr, tmp_sel, (const uint32_t*)nullptr, (const uint8_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, b.st));
  const size_t tmp_bytes = std::max(tmp_sort, std::max(tmp_scan, tmp_sel)) + 256;
  int slot = 0;
  auto alloc = [&](void** out, size_t bytes) {
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaSuccess) b.p[slot++] = *out;
    return e;
  };
  SA_TRY(alloc((void**)&d_text, words * 8));
  SA_TRY(alloc((void**)&d_key, (size_t)n * 8));
  SA_TRY(alloc((void**)&d_skey, (size_t)n * 8));
  SA_TRY(alloc((void**)&d_val, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_sval, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_sa, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_rank, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_head, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_pos, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_pos2, (size_t)n * 4));
  SA_TRY(alloc((void**)&d_m, 256));
  SA_TRY(alloc((void**)&d_act, (size_t)n));
  SA_TRY(alloc(&d_tmp, tmp_bytes));

  SA_TRY(cudaMemcpyAsync(d_text, text4, words * 8, cudaMemcpyHostToDevice, b.st));
  SA_TRY(cudaEventRecord(b.e0, b.st));

  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  auto grid = [&](uint32_t items) {
    const uint64_t want = ((uint64_t)items + 255) / 256;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms * 8));
  };

  uint32_t m = n, h = 16, n_steps = 0;
  k_sa_key16<<<grid(n), 256, 0, b.st>>>(d_text, n, d_key, d_val, d_pos);
  int end_bit = 64;
  for (;;) {
    size_t tb = tmp_bytes;
    SA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tb, (const uint64_t*)d_key, d_skey, (const uint32_t*)d_val, d_sval,
                                           (int64_t)m, 0, end_bit, b.st));
    k_sa_heads<<<grid(m), 256, 0, b.st>>>(d_skey, d_pos, m, d_head, d_act);
    tb = tmp_bytes;
    // group rank = position of the group's first member + 1; the scan result goes to d_val (free after the sort)
    SA_TRY(cub::DeviceScan::InclusiveScan(d_tmp, tb, (const uint32_t*)d_head, d_val, MaxU32(), (int64_t)m, b.st));
    k_sa_commit<<<grid(m), 256, 0, b.st>>>(d_sval, d_pos, d_val, m, d_sa, d_rank);
    tb = tmp_bytes;
    SA_TRY(cub::DeviceSelect::Flagged(d_tmp, tb, (const uint32_t*)d_pos, (const uint8_t*)d_act, d_pos2, d_m,
                                      (int64_t)m, b.st));
    uint32_t m_next = 0;
    SA_TRY(cudaMemcpyAsync(&m_next, d_m, 4, cudaMemcpyDeviceToHost, b.st));
    SA_TRY(cudaStreamSynchronize(b.st));
    n_steps++;
    std::swap(d_pos, d_pos2);
    m = m_next;
    if (m == 0) break;
    if (h >= n) return tg_fail(TG_ERR_INTERNAL, "suffix array: prefix doubling did not converge");
    k_sa_key_pair<<<grid(m), 256, 0, b.st>>>(d_sa, d_pos, d_rank, m, n, h, bits, d_key, d_val);
    end_bit = (int)(2 * bits);
    h = h > (1u << 30) ? n : h * 2;
  }
  SA_TRY(cudaGetLastError());
  SA_TRY(cudaEventRecord(b.e1, b.st));
  SA_TRY(cudaMemcpyAsync(sa_out, d_sa, (size_t)n * 4, cudaMemcpyDeviceToHost, b.st));
  SA_TRY(cudaStreamSynchronize(b.st));
  if (ms) SA_TRY(cudaEventElapsedTime(ms, b.e0, b.e1));
  if (steps) *steps = n_steps;
  return TG_OK;
}

struct SaHook {
  SaHook() { g_tg_sa_device = &sa_device; }
} g_sa_hook;

}  // namespace

extern "C" tg_status tg_suffix_array_gpu(const uint64_t* text4, uint64_t text_len, int device, uint32_t* sa_out,
                                         float* device_ms, uint32_t* n_steps) {
  try {
    return sa_device(text4, text_len, device, sa_out, device_ms, n_steps);
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  } catch (...) {
    return tg_fail(TG_ERR_INTERNAL, "unknown exception");
  }
}
