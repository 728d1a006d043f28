// tg_paf.cu -- PAF lines written on the GPU (SURVEY 8f row N1 moved onto the device).
//
// The host writers (host_io.cpp) are byte-identical to the reference's PafEntry (src/aln_writer.rs:47-116) but cost 23 ms of
// 16-thread time per 1 M reads inside the file pipeline, next to a parser that wants the same cores, while the GPU needs
// 13 ms for the alignment itself.  Here a batch is aligned with its records left in HBM (tg_align_batch_device), the lines
// are formatted by two kernels (bytes per read -> exclusive scan -> one thread per read writes its lines) and only the
// text crosses PCIe (66 B per read instead of 69 B of records): the host does nothing but copy names up and text out.
// The line layout is restated from host_io.cpp's PAF branch; tests compare the two byte for byte.
// SAM records (aln_to_sam_record / unmapped_sam_record, src/aln_writer.rs:118-253) take the same route with the per-read
// routine of tg_textfmt.h (also built for the host and compared with the oracle's text by the CPU tests): qualities go up
// as well and 319 B of text per read come back, which frees the host's cores for the parser and leaves the file copy
// as the bound of FASTQ -> SAM.
// The scan is a CUB device primitive (library code, like the sorts of tg_sa.cu); the kernels are ours.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "tg_internal.h"
#include "tg_textfmt.h"

namespace {

struct PafParams {
  uint32_t n_reads;
  const uint64_t* offs;        // read offsets: L = offs[r + 1] - offs[r]
  const uint64_t* aln_first;   // device result of tg_align_batch_device (wide records)
  const uint32_t* aln_count;
  const tg_aln* alns;
  const uint32_t* ops;
  const uint8_t* names;        // header lines without '@'
  const uint64_t* name_offs;
  const char* ref_names;       // names of refs() back to back
  const uint32_t* ref_name_offs;
  unsigned long long* line_off;  // [n_reads + 1]: bytes of every read's lines, then their exclusive scan
  char* text;
  uint32_t mapq[6];            // multimapq (src/aln_writer.rs:332-340) of 0 .. 5 records
};

__device__ __forceinline__ uint32_t n_digits(unsigned long long v) {
  uint32_t n = 1;
  while (v >= 10ull) { v /= 10ull; n++; }
  return n;
}
__device__ __forceinline__ char* put_num(char* d, unsigned long long v) {
  const uint32_t n = n_digits(v);
  for (uint32_t i = n; i-- > 0;) { d[i] = (char)('0' + (uint32_t)(v % 10ull)); v /= 10ull; }
  return d + n;
}
// n_match (":" Match runs) and the block length (src/aln_writer.rs:61-88): every aligned or gapped position, one per clip
__device__ __forceinline__ void paf_counts(const uint32_t* w, uint32_t n, unsigned long long& n_match, unsigned long long& n_block) {
  n_match = 0; n_block = 0;
  for (uint32_t k = 0; k < n; k++) {
    const uint32_t kind = w[k] & 7u, run = w[k] >> 3;
    if (kind == TG_OP_MATCH) n_match += run;
    if (kind <= TG_OP_INS) n_block += run;
    else if (kind == TG_OP_XCLIP) n_block += 1;
  }
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k_paf(PafParams p) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    const uint32_t cnt = p.aln_count[r];
    if (!WRITE && cnt == 0) { p.line_off[r] = 0; continue; }  // PAF prints nothing for an unmapped read (src/aligner.rs:77)
    if (cnt == 0) continue;
    const unsigned long long L = p.offs[r + 1] - p.offs[r];
    const unsigned long long nm0 = p.name_offs[r], nm_len = p.name_offs[r + 1] - nm0;
    const uint32_t mq = p.mapq[cnt < 5u ? cnt : 5u];
    const unsigned long long first = p.aln_first[r];
    unsigned long long bytes = 0;
    char* d = WRITE ? p.text + p.line_off[r] : nullptr;
    for (uint32_t i = 0; i < cnt; i++) {
      const tg_aln a = p.alns[first + i];
      unsigned long long n_match, n_block;
      paf_counts(p.ops + a.ops_off, a.ops_len, n_match, n_block);
      const uint32_t rn0 = p.ref_name_offs[a.ref_id], rn_len = p.ref_name_offs[a.ref_id + 1] - rn0;
      if (!WRITE) {
        bytes += nm_len + rn_len + n_digits(L) + n_digits(a.xstart) + n_digits(a.xend) + n_digits(a.ylen) + n_digits(a.ystart) +
                 n_digits(a.yend) + n_digits(n_match) + n_digits(n_block) + n_digits(mq) + 14;  // 12 tabs, strand, newline
      } else {
        for (unsigned long long k = 0; k < nm_len; k++) d[k] = (char)p.names[nm0 + k];
        d += nm_len; *d++ = '\t';
        d = put_num(d, L); *d++ = '\t';
        d = put_num(d, a.xstart); *d++ = '\t';
        d = put_num(d, a.xend); *d++ = '\t';
        *d++ = a.strand ? '+' : '-'; *d++ = '\t';
        for (uint32_t k = 0; k < rn_len; k++) d[k] = p.ref_names[rn0 + k];
        d += rn_len; *d++ = '\t';
        d = put_num(d, a.ylen); *d++ = '\t';
        d = put_num(d, a.ystart); *d++ = '\t';
        d = put_num(d, a.yend); *d++ = '\t';
        d = put_num(d, n_match); *d++ = '\t';
        d = put_num(d, n_block); *d++ = '\t';
        d = put_num(d, mq); *d++ = '\t'; *d++ = '\n';
      }
    }
    if (!WRITE) p.line_off[r] = bytes;
  }
  if (!WRITE && blockIdx.x == 0 && threadIdx.x == 0) p.line_off[p.n_reads] = 0;
}

double tg_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <bool WRITE>
__global__ void __launch_bounds__(128) k_sam(TgTextParams p) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    const unsigned long long bytes = tg_sam_read<WRITE>(p, r);
    if (!WRITE) p.line_off[r] = bytes;
  }
  if (!WRITE && blockIdx.x == 0 && threadIdx.x == 0) p.line_off[p.n_reads] = 0;
}

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  tg_status ensure(size_t bytes) {
    if (bytes <= cap) return TG_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return tg_fail(TG_ERR_CUDA, "cudaMalloc failed (PAF formatter)"); }
    cap = want;
    return TG_OK;
  }
  ~DBuf() { if (p) cudaFree(p); }
};
struct HBuf {
  void* p = nullptr;
  size_t cap = 0;
  tg_status ensure(size_t bytes) {
    if (bytes <= cap) return TG_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return tg_fail(TG_ERR_CUDA, "cudaHostAlloc failed (PAF formatter)"); }
    cap = want;
    return TG_OK;
  }
  ~HBuf() { if (p) cudaFreeHost(p); }
};

}  // namespace

struct tg_paf {
  tg_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  int device = 0, n_sms = 0;
  DBuf d_bases, d_offs, d_names, d_name_offs, d_line_off, d_text[2], d_ref_names, d_ref_name_offs, d_tmp;
  // the text goes home on a stream of its own, so the next batch is aligned while it travels
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t written = nullptr, text_ready[2] = {nullptr, nullptr};
  bool in_flight[2] = {false, false};
  bool timing = false;
  // first batch: the page-locked text buffers (155 ms per 400 MB) are allocated by a helper while the batch is aligned
  std::thread prealloc;
  bool prealloc_started = false;
  bool sam = false;  // SAM records instead of PAF lines: qualities and the annotation's names are needed as well
  DBuf d_quals, d_qual_offs, d_tx_ids, d_tx_id_offs, d_gene_ids, d_gene_id_offs, d_gene_names, d_gene_name_offs, d_tx_gene;
  HBuf h_text[2];
  unsigned long long* h_total = nullptr;  // pinned
  int cur = 0;
  uint32_t mapq[6];
};

#define PAF_CHECK(call)                                                                                \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" void tg_paf_destroy(tg_paf* f);
static tg_status formatter_create(const tg_index_host* ix, tg_ctx* ctx, int device, bool sam, tg_paf** out) {
  TG_GUARD_BEGIN
  if (!ix || !ctx || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  PAF_CHECK(cudaSetDevice(device));
  auto* f = new tg_paf();
  f->ctx = ctx; f->device = device;
  f->stream = (cudaStream_t)tg_ctx_stream(ctx);
  f->timing = getenv("TG_PAF_TIMING") != nullptr;
  if (cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&f->written, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&f->text_ready[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&f->text_ready[1], cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError(); tg_paf_destroy(f); return tg_fail(TG_ERR_CUDA, "cannot create the formatter's copy stream");
  }
  cudaDeviceGetAttribute(&f->n_sms, cudaDevAttrMultiProcessorCount, device);
  std::string names;
  std::vector<uint32_t> offs(1, 0);
  for (uint32_t i = 0; i < tg_index_host_n_refs(ix); i++) {
    uint64_t v[4];
    names += tg_index_host_ref(ix, i, v);
    offs.push_back((uint32_t)names.size());
  }
  tg_status st;
  if ((st = f->d_ref_names.ensure(names.size() + 1)) != TG_OK || (st = f->d_ref_name_offs.ensure(offs.size() * 4)) != TG_OK) { tg_paf_destroy(f); return st; }
  cudaMemcpy(f->d_ref_names.p, names.data(), names.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(f->d_ref_name_offs.p, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice);
  if (cudaMallocHost(&f->h_total, 8) != cudaSuccess) { cudaGetLastError(); f->h_total = nullptr; tg_paf_destroy(f); return tg_fail(TG_ERR_CUDA, "cudaMallocHost failed"); }
  // multimapq (src/aln_writer.rs:332-340): 255 for a unique hit, -10 log10(1 - 1/n) rounded for 2 .. 4, 0 from 5 on
  f->mapq[0] = 255; f->mapq[1] = 255; f->mapq[5] = 0;
  for (int n = 2; n <= 4; n++) f->mapq[n] = (uint32_t)std::lround(-10.0f * std::log10(1.0f - 1.0f / (float)n));
  f->sam = sam;
  if (sam) {
    TgTextTables tb;
    tb.build(ix);
    auto up = [&](DBuf& d, const void* src, size_t bytes) -> tg_status {
      tg_status s2 = d.ensure(bytes + 16);
      if (s2 == TG_OK && bytes && cudaMemcpy(d.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) s2 = tg_fail(TG_ERR_CUDA, "cudaMemcpy failed (SAM formatter tables)");
      return s2;
    };
    if ((st = up(f->d_tx_ids, tb.tx_ids.data(), tb.tx_ids.size())) != TG_OK || (st = up(f->d_tx_id_offs, tb.tx_id_offs.data(), tb.tx_id_offs.size() * 4)) != TG_OK ||
        (st = up(f->d_gene_ids, tb.gene_ids.data(), tb.gene_ids.size())) != TG_OK || (st = up(f->d_gene_id_offs, tb.gene_id_offs.data(), tb.gene_id_offs.size() * 4)) != TG_OK ||
        (st = up(f->d_gene_names, tb.gene_names.data(), tb.gene_names.size())) != TG_OK ||
        (st = up(f->d_gene_name_offs, tb.gene_name_offs.data(), tb.gene_name_offs.size() * 4)) != TG_OK ||
        (st = up(f->d_tx_gene, tb.tx_gene.data(), tb.tx_gene.size() * 4)) != TG_OK) {
      tg_paf_destroy(f); return st;
    }
  }
  *out = f;
  return TG_OK;
  TG_GUARD_END
}

extern "C" {

tg_status tg_paf_create(const tg_index_host* ix, tg_ctx* ctx, int device, tg_paf** out) { return formatter_create(ix, ctx, device, false, out); }
tg_status tg_sam_create(const tg_index_host* ix, tg_ctx* ctx, int device, tg_paf** out) { return formatter_create(ix, ctx, device, true, out); }

void tg_paf_destroy(tg_paf* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaStreamSynchronize(f->stream);
  if (f->prealloc.joinable()) f->prealloc.join();
  if (f->copy_stream) { cudaStreamSynchronize(f->copy_stream); cudaStreamDestroy(f->copy_stream); }
  if (f->written) cudaEventDestroy(f->written);
  for (int i = 0; i < 2; i++) if (f->text_ready[i]) cudaEventDestroy(f->text_ready[i]);
  if (f->h_total) cudaFreeHost(f->h_total);
  delete f;
}

tg_status tg_paf_wait(tg_paf* f, const char* text) {
  TG_GUARD_BEGIN
  if (!f) return tg_fail(TG_ERR_INVALID, "null argument");
  for (int i = 0; i < 2; i++) {
    if (text && (const char*)f->h_text[i].p != text) continue;
    if (!f->in_flight[i]) continue;
    PAF_CHECK(cudaSetDevice(f->device));
    PAF_CHECK(cudaEventSynchronize(f->text_ready[i]));
    f->in_flight[i] = false;
  }
  return TG_OK;
  TG_GUARD_END
}

tg_status tg_paf_align_batch(tg_paf* f, const tg_read_batch* b, const char** text, size_t* text_len, tg_result* counters) {
  tg_status st = tg_paf_align_batch_async(f, b, text, text_len, counters);
  return st == TG_OK && text ? tg_paf_wait(f, *text) : st;
}

tg_status tg_paf_align_batch_async(tg_paf* f, const tg_read_batch* b, const char** text, size_t* text_len, tg_result* counters) {
  TG_GUARD_BEGIN
  if (!f || !b || !text || !text_len) return tg_fail(TG_ERR_INVALID, "null argument");
  *text = nullptr; *text_len = 0;
  const uint32_t n = b->n_reads;
  if (n == 0) return TG_OK;
  PAF_CHECK(cudaSetDevice(f->device));
  const double t0 = f->timing ? tg_now_ms() : 0;
  const uint64_t nb = b->offs[n] - b->offs[0], nn = b->name_offs[n];
  if (b->offs[0] != 0 || b->name_offs[0] != 0) return tg_fail(TG_ERR_INVALID, "batch offsets must start at 0");
  if (f->sam && (!b->quals || !b->qual_offs || b->qual_offs[0] != 0)) return tg_fail(TG_ERR_INVALID, "SAM output needs qualities (offsets starting at 0)");
  const uint64_t nq = f->sam ? b->qual_offs[n] : 0;
  tg_status st;
  if (f->sam && ((st = f->d_quals.ensure(nq + 64)) != TG_OK || (st = f->d_qual_offs.ensure((size_t)(n + 1) * 8)) != TG_OK)) return st;
  if ((st = f->d_bases.ensure(nb + 64)) != TG_OK || (st = f->d_offs.ensure((size_t)(n + 1) * 8)) != TG_OK ||
      (st = f->d_names.ensure(nn + 64)) != TG_OK || (st = f->d_name_offs.ensure((size_t)(n + 1) * 8)) != TG_OK ||
      (st = f->d_line_off.ensure((size_t)(n + 2) * 8)) != TG_OK)
    return st;
  // read lengths on the host: the device entry point wants the longest read (one pass over the offsets)
  uint32_t maxL = 1;
  for (uint32_t r = 0; r < n; r++) {
    const uint64_t d = b->offs[r + 1] - b->offs[r];
    if (b->offs[r + 1] < b->offs[r] || d > TG_MAX_READ_LEN) return tg_fail(TG_ERR_INVALID, "read offsets must be non-decreasing and reads at most TG_MAX_READ_LEN long");
    maxL = d > maxL ? (uint32_t)d : maxL;
  }
  PAF_CHECK(cudaMemcpyAsync(f->d_offs.p, b->offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, f->stream));
  if (nb) PAF_CHECK(cudaMemcpyAsync(f->d_bases.p, b->bases, nb, cudaMemcpyHostToDevice, f->stream));
  PAF_CHECK(cudaMemcpyAsync(f->d_name_offs.p, b->name_offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, f->stream));
  if (nn) PAF_CHECK(cudaMemcpyAsync(f->d_names.p, b->names, nn, cudaMemcpyHostToDevice, f->stream));
  if (f->sam) {
    PAF_CHECK(cudaMemcpyAsync(f->d_qual_offs.p, b->qual_offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, f->stream));
    if (nq) PAF_CHECK(cudaMemcpyAsync(f->d_quals.p, b->quals, nq, cudaMemcpyHostToDevice, f->stream));
  }
  if (!f->prealloc_started) {
    f->prealloc_started = true;
    const size_t est = f->sam ? (size_t)(nb + nq + nn) + (size_t)n * 170 : (size_t)nn + (size_t)n * 64;
    f->prealloc = std::thread([f, est]() {
      if (cudaSetDevice(f->device) != cudaSuccess) return;
      f->h_text[0].ensure(est);
      f->h_text[1].ensure(est);
    });
  }
  struct JoinGuard { std::thread& t; ~JoinGuard() { if (t.joinable()) t.join(); } } join_guard{f->prealloc};
  tg_result res;
  if ((st = tg_align_batch_device(f->ctx, (const uint8_t*)f->d_bases.p, (const uint64_t*)f->d_offs.p, n, nb, maxL, &res)) != TG_OK) return st;
  if (counters) *counters = res;
  PafParams p;
  p.n_reads = n; p.offs = (const uint64_t*)f->d_offs.p;
  p.aln_first = res.read_aln_first; p.aln_count = res.read_aln_count; p.alns = res.alns; p.ops = res.ops;
  p.names = (const uint8_t*)f->d_names.p; p.name_offs = (const uint64_t*)f->d_name_offs.p;
  p.ref_names = (const char*)f->d_ref_names.p; p.ref_name_offs = (const uint32_t*)f->d_ref_name_offs.p;
  p.line_off = (unsigned long long*)f->d_line_off.p; p.text = nullptr;
  memcpy(p.mapq, f->mapq, sizeof(p.mapq));
  TgTextParams q;
  memset(&q, 0, sizeof(q));
  if (f->sam) {
    q.n_reads = n; q.bases = (const uint8_t*)f->d_bases.p; q.offs = p.offs;
    q.aln_first = p.aln_first; q.aln_count = p.aln_count; q.alns = p.alns; q.ops = p.ops;
    q.names = p.names; q.name_offs = p.name_offs;
    q.quals = (const uint8_t*)f->d_quals.p; q.qual_offs = (const uint64_t*)f->d_qual_offs.p;
    q.ref_names = p.ref_names; q.ref_name_offs = p.ref_name_offs;
    q.tx_ids = (const char*)f->d_tx_ids.p; q.tx_id_offs = (const uint32_t*)f->d_tx_id_offs.p;
    q.gene_ids = (const char*)f->d_gene_ids.p; q.gene_id_offs = (const uint32_t*)f->d_gene_id_offs.p;
    q.gene_names = (const char*)f->d_gene_names.p; q.gene_name_offs = (const uint32_t*)f->d_gene_name_offs.p;
    q.tx_gene = (const uint32_t*)f->d_tx_gene.p;
    q.line_off = p.line_off; q.text = nullptr;
    memcpy(q.mapq, f->mapq, sizeof(q.mapq));
  }
  const int blocks = (int)std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)f->n_sms * 16);
  if (f->sam) k_sam<false><<<blocks, 128, 0, f->stream>>>(q);
  else k_paf<false><<<blocks, 128, 0, f->stream>>>(p);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, p.line_off, p.line_off, (int)(n + 1), f->stream);
  if ((st = f->d_tmp.ensure(tmp_bytes + 16)) != TG_OK) return st;
  PAF_CHECK(cub::DeviceScan::ExclusiveSum(f->d_tmp.p, tmp_bytes, p.line_off, p.line_off, (int)(n + 1), f->stream));
  PAF_CHECK(cudaMemcpyAsync(f->h_total, p.line_off + n, 8, cudaMemcpyDeviceToHost, f->stream));
  PAF_CHECK(cudaStreamSynchronize(f->stream));
  const double t1 = f->timing ? tg_now_ms() : 0;
  const size_t total = (size_t)*f->h_total;
  if (f->prealloc.joinable()) f->prealloc.join();
  const int cur = f->cur;
  f->cur ^= 1;
  HBuf& H = f->h_text[cur];
  DBuf& D = f->d_text[cur];
  // this buffer pair carried the batch before the previous one: its copy has long ended, but make sure before reuse
  if (f->in_flight[cur]) { PAF_CHECK(cudaEventSynchronize(f->text_ready[cur])); f->in_flight[cur] = false; }
  if ((st = D.ensure(total + 64)) != TG_OK || (st = H.ensure(total + 64)) != TG_OK) return st;
  const double t2 = f->timing ? tg_now_ms() : 0;
  if (total) {
    p.text = (char*)D.p;
    q.text = p.text;
    if (f->sam) k_sam<true><<<blocks, 128, 0, f->stream>>>(q);
    else k_paf<true><<<blocks, 128, 0, f->stream>>>(p);
    PAF_CHECK(cudaGetLastError());
    PAF_CHECK(cudaEventRecord(f->written, f->stream));
    PAF_CHECK(cudaStreamWaitEvent(f->copy_stream, f->written, 0));
    PAF_CHECK(cudaMemcpyAsync(H.p, D.p, total, cudaMemcpyDeviceToHost, f->copy_stream));
    PAF_CHECK(cudaEventRecord(f->text_ready[cur], f->copy_stream));
    f->in_flight[cur] = true;
    // the records, reads and names of this batch are free for the next one once the kernel has run
    PAF_CHECK(cudaStreamSynchronize(f->stream));
  }
  if (f->timing) {
    const double t3 = tg_now_ms();
    cudaStreamSynchronize(f->copy_stream);
    fprintf(stderr, "[tg_paf] %u reads, %zu B of text: copies up + alignment + line lengths %.2f ms, buffers %.2f ms, text kernel %.2f ms, text copy (rest) %.2f ms\n",
            n, total, t1 - t0, t2 - t1, t3 - t2, tg_now_ms() - t3);
  }
  *text = (const char*)H.p;
  *text_len = total;
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"
