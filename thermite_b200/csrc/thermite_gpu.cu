// thermite_gpu.cu -- CUDA kernels (sm_100a) and the device half of the C ABI of libthermite_gpu.so.
//
// Kernels (reference file:line of what each replaces; design in DESIGN.md section 4):
//   k_kmer_count / k_kmer_insert      open-addressing k-mer table over the suffix array (ctx create)
//   k_pack_reads, k_probe_light, k_probe_heavy, k_seed_select
//                                     Index::all_smems (src/index.rs:228-255): thread per (read, offset) probes in
//                                     three waves, thread per read SMEM selection.  HBM random-access bound
//   k_round_{init,plan,prep,hist,binscan,scatter,post,scan,final}
//                                     align_read / align_seed_hit (src/aligner.rs:123-449, src/txome.rs:82-160) as the
//                                     speculative round pipeline of tg_rounds.h: thread per read / per hit
//   k_round_dpt<1..11>                SwgExtend::extend / trace (src/swg.rs:31-207): thread per extension, band in
//                                     registers (tg_dpt.h), one kernel per group of band classes.  INT-pipe bound
//   k_round_dp<R>, k_swg_batch<R>     the same on the warp-cooperative wavefront (tg_core.h): long reads, very wide
//                                     bands, raw-byte pairs of tg_swg_extend_batch
//   k_extend<R>                       one warp walks one read start to finish (tg_align_read): reads the round
//                                     pipeline cannot hold, and tg_ctx_set_round_pipeline(ctx, 0)
//   k_swg_prepare / k_swg_collect     tg_swg_extend_batch on the thread kernels;  k_random_gather: HBM yardstick
// No CPU fallback exists: every entry point below needs a device.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "tg_rounds.h"
#include "tg_dpt.h"

#define TG_FULL 0xffffffffu
#define TG_WARPS_PER_CTA 4

namespace {

struct DevWarp {
  static constexpr int LANES = 32;
  TG_HD int lane() const {
#ifdef __CUDA_ARCH__
    return (int)(threadIdx.x & 31);
#else
    return 0;
#endif
  }
  TG_HD int shfl_up(int v, int d) {
#ifdef __CUDA_ARCH__
    return __shfl_up_sync(TG_FULL, v, d);
#else
    return v;
#endif
  }
  TG_HD int shfl(int v, int src) {
#ifdef __CUDA_ARCH__
    return __shfl_sync(TG_FULL, v, src);
#else
    return v;
#endif
  }
  TG_HD unsigned long long shfl64(unsigned long long v, int src) {
#ifdef __CUDA_ARCH__
    return __shfl_sync(TG_FULL, v, src);
#else
    return v;
#endif
  }
  TG_HD bool any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(TG_FULL, p) != 0;
#else
    return p;
#endif
  }
  TG_HD uint32_t ballot(bool p) {
#ifdef __CUDA_ARCH__
    return __ballot_sync(TG_FULL, p);
#else
    return p ? 1u : 0u;
#endif
  }
  TG_HD int reduce_max_i32(int v) {
#ifdef __CUDA_ARCH__
    return __reduce_max_sync(TG_FULL, v);
#else
    return v;
#endif
  }
  TG_HD uint32_t reduce_min_u32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __reduce_min_sync(TG_FULL, v);
#else
    return v;
#endif
  }
  TG_HD unsigned long long sum64(unsigned long long v) {
#ifdef __CUDA_ARCH__
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(TG_FULL, v, d);
#endif
    return v;
  }
  TG_HD void sync() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
  }
  TG_HD void sync_global() {
#ifdef __CUDA_ARCH__
    __threadfence_block();
    __syncwarp();
#endif
  }
  TG_HD unsigned long long atomic_add(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    unsigned long long o = *p; *p += v; return o;
#endif
  }
  TG_HD void atomic_or(int* p, int v) {
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
  }
};

// Pool allocation from thread-per-item kernels: the threads of a warp that reach the call together combine their
// requests into ONE atomicAdd (a million same-address atomics per kernel otherwise serialise in L2).
__device__ __forceinline__ unsigned long long warp_agg_add(unsigned long long* ctr, unsigned long long n) {
  namespace cg = cooperative_groups;
  cg::coalesced_group g = cg::coalesced_threads();
  const unsigned long long prefix = cg::exclusive_scan(g, n);
  unsigned long long base = 0;
  if (g.thread_rank() == g.size() - 1) base = atomicAdd(ctr, prefix + n);
  base = g.shfl(base, g.size() - 1);
  return base + prefix;
}

// One thread = one "warp" of a single lane: lets the thread-per-read round kernels reuse the templated code.
struct DevThread {
  static constexpr int LANES = 1;
  TG_HD int lane() const { return 0; }
  TG_HD int shfl_up(int v, int) { return v; }
  TG_HD int shfl(int v, int) { return v; }
  TG_HD unsigned long long shfl64(unsigned long long v, int) { return v; }
  TG_HD bool any(bool p) { return p; }
  TG_HD uint32_t ballot(bool p) { return p ? 1u : 0u; }
  TG_HD int reduce_max_i32(int v) { return v; }
  TG_HD uint32_t reduce_min_u32(uint32_t v) { return v; }
  TG_HD unsigned long long sum64(unsigned long long v) { return v; }
  TG_HD void sync() {}
  TG_HD void sync_global() {}
  TG_HD unsigned long long atomic_add(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    return warp_agg_add(p, v);
#else
    unsigned long long o = *p; *p += v; return o;
#endif
  }
  TG_HD void atomic_or(int* p, int v) {
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
  }
};

#define TG_DPT_CBINS 512  // per class: 32 band-width bins x 16 column-count bins
#define TG_DPT_NBINS (TG_DPT_NCLS * TG_DPT_CBINS)
struct DevCounters {
  unsigned long long seed_used, n_smems, alns_used, ops_used, cells, n_ext, hits, work_seed, work_ext, swg_ops_used,
      work_swg, kmer_groups;
  int flags;
  int pad;
  unsigned long long phase[16];
  // round pipeline
  unsigned long long n_complex, work_complex;
  unsigned long long probe_n[3][2];  // per probe wave: queued probes, deferred (multi-occurrence) probes
  unsigned long long items_used, hops_used, fin_used;
  unsigned long long n_late;  // reads finalised by the LAST pass (their first/count reach the host as a fix-up list)
  unsigned long long round_end[TG_MAX_ROUNDS];  // items_used after round r
  unsigned long long round_active[TG_MAX_ROUNDS];  // reads still unfinished after round r
  unsigned long long round_tasks[TG_MAX_ROUNDS], round_ops[TG_MAX_ROUNDS], round_work[TG_MAX_ROUNDS], round_work2[TG_MAX_ROUNDS][TG_DPT_NCLS];
  // task sorting for the thread-per-extension kernel: bins = class * TG_DPT_CBINS + column bucket
  uint32_t bin_count[TG_DPT_NBINS], bin_cursor[TG_DPT_NBINS];
  uint32_t cls_start[TG_DPT_NCLS + 1], cls_end[TG_DPT_NCLS + 1], cls_chunk0[TG_DPT_NCLS + 1];
  uint32_t warp_tasks;  // sorted[0 .. warp_tasks): tasks for the warp-cooperative kernel
  uint32_t round_cls[TG_MAX_ROUNDS][TG_DPT_NCLS];  // debug: tasks per band class
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(TG_FULL, v, d);
  return v;
}
__device__ __forceinline__ uint32_t next_work(unsigned long long* ctr) {
  unsigned long long r = 0;
  if ((threadIdx.x & 31) == 0) r = atomicAdd(ctr, 1ull);
  return (uint32_t)__shfl_sync(TG_FULL, r, 0);
}
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ---------------------------------------------------------------------------------------------------
// k-mer table build
// ---------------------------------------------------------------------------------------------------
__global__ void k_kmer_count(const uint64_t* __restrict__ text4, uint64_t T, const uint32_t* __restrict__ sa, uint32_t k,
                             DevCounters* ctr) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (; r < T; r += stride) {
    uint64_t w0, w1;
    if (tg_kmer_group_start(text4, T, sa, r, k, w0, w1)) local++;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&ctr->kmer_groups, local);
}

__global__ void k_kmer_insert(const uint64_t* __restrict__ text4, uint64_t T, const uint32_t* __restrict__ sa, uint32_t k,
                              TgSlot* slots, uint64_t slot_mask) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; r < T; r += stride) {
    uint64_t w0, w1;
    if (!tg_kmer_group_start(text4, T, sa, r, k, w0, w1)) continue;
    uint32_t cnt = tg_kmer_group_count(text4, T, sa, r, k, w0, w1);
    uint64_t h = tg_hash_kmer(w0, w1);
    uint32_t tag = tg_tag_of(h);
    uint64_t idx = h & slot_mask;
    for (;;) {
      uint32_t old = atomicCAS(&slots[idx].tag, 0u, tag);
      if (old == 0u) {
        slots[idx].lo = cnt == 1 ? sa[r] : (uint32_t)r;
        slots[idx].count = cnt;
        break;
      }
      idx = (idx + 1) & slot_mask;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// seeding
// ---------------------------------------------------------------------------------------------------
struct SeedParams {
  const uint8_t* bases;
  const uint64_t* offs;
  uint32_t n_reads, k, max_len, max_q, rp_words;  // max_q = max_len - k + 1 probe offsets per read
  const TgSlot* slots;
  uint64_t slot_mask;
  const uint64_t* text4;
  const uint32_t* sa;
  uint64_t* rp;        // [n_reads][rp_words] packed reads (also used by the extension stage)
  TgSeedHit* hits;     // [n_reads][max_q]
  uint64_t* queue;     // probes of the current wave: read << 16 | offset
  uint64_t* queue2;    // the ones whose k-mer occurs more than once
  int wave;            // 0: offset 0 of every read; 1: the other offsets, skipped when the whole read matched at 0
  TgSeedOut out;
  DevCounters* ctr;
};

// ASCII -> 4-bit codes, one thread per packed word
__global__ void __launch_bounds__(256) k_pack_reads(SeedParams p) {
  const unsigned long long total = (unsigned long long)p.n_reads * p.rp_words;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(t / p.rp_words), wi = (uint32_t)(t % p.rp_words);
    const uint64_t off = p.offs[r];
    const uint32_t L = (uint32_t)(p.offs[r + 1] - off);
    p.rp[t] = wi < L / 16 + 3 ? tg_pack_word(p.bases, off, L, wi) : 0xFFFFFFFFFFFFFFFFull;
  }
}

// E(q) in the three waves of tg_core.h (TG_PROBE_STRIDE), two kernels per wave:
//   light  one thread per candidate (read, offset): skip rules, hash, slot, verification of a single-occurrence k-mer;
//          k-mers that occur more than once are queued
//   heavy  one thread per queued probe: the suffix-array searches, 32 real searches per warp instead of a few live
//          lanes holding a warp of finished ones
__device__ __forceinline__ void probe_push(unsigned long long* ctr, uint64_t* queue, uint32_t r, uint32_t q) {
  queue[warp_agg_add(ctr, 1ull)] = ((uint64_t)r << 16) | q;
}
// wave 2, step 1: one thread per (read, bracket of TG_PROBE_STRIDE offsets).  A bracket whose two samples ended at the same
// place pins every E in between (closed: nothing to do, the usual case); the interior offsets of an open bracket are queued.
// (The first version gave every interior offset of every read its own thread, 7 of 8 of which only re-read the two samples
// and left: 284 M threads and 2.1 kB of table traffic per read for a 4 M-read batch.)
__global__ void __launch_bounds__(256) k_probe_brackets(SeedParams p) {
  const uint32_t nb = (p.max_q + TG_PROBE_STRIDE - 1) / TG_PROBE_STRIDE;
  const unsigned long long total = (unsigned long long)p.n_reads * nb;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(t / nb), j = (uint32_t)(t % nb);
    const uint32_t L = (uint32_t)(p.offs[r + 1] - p.offs[r]);
    if (L < p.k) continue;
    const uint32_t q_last = L - p.k, qa = j * TG_PROBE_STRIDE;
    if (qa >= q_last) continue;
    const TgSeedHit* row = p.hits + (size_t)r * p.max_q;
    if (row[0].e == L) continue;  // read[0..L) occurs: E(q) = L for every q, no other SMEM
    const uint32_t qb = qa + TG_PROBE_STRIDE < q_last ? qa + TG_PROBE_STRIDE : q_last;
    if (qb - qa < 2) continue;
    const uint32_t ea = row[qa].e, eb = row[qb].e;
    if (ea == eb && ea != 0) continue;  // tg_probe_bracketed
    const unsigned long long base = warp_agg_add(&p.ctr->probe_n[2][0], (unsigned long long)(qb - qa - 1));
    for (uint32_t q = qa + 1; q < qb; q++) p.queue[base + (q - qa - 1)] = ((uint64_t)r << 16) | q;
  }
}
__global__ void __launch_bounds__(256) k_probe_light(SeedParams p) {
  const uint32_t qn = p.wave == 0 ? 1u : (p.max_q + TG_PROBE_STRIDE - 1) / TG_PROBE_STRIDE;
  const unsigned long long total = p.wave == 2 ? p.ctr->probe_n[2][0] : (unsigned long long)p.n_reads * qn;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t r, q = 0, L;
    TgSeedHit* row;
    if (p.wave == 2) {  // step 2: one thread per queued offset
      const uint64_t e = p.queue[t];
      r = (uint32_t)(e >> 16); q = (uint32_t)(e & 0xFFFFu);
      L = (uint32_t)(p.offs[r + 1] - p.offs[r]);
      row = p.hits + (size_t)r * p.max_q;
    } else {
      r = (uint32_t)(t / qn);
      const uint32_t j = (uint32_t)(t % qn);
      L = (uint32_t)(p.offs[r + 1] - p.offs[r]);
      if (L < p.k) continue;
      const uint32_t q_last = L - p.k;
      row = p.hits + (size_t)r * p.max_q;
      if (p.wave == 1) {
        if (row[0].e == L) continue;  // read[0..L) occurs: E(q) = L for every q, no other SMEM
        q = tg_probe_sample(j, q_last);
        if (q == 0xFFFFFFFFu) continue;
      }
    }
    TgSeedHit h;
    if (tg_seed_offset<true>(p.rp + (size_t)r * p.rp_words, L, q, p.k, p.slots, p.slot_mask, p.text4, p.sa, h))
      row[q] = h;
    else
      probe_push(&p.ctr->probe_n[p.wave][1], p.queue2, r, q);
  }
}
__global__ void __launch_bounds__(256) k_probe_heavy(SeedParams p) {
  const unsigned long long total = p.ctr->probe_n[p.wave][1];
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t e = p.queue2[t];
    const uint32_t r = (uint32_t)(e >> 16), q = (uint32_t)(e & 0xFFFFu);
    const uint32_t L = (uint32_t)(p.offs[r + 1] - p.offs[r]);
    TgSeedHit h;
    tg_seed_offset<false>(p.rp + (size_t)r * p.rp_words, L, q, p.k, p.slots, p.slot_mask, p.text4, p.sa, h);
    p.hits[(size_t)r * p.max_q + q] = h;
  }
}

// Index::all_smems per read from E[]: one thread per read
__global__ void __launch_bounds__(128) k_seed_select(SeedParams p) {
  DevThread w;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    const uint32_t L = (uint32_t)(p.offs[r + 1] - p.offs[r]);
    tg_seed_select_read<DevThread>(w, p.hits + (size_t)r * p.max_q, L, p.k, p.out, r);
  }
}

// ---------------------------------------------------------------------------------------------------
// extension
// ---------------------------------------------------------------------------------------------------
struct ExtParams {
  TgAlignParams P;
  const uint8_t* bases;
  const uint64_t* offs;
  uint32_t n_reads, max_len, max_cols, trace_bytes, ops_cap;
  const tg_seed* seeds;
  const uint64_t* read_seed_first;
  const uint32_t* read_seed_count;
  TgCand* cands;      // [n_warps][TG_MAX_ALNS_PER_READ]
  uint32_t* arena;    // [n_warps][arena_cap]
  uint32_t arena_cap;
  uint16_t* order;    // [n_warps][2 * TG_MAX_ALNS_PER_READ]
  int bound_stop;
  const uint32_t* read_list;  // when set: only these reads (the "complex" ones the round pipeline handed over)
  TgAlignOut out;
  DevCounters* ctr;
};
struct ExtSmemLayout {
  size_t rd, xs, ys, trace, ops, stack, rp, total;
};
__host__ __device__ inline ExtSmemLayout ext_smem_layout(uint32_t maxL, uint32_t max_cols, uint32_t trace_bytes, uint32_t ops_cap) {
  ExtSmemLayout l;
  size_t o = 0;
  l.rd = o; o += align16(maxL + 16);
  l.xs = o; o += align16(maxL + 16);
  l.ys = o; o += align16(max_cols + 16);
  l.trace = o; o += align16(trace_bytes);
  l.ops = o; o += align16((size_t)ops_cap * 4) * 4;
  l.stack = o; o += align16(TG_TREE_STACK * 4);
  l.rp = o; o += align16((maxL / 16 + 4) * 8);
  l.total = o;
  return l;
}

template <int RMAX>
__global__ void __launch_bounds__(TG_WARPS_PER_CTA * 32) k_extend(ExtParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const ExtSmemLayout lay = ext_smem_layout(p.max_len, p.max_cols, p.trace_bytes, p.ops_cap);
  uint8_t* base = smem + (threadIdx.x >> 5) * lay.total;
  TgWarpMem m;
  m.rd = base + lay.rd; m.xs = base + lay.xs; m.ys = base + lay.ys; m.trace = base + lay.trace;
  const size_t ob = align16((size_t)p.ops_cap * 4);
  m.opsA = (uint32_t*)(base + lay.ops); m.opsB = (uint32_t*)(base + lay.ops + ob);
  m.opsC = (uint32_t*)(base + lay.ops + 2 * ob); m.opsT = (uint32_t*)(base + lay.ops + 3 * ob);
  m.stack = (int32_t*)(base + lay.stack);
  m.rp = (uint64_t*)(base + lay.rp);
  m.ops_cap = p.ops_cap;
  m.bound_stop = p.bound_stop != 0;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  TgWarpScratch sc{p.cands + (size_t)gw * TG_MAX_ALNS_PER_READ, p.arena + (size_t)gw * p.arena_cap, p.arena_cap,
                   p.order + (size_t)gw * 2 * TG_MAX_ALNS_PER_READ};
  DevWarp w;
  TgCounters ctr{};
  for (;;) {
    uint32_t r;
    if (p.read_list) {
      uint32_t wi = next_work(&p.ctr->work_complex);
      if ((unsigned long long)wi >= p.ctr->n_complex) break;
      r = p.read_list[wi];
    } else {
      r = next_work(&p.ctr->work_ext);
      if (r >= p.n_reads) break;
    }
    uint64_t off = p.offs[r];
    uint32_t L = (uint32_t)(p.offs[r + 1] - off);
    tg_align_read<DevWarp, RMAX>(w, m, p.P, p.bases, off, L, p.seeds + p.read_seed_first[r], p.read_seed_count[r], sc, p.out, r, ctr);
  }
  unsigned long long c = warp_sum(ctr.cells), e = warp_sum(ctr.n_ext), h = warp_sum(ctr.hits);
  if ((threadIdx.x & 31) == 0) {
    if (c) atomicAdd(&p.ctr->cells, c);
    if (e) atomicAdd(&p.ctr->n_ext, e);
    if (h) atomicAdd(&p.ctr->hits, h);
#ifdef TG_PROFILE_PHASES
    for (int k = 0; k < TG_NPHASE; k++) atomicAdd(&p.ctr->phase[k], ctr.ph[k]);
#endif
  }
}

__host__ __device__ inline size_t swg_smem_per_warp(uint32_t max_xlen, uint32_t max_cols, uint32_t trace_bytes, uint32_t ops_words);

// ---------------------------------------------------------------------------------------------------
// round pipeline (tg_rounds.h): thread-per-read / thread-per-hit control kernels + a warp-per-task extension kernel
// ---------------------------------------------------------------------------------------------------
struct RoundParams {
  TgAlignParams P;
  const uint8_t* bases;
  const uint64_t* offs;
  uint32_t n_reads, max_len, rp_words;
  const tg_seed* seeds;
  const uint64_t* read_seed_first;
  const uint32_t* read_seed_count;
  TgReadState* st;
  uint64_t* rp;       // [n_reads][rp_words] packed reads
  // items (append-only within the batch): one per evaluated (read, hit)
  TgHit* hits;
  TgItemRes* ires;
  TgCand* cands;
  unsigned long long item_cap;
  TgHopsPool hp;      // operations of the kept items
  uint32_t* fin;      // scratch of the finaliser for reads with many accepted alignments
  unsigned long long fin_cap;
  // per round
  TgTask* tasks;
  unsigned long long task_cap;
  uint32_t* ops_pool;
  unsigned long long ops_cap;
  uint32_t* complex_list;
  uint32_t round;
  int early;               // k_round_final: first pass (finished reads only)
  ulonglong2* late;        // last pass: {first, read | count << 32} of every read it finalises (null: not recorded)
  uint32_t* sorted;        // task indices of the round, grouped by class and (descending) column count
  uint32_t* perm;          // items of the round ordered by text locus (null: item order); see k_round_ikey
  uint32_t* ikey;          // [item] locus bucket
  uint32_t* ibins;         // [2][TG_IB_N] bucket counts, cursors
  uint32_t ib_shift;
  int all_warp;            // 1: every task of the round goes to the warp-cooperative kernel (tiny batches: no per-class launches)
  uint32_t* dpt_trace;     // thread kernels: per class group, [warp][col][word][lane]
  size_t dpt_trace_off[TG_DPT_NCLS], dpt_trace_words[TG_DPT_NCLS];  // region of each band class, words per warp
  int dpt_one, dpt_k128;   // 1 and 128 (see TgDptMem)
  uint32_t max_xlen, max_cols, trace_bytes, ops_words;
  int bound_stop;
  TgAlignOut out;
  DevCounters* ctr;
};

__device__ __forceinline__ void mark_complex(const RoundParams& p, uint32_t r, int reason) {
  p.st[r].status = TG_RS_COMPLEX;
  atomicAdd(&p.ctr->phase[12 + reason], 1ull);  // debug statistics: why reads leave the round path
  unsigned long long i = atomicAdd(&p.ctr->n_complex, 1ull);
  p.complex_list[i] = r;
}
__device__ __forceinline__ void round_item_range(const RoundParams& p, unsigned long long& lo, unsigned long long& hi) {
  lo = p.round ? p.ctr->round_end[p.round - 1] : 0ull;
  hi = p.ctr->items_used;
  if (hi > p.item_cap) hi = p.item_cap;
  if (lo > hi) lo = hi;
}

__global__ void __launch_bounds__(128) k_round_init(RoundParams p) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    const uint64_t off = p.offs[r];
    const uint32_t L = (uint32_t)(p.offs[r + 1] - off);
    const uint32_t ns = p.read_seed_count[r];
    const tg_seed* sd = p.seeds + p.read_seed_first[r];
    TgReadState st;
    const bool ok = tg_read_state_init(st, L, p.P.opts, ns, sd);
    p.st[r] = st;
    if (!ok) mark_complex(p, r, 0);
  }
}

// every active read submits its next batch of hits
__global__ void __launch_bounds__(128) k_round_plan(RoundParams p) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    if (p.st[r].status != TG_RS_ACTIVE || p.st[r].planned) continue;  // (scan already submitted the batch)
    TgReadState st = p.st[r];
    uint32_t b = tg_plan_batch(st, p.round);
    const unsigned long long base = warp_agg_add(&p.ctr->items_used, (unsigned long long)b);
    if (base + b > p.item_cap) {
      atomicOr(&p.ctr->flags, TG_FLAG_ITEM_POOL);
      b = 0;
    }
    st.batch_first = (uint32_t)base; st.batch_n = b;
    for (uint32_t i = 0; i < b; i++) {
      TgItemRes& ir = p.ires[base + i];
      ir.read = r; ir.hit = st.next_hit + i; ir.flags = 0; ir.prev_acc = TG_NONE; ir.state = tg_pack_state(st.bw, st.x_drop);
    }
    p.st[r] = st;
  }
}

// ---- item order: the thread-per-hit kernels (prep, post) walk data-dependent loops over the transcripts and exons at the
// hit's locus, so a warp whose 32 hits sit at 32 unrelated loci runs 32 different control flows (3.8 active lanes per
// instruction, profiles/r1_ncu_prep.csv).  Hits are therefore handed to threads in the order of their text position
// (2^TG_IB_BITS buckets of 2^ib_shift symbols, counting sort: count -> scan -> scatter): neighbours in a warp then see the same gene,
// the same transcripts and the same cache lines (prep, round 0, 1 M hits: 436 M -> 190 M warp instructions, 4.7 -> 10.7
// active lanes, 1.05 -> 0.63 ms).  Only prep's thread -> item mapping changes; items stay where they are.
#ifndef TG_IB_BITS
#define TG_IB_BITS 14
#endif
#define TG_IB_N (1 << TG_IB_BITS)
__global__ void __launch_bounds__(256) k_round_ikey(RoundParams p) {
  unsigned long long lo, hi;
  round_item_range(p, lo, hi);
  for (unsigned long long it = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < hi; it += (unsigned long long)gridDim.x * blockDim.x) {
    const TgItemRes& ir = p.ires[it];
    const tg_seed* seeds = p.seeds + p.read_seed_first[ir.read];
    uint32_t si, rk;
    tg_hit_locate(seeds, p.read_seed_count[ir.read], ir.hit, si, rk);
    const tg_seed sd = seeds[si];
    const uint32_t ref_idx = sd.direct ? sd.sa_lo : TG_LDG(p.P.ix.sa + sd.sa_lo + rk);
    uint32_t b = ref_idx >> p.ib_shift;
    if (b >= TG_IB_N) b = TG_IB_N - 1;
    p.ikey[it] = b;
    atomicAdd(&p.ibins[b], 1u);
  }
}
// one block: exclusive scan of the bucket counts into the cursors; clears the counts for the next round
__global__ void __launch_bounds__(1024) k_round_iscan(RoundParams p) {
  constexpr int PER = TG_IB_N / 1024;
  __shared__ uint32_t wsum[32];
  uint32_t sum = 0;
  const int b0 = threadIdx.x * PER;
  for (int k = 0; k < PER; k++) sum += p.ibins[b0 + k];
  uint32_t inc = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t x = wsum[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    wsum[lane] = xi - x;
  }
  __syncthreads();
  uint32_t acc = wsum[wid] + inc - sum;
  for (int k = 0; k < PER; k++) {
    const uint32_t v = p.ibins[b0 + k];
    p.ibins[b0 + k] = 0;
    p.ibins[TG_IB_N + b0 + k] = acc;
    acc += v;
  }
}
__global__ void __launch_bounds__(256) k_round_iscatter(RoundParams p) {
  unsigned long long lo, hi;
  round_item_range(p, lo, hi);
  for (unsigned long long it = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < hi; it += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t pos = atomicAdd(&p.ibins[TG_IB_N + p.ikey[it]], 1u);
    p.perm[pos] = (uint32_t)it;
  }
}

__global__ void __launch_bounds__(128) k_round_prep(RoundParams p) {
  DevThread w;
  unsigned long long lo, hi;
  round_item_range(p, lo, hi);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi - lo; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long it = p.perm ? p.perm[i] : lo + i;
    TgItemRes& ir = p.ires[it];
    const uint32_t r = ir.read;
    const bool ok = tg_item_prep<DevThread>(w, p.P, p.rp + (size_t)r * p.rp_words, p.st[r], ir.state, p.seeds + p.read_seed_first[r],
                                            p.read_seed_count[r], r, ir.hit, p.hits[it], p.tasks, &p.ctr->round_tasks[p.round],
                                            p.task_cap, &p.ctr->flags);
    if (!ok) ir.flags = TG_IF_FAIL;
  }
}

template <int RMAX>
__global__ void __launch_bounds__(TG_WARPS_PER_CTA * 32) k_round_dp(RoundParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* base = smem + (threadIdx.x >> 5) * swg_smem_per_warp(p.max_xlen, p.max_cols, p.trace_bytes, p.ops_words);
  uint8_t* sx = base; base += align16(p.max_xlen + 16);
  uint8_t* sy = base; base += align16(p.max_cols + 16);
  uint8_t* trace = base; base += align16(p.trace_bytes);
  uint32_t* obuf = (uint32_t*)base;
  DevWarp w;
  const uint32_t n_tasks = p.ctr->warp_tasks;  // not eligible for the thread kernel, or too few of their class
  for (;;) {
    uint32_t t = next_work(&p.ctr->round_work[p.round]);
    if (t >= n_tasks) break;
    tg_task_run<DevWarp, RMAX>(w, p.P.ix, p.bases, p.offs, p.tasks[p.sorted[t]], sx, sy, trace, obuf, p.ops_pool, &p.ctr->round_ops[p.round],
                               p.ops_cap, &p.ctr->flags, p.bound_stop != 0);
  }
}

// ---- task sorting: histogram -> bin starts -> scatter ------------------------------------------------------------
__device__ __forceinline__ uint32_t task_bin(const TgTask& t) {
  const int xlen = (int)t.xlen, bw = (int)t.bw, ylen = (int)t.ylen;
  const int ncols = ylen < xlen + bw ? ylen : xlen + bw;
  const int cls = t.pad0 ? 0 : tg_dpt_class(xlen, bw, t.x_drop);  // pad0: not for the thread kernel (tg_swg_extend_batch)
  // lanes of a warp should agree on the phase boundaries (bw) and on the number of columns: sort by both
  const int bwb = bw < 16 ? bw : 16 + (bw - 16 < 60 ? (bw - 16) >> 2 : 15);
  const int cb = (ncols < 255 ? ncols : 255) >> 4;
  return (uint32_t)(cls * TG_DPT_CBINS + bwb * 16 + (15 - cb));
}
__global__ void __launch_bounds__(256) k_round_hist(RoundParams p) {
  __shared__ uint32_t h[TG_DPT_NBINS];
  for (int i = threadIdx.x; i < TG_DPT_NBINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  unsigned long long n_tasks = p.ctr->round_tasks[p.round];
  if (n_tasks > p.task_cap) n_tasks = p.task_cap;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tasks; t += (unsigned long long)gridDim.x * blockDim.x)
    atomicAdd(&h[task_bin(p.tasks[t])], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < TG_DPT_NBINS; i += blockDim.x)
    if (h[i]) atomicAdd(&p.ctr->bin_count[i], h[i]);
}
// one block: exclusive scan of the bins; class ranges and warp-chunk table; clears the counts for the next round
#define TG_BINSCAN_THREADS 768
__global__ void __launch_bounds__(TG_BINSCAN_THREADS) k_round_binscan(RoundParams p) {
  __shared__ uint32_t cls_n[TG_DPT_NCLS + 1];
  __shared__ uint32_t cls_base[TG_DPT_NCLS + 1];
  if (threadIdx.x <= TG_DPT_NCLS) cls_n[threadIdx.x] = 0;
  __syncthreads();
  // thread t owns the bins [t * PER, t * PER + PER) of one class (TG_DPT_CBINS is a multiple of PER)
  constexpr int PER = TG_DPT_NBINS / TG_BINSCAN_THREADS;
  static_assert(TG_DPT_NBINS % TG_BINSCAN_THREADS == 0 && TG_DPT_CBINS % PER == 0, "bin layout");
  uint32_t v[PER], sum = 0;
  const int b0 = threadIdx.x * PER, cls = b0 / TG_DPT_CBINS;
#pragma unroll
  for (int k = 0; k < PER; k++) { v[k] = p.ctr->bin_count[b0 + k]; p.ctr->bin_count[b0 + k] = 0; sum += v[k]; }
  // exclusive scan of `sum` inside the class: warp scan + per-class atomics are overkill for <= 86 threads per class;
  // a shared array and a serial pass over the threads of the class by its first thread is cheap enough
  __shared__ uint32_t tsum[TG_BINSCAN_THREADS];
  tsum[threadIdx.x] = sum;
  __syncthreads();
  constexpr int TPC = TG_DPT_CBINS / PER;  // threads per class
  if (threadIdx.x % TPC == 0) {
    uint32_t acc = 0;
    for (int t = 0; t < TPC; t++) { const uint32_t x = tsum[threadIdx.x + t]; tsum[threadIdx.x + t] = acc; acc += x; }
    cls_n[cls] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // A class with few tasks cannot fill the GPU with one thread per extension and would hold the round for the
    // latency of a single long task (a wide band is ~100 us on one thread): such classes join class 0 on the
    // warp-cooperative kernel, which spreads one extension over 32 lanes.  Layout: class 0, small classes, big classes.
    uint32_t acc = 0, chunks = 0;
    bool small[TG_DPT_NCLS];
    for (int c = 0; c < TG_DPT_NCLS; c++) {
      const uint32_t thr = tg_dpt_wb(c) <= 32 ? 4096u : 8192u;
      small[c] = c == 0 || p.all_warp || cls_n[c] < thr;
      p.ctr->round_cls[p.round][c] = cls_n[c];
    }
    for (int c = 0; c < TG_DPT_NCLS; c++)
      if (small[c]) { cls_base[c] = acc; acc += cls_n[c]; }
    p.ctr->warp_tasks = acc;
    for (int c = 0; c < TG_DPT_NCLS; c++) {
      if (!small[c]) { cls_base[c] = acc; acc += cls_n[c]; }
      p.ctr->cls_start[c] = cls_base[c];
      p.ctr->cls_end[c] = cls_base[c] + cls_n[c];
      p.ctr->cls_chunk0[c] = chunks;
      if (!small[c]) chunks += (cls_n[c] + 31) / 32;
    }
    p.ctr->cls_chunk0[TG_DPT_NCLS] = chunks;
  }
  __syncthreads();
  uint32_t acc = cls_base[cls] + tsum[threadIdx.x];
#pragma unroll
  for (int k = 0; k < PER; k++) { p.ctr->bin_cursor[b0 + k] = acc; acc += v[k]; }
}
__global__ void __launch_bounds__(256) k_round_scatter(RoundParams p) {
  __shared__ uint32_t h[TG_DPT_NBINS];
  unsigned long long n_tasks = p.ctr->round_tasks[p.round];
  if (n_tasks > p.task_cap) n_tasks = p.task_cap;
  const unsigned long long per_block = 256ull * 8;
  for (unsigned long long base = (unsigned long long)blockIdx.x * per_block; base < n_tasks; base += (unsigned long long)gridDim.x * per_block) {
    for (int i = threadIdx.x; i < TG_DPT_NBINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t bin[8], rank[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned long long t = base + (unsigned long long)k * 256 + threadIdx.x;
      bin[k] = 0xFFFFFFFFu;
      if (t < n_tasks) { bin[k] = task_bin(p.tasks[t]); rank[k] = atomicAdd(&h[bin[k]], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TG_DPT_NBINS; i += blockDim.x)
      if (h[i]) h[i] = atomicAdd(&p.ctr->bin_cursor[i], h[i]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned long long t = base + (unsigned long long)k * 256 + threadIdx.x;
      if (bin[k] != 0xFFFFFFFFu) p.sorted[h[bin[k]] + rank[k]] = (uint32_t)t;
    }
    __syncthreads();
  }
}

// ---- thread-per-extension kernel (tg_dpt.h) ----------------------------------------------------------------------------
#define DPT_OBUF 12u
template <int WB>
__device__ __forceinline__ void dpt_task(const RoundParams& p, TgTask& t, bool active, const TgDptMem& m, uint32_t* obuf, int lane) {
  TgDptResult res{0, 0, 0, 0};
  TgDptY ys;
  uint32_t n_ops = 0;
  int xlen = 0, bw = 0;
  if (active) {
    xlen = (int)t.xlen; bw = (int)t.bw;
    const int ylen = (int)t.ylen;
    const int ncols = ylen < xlen + bw ? ylen : xlen + bw;
    ys.seq = tg_seq_of(p.P.ix, t.seqsel); ys.y0 = t.y0; ys.ncols = ncols; ys.side = t.side; ys.word = 0; ys.need = 0;
    ys.word_next = 0; ys.need_next = 0; ys.t_next = -1;
    tg_dpt_profile(m, p.rp + (size_t)t.read * p.rp_words, t.xoff, xlen, t.side);
    tg_dpt_fill<WB>(m, ys, xlen, ncols, bw, t.x_drop, p.bound_stop != 0, res);
    // one traceback pass: the RLE words go to a small shared-memory buffer (alignments have few runs); only an
    // alignment with more than DPT_OBUF runs is walked a second time after the pool allocation
    uint32_t* ob = obuf;
    n_ops = tg_dpt_traceback<WB>(m, ys, xlen, bw, res, [ob](uint32_t i, uint32_t kind, uint32_t run) {
      if (i < DPT_OBUF) ob[i * 128] = kind | (run << 3);
    });
  }
  __syncwarp();
  // one pool allocation per warp
  uint32_t incl = n_ops;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(TG_FULL, incl, d);
    if (lane >= d) incl += v;
  }
  const uint32_t total = __shfl_sync(TG_FULL, incl, 31);
  unsigned long long base = 0;
  if (lane == 0 && total) base = atomicAdd(&p.ctr->round_ops[p.round], (unsigned long long)total);
  base = __shfl_sync(TG_FULL, base, 0);
  if (base + total > p.ops_cap) {
    if (lane == 0) atomicOr(&p.ctr->flags, TG_FLAG_OPS_POOL);
    n_ops = 0;
  }
  if (active) {
    const unsigned long long dst = base + incl - n_ops;
    uint32_t* out = p.ops_pool + dst;
    if (n_ops > DPT_OBUF) tg_dpt_traceback<WB>(m, ys, xlen, bw, res, [out](uint32_t i, uint32_t kind, uint32_t run) { out[i] = kind | (run << 3); });
    else
      for (uint32_t i = 0; i < n_ops; i++) out[i] = obuf[i * 128];
    t.score = res.score; t.xend = (uint32_t)res.xend; t.yend = (uint32_t)res.yend; t.cells = res.cells;
    t.ops_off = (uint32_t)dst; t.ops_n = n_ops;
  }
}

// One kernel per band class (round 1 had four class groups): a class is not held to the register budget of a wider one
// that shares its kernel, and an SM's instruction cache sees one unrolled band per kernel.  Resident CTAs per SM were
// tuned per class on the bench workload (DP section per 4 M reads: four groups 19.5 ms; per class with the groups'
// budgets 18.4 ms; natural register counts without spills -- 72 ... 255 -- 20.4 ms; the budgets below, with small spills
// in the 40- to 64-slot classes, 18.0 ms): occupancy is worth more than spill-free code here.
template <int CLS> struct DptCls { static constexpr int min_blocks = CLS == 1 ? 8 : CLS == 2 ? 7 : CLS == 3 ? 6 : CLS == 4 ? 5 : CLS <= 6 ? 4 : CLS <= 9 ? 3 : 2; };

template <int CLS>
__global__ void __launch_bounds__(128, DptCls<CLS>::min_blocks) k_round_dpt(RoundParams p) {
  __shared__ uint32_t msk[32 * 128];
  __shared__ uint32_t obuf_s[DPT_OBUF * 128];
  uint32_t* obuf = obuf_s + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  TgDptMem m;
  m.msk = msk + threadIdx.x; m.mstride = 128;
  m.tr = p.dpt_trace + p.dpt_trace_off[CLS] + (size_t)gw * p.dpt_trace_words[CLS] + lane; m.tstride = 32;
  m.one = p.dpt_one; m.k128 = p.dpt_k128;
  const uint32_t n_chunks = p.ctr->cls_chunk0[CLS + 1] - p.ctr->cls_chunk0[CLS];
  const uint32_t start = p.ctr->cls_start[CLS], end = p.ctr->cls_end[CLS];
  for (;;) {
    const uint32_t g = next_work(&p.ctr->round_work2[p.round][CLS]);
    if (g >= n_chunks) break;
    const uint32_t first = start + g * 32u;
    const bool active = first + lane < end;
    TgTask& t = p.tasks[active ? p.sorted[first + lane] : p.sorted[first]];
    dpt_task<tg_dpt_wb(CLS)>(p, t, active, m, obuf, lane);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) k_round_post(RoundParams p) {
  DevThread w;
  unsigned long long lo, hi;
  round_item_range(p, lo, hi);
  // items appended from here on (by scan) belong to the next round
  if (blockIdx.x == 0 && threadIdx.x == 0) p.ctr->round_end[p.round] = p.ctr->items_used;
  // (item order, not the locus order prep uses: post is bound by the latency of its loads, and items are in read order,
  // so neighbouring threads find their read states, packed reads and outputs in neighbouring memory -- 0.53 ms against
  // 0.63 ms per 1 M items in locus order; storing the hit tables in prep's order as well did not help: 46.6 -> 47.1 ms per step)
  for (unsigned long long it = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < hi; it += (unsigned long long)gridDim.x * blockDim.x) {
    TgItemRes& ir = p.ires[it];
    if (ir.flags & TG_IF_FAIL) continue;
    tg_item_post<DevThread>(w, p.P, p.st[ir.read], p.hits[it], p.tasks, p.ops_pool, ir, p.cands[it], p.hp, &p.ctr->flags);
  }
}

__global__ void __launch_bounds__(128) k_round_scan(RoundParams p) {
  DevThread w;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    if (p.st[r].status != TG_RS_ACTIVE) continue;
    TgReadState st = p.st[r];
    if (st.batch_n == 0) continue;
    if (!tg_scan_read<DevThread>(w, p.P.opts, st, p.ires, r, &p.ctr->items_used, p.item_cap, &p.ctr->flags)) { mark_complex(p, r, 1); continue; }
    p.st[r] = st;
    if (st.status == TG_RS_ACTIVE) warp_agg_add(&p.ctr->round_active[p.round], 1ull);
  }
}

// src/aligner.rs:177-187 for the reads that finished on the round path; the others join the single-warp list
__global__ void __launch_bounds__(128) k_round_final(RoundParams p) {
  DevThread w;
  unsigned long long cells = 0, n_ext = 0, hits = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += gridDim.x * blockDim.x) {
    const TgReadState st = p.st[r];
    if (st.status == TG_RS_COMPLEX || st.status == TG_RS_FINAL) continue;
    if (st.status == TG_RS_ACTIVE) {
      if (!p.early) mark_complex(p, r, 3);  // early pass: unfinished reads simply wait for the last one
      continue;
    }
    p.st[r].status = TG_RS_FINAL;
    if (st.n_acc <= TG_FINAL_SMALL) {
      uint32_t items[TG_FINAL_SMALL];
      uint16_t order[TG_FINAL_SMALL], tmp[TG_FINAL_SMALL];
      tg_round_final<DevThread, uint16_t>(w, p.P, st, p.cands, p.ires, p.hp.w, items, order, tmp, p.out, r);
    } else {
      const unsigned long long base = warp_agg_add(&p.ctr->fin_used, 3ull * st.n_acc);
      if (base + 3ull * st.n_acc > p.fin_cap) { mark_complex(p, r, 2); continue; }  // (status becomes COMPLEX again)
      uint32_t* f = p.fin + base;
      tg_round_final<DevThread, uint32_t>(w, p.P, st, p.cands, p.ires, p.hp.w, f, f + st.n_acc, f + 2 * (size_t)st.n_acc, p.out, r);
    }
    cells += st.cells; n_ext += st.n_ext; hits += st.hits;
    if (!p.early && p.late && p.st[r].status == TG_RS_FINAL)
      p.late[warp_agg_add(&p.ctr->n_late, 1ull)] = make_ulonglong2(p.out.first32 ? (unsigned long long)p.out.first32[r] : p.out.read_aln_first[r],
                                                                   (unsigned long long)r | ((unsigned long long)p.out.read_aln_count[r] << 32));
  }
  cells = warp_sum(cells); n_ext = warp_sum(n_ext); hits = warp_sum(hits);
  if ((threadIdx.x & 31) == 0) {
    if (cells) atomicAdd(&p.ctr->cells, cells);
    if (n_ext) atomicAdd(&p.ctr->n_ext, n_ext);
    if (hits) atomicAdd(&p.ctr->hits, hits);
  }
}

// ---------------------------------------------------------------------------------------------------
// SWG microbench / batched SwgExtend::extend
// ---------------------------------------------------------------------------------------------------
struct SwgParams {
  const uint8_t* xs; const uint64_t* xoff; const uint8_t* ys; const uint64_t* yoff;
  uint32_t n;
  const uint32_t* bw; const int32_t* x_drop;
  int32_t* score; uint32_t* xend; uint32_t* yend;
  uint64_t* task_off; uint32_t* task_len;
  uint32_t* ops; unsigned long long ops_cap;
  uint32_t max_xlen, max_cols, trace_bytes, ops_words;
  int bound_stop;
  DevCounters* ctr;
  const uint32_t* list;  // when set: only the tasks list[0 .. ctr->warp_tasks) (those the thread kernels do not take)
  // thread-kernel path: tasks in round-pipeline form, x as packed "reads", y as one packed text
  TgTask* tasks;
  uint64_t* xpk;
  uint64_t* ypk;
  const uint64_t* ysym_off;  // [n] first symbol of task t in ypk (multiple of 16)
  uint32_t rp_words;
  const uint32_t* dp_ops;    // operations written by the thread kernels (generation order)
};
__host__ __device__ inline size_t swg_smem_per_warp(uint32_t max_xlen, uint32_t max_cols, uint32_t trace_bytes, uint32_t ops_words) {
  return align16(max_xlen + 16) + align16(max_cols + 16) + align16(trace_bytes) + align16((size_t)ops_words * 4);
}
template <int RMAX>
__global__ void __launch_bounds__(TG_WARPS_PER_CTA * 32) k_swg_batch(SwgParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* base = smem + (threadIdx.x >> 5) * swg_smem_per_warp(p.max_xlen, p.max_cols, p.trace_bytes, p.ops_words);
  uint8_t* sx = base; base += align16(p.max_xlen + 16);
  uint8_t* sy = base; base += align16(p.max_cols + 16);
  uint8_t* trace = base; base += align16(p.trace_bytes);
  uint32_t* obuf = (uint32_t*)base;
  DevWarp w;
  const int lane = threadIdx.x & 31;
  unsigned long long cells = 0, n_ext = 0;
  for (;;) {
    uint32_t t = next_work(&p.ctr->work_swg);
    if (p.list) {
      if (t >= p.ctr->warp_tasks) break;
      t = p.list[t];
    } else if (t >= p.n) break;
    const uint64_t x0 = p.xoff[t], y0 = p.yoff[t];
    const int xlen = (int)(p.xoff[t + 1] - x0);
    const int ylen_full = (int)min((unsigned long long)(p.yoff[t + 1] - y0), 0x7fffffffull);
    const int bw = (int)p.bw[t];
    const int ylen = ylen_full > xlen + bw ? xlen + bw + 1 : ylen_full;
    const int ncols = ylen < xlen + bw ? ylen : xlen + bw;
    for (int i = lane; i < xlen; i += 32) sx[i] = p.xs[x0 + i];
    for (int i = lane; i < ncols; i += 32) sy[i] = p.ys[y0 + i];
    __syncwarp();
    TgSwgResult res{0, 0, 0};
    TgOps o{obuf, 0};
    tg_swg_extend<DevWarp, RMAX>(w, sx, sy, xlen, ylen, bw, p.x_drop[t], trace, res, o, cells, n_ext, p.bound_stop != 0);
    unsigned long long dst = 0;
    if (lane == 0 && o.n) dst = atomicAdd(&p.ctr->swg_ops_used, (unsigned long long)o.n);
    dst = __shfl_sync(TG_FULL, dst, 0);
    if (dst + o.n > p.ops_cap) {
      if (lane == 0) atomicOr(&p.ctr->flags, TG_FLAG_OPS_POOL);
    } else {
      for (uint32_t i = lane; i < o.n; i += 32) p.ops[dst + i] = obuf[o.n - 1 - i];  // buffer holds rev(operations)
    }
    if (lane == 0) {
      p.score[t] = res.score; p.xend[t] = (uint32_t)res.xend; p.yend[t] = (uint32_t)res.yend;
      p.task_off[t] = dst; p.task_len[t] = o.n;
      if (p.list) p.tasks[t].pad0 = 1;  // done here: k_swg_collect must not take it
    }
    __syncwarp();
  }
  cells = warp_sum(cells);
  n_ext = warp_sum(n_ext);
  if (lane == 0) {
    if (cells) atomicAdd(&p.ctr->cells, cells);
    if (n_ext) atomicAdd(&p.ctr->n_ext, n_ext);
  }
}

// SwgExtend::extend batch on the thread-per-extension kernels: every (x, y) pair becomes a round-pipeline task whose
// x is a packed "read" and whose y lives in one packed text.  Bytes are compared AS GIVEN by the reference, so only
// pairs made of the symbols A C G N T (upper case) may use the 4-bit codes; the others keep pad0 = 1 and run on
// k_swg_batch with raw bytes.
__device__ __forceinline__ uint32_t exact_code(uint8_t c) {
  switch (c) {
    case 'A': return TG_C_A;
    case 'C': return TG_C_C;
    case 'G': return TG_C_G;
    case 'N': return TG_C_N;
    case 'T': return TG_C_T;
    default: return 0xFFu;
  }
}
__global__ void __launch_bounds__(128) k_swg_prepare(SwgParams p) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < p.n; t += gridDim.x * blockDim.x) {
    const uint64_t x0 = p.xoff[t], y0 = p.yoff[t];
    const uint32_t xlen = (uint32_t)(p.xoff[t + 1] - x0);
    const uint64_t ylen_full = p.yoff[t + 1] - y0;
    const uint32_t bw = p.bw[t];
    const uint32_t ylen = ylen_full > (uint64_t)xlen + bw ? xlen + bw + 1 : (uint32_t)ylen_full;
    const uint32_t ncols = ylen < xlen + bw ? ylen : xlen + bw;
    TgTask& k = p.tasks[t];
    k.read = t; k.xoff = 0; k.xlen = xlen; k.ylen = ylen; k.y0 = p.ysym_off[t]; k.bw = bw; k.x_drop = p.x_drop[t];
    k.side = 0; k.seqsel = 0; k.pad1 = 0;
    k.score = 0; k.xend = 0; k.yend = 0; k.cells = 0; k.ops_off = 0; k.ops_n = 0;
    bool bad = xlen == 0 || ylen == 0 || tg_dpt_class((int)xlen, (int)bw, k.x_drop) == 0;
    if (!bad) {
      uint64_t* xw = p.xpk + (size_t)t * p.rp_words;
      for (uint32_t wi = 0; wi * 16 < xlen; wi++) {
        uint64_t word = 0;
        for (uint32_t u = 0; u < 16; u++) {
          const uint32_t i = wi * 16 + u;
          uint32_t c = TG_C_PAD;
          if (i < xlen) { c = exact_code(p.xs[x0 + i]); bad |= c == 0xFFu; }
          word |= (uint64_t)(c & 15u) << ((15 - u) * 4);
        }
        xw[wi] = word;
      }
      uint64_t* yw = p.ypk + (k.y0 >> 4);
      for (uint32_t wi = 0; wi * 16 < ncols; wi++) {
        uint64_t word = 0;
        for (uint32_t u = 0; u < 16; u++) {
          const uint32_t i = wi * 16 + u;
          uint32_t c = TG_C_PAD;
          if (i < ncols) { c = exact_code(p.ys[y0 + i]); bad |= c == 0xFFu; }
          word |= (uint64_t)(c & 15u) << ((15 - u) * 4);
        }
        yw[wi] = word;
      }
    }
    k.pad0 = bad ? 1 : 0;
  }
}
// results of the thread-kernel tasks -> the flat result arrays of tg_swg_extend_batch (operations in forward order)
__global__ void __launch_bounds__(128) k_swg_collect(SwgParams p) {
  unsigned long long cells = 0, n_ext = 0;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < p.n; t += gridDim.x * blockDim.x) {
    const TgTask k = p.tasks[t];
    if (k.pad0) continue;
    const unsigned long long dst = warp_agg_add(&p.ctr->swg_ops_used, (unsigned long long)k.ops_n);
    if (dst + k.ops_n > p.ops_cap) atomicOr(&p.ctr->flags, TG_FLAG_OPS_POOL);
    else
      for (uint32_t i = 0; i < k.ops_n; i++) p.ops[dst + i] = p.dp_ops[k.ops_off + k.ops_n - 1 - i];
    p.score[t] = k.score; p.xend[t] = k.xend; p.yend[t] = k.yend;
    p.task_off[t] = dst; p.task_len[t] = k.ops_n;
    cells += k.cells; n_ext++;
  }
  cells = warp_sum(cells); n_ext = warp_sum(n_ext);
  if ((threadIdx.x & 31) == 0) {
    if (cells) atomicAdd(&p.ctr->cells, cells);
    if (n_ext) atomicAdd(&p.ctr->n_ext, n_ext);
  }
}

// HBM random-access yardstick for the seeding roofline: every thread issues independent 16-B loads (one
// 32-B sector each) at hashed positions of a table much larger than L2.
__global__ void __launch_bounds__(256) k_random_gather(const uint4* __restrict__ table, uint64_t mask, uint32_t per_thread, uint32_t* sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < per_thread; i += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = __ldg(table + (tg_mix64(tid * per_thread + i + u + 0x9e3779b97f4a7c15ULL) & mask));
#pragma unroll
    for (int u = 0; u < 4; u++) acc ^= v[u].x ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;  // keeps the loads alive
}

// A few counters -> pinned host memory, written by the device itself.  The host decisions inside a batch (how much early
// output there is, whether a read is still active) used to travel as 16-byte D2H copies on the main stream; those queue on
// the copy engine BEHIND the early record copies of the same batch (hundreds of MB, slow when 8 GPUs share one host) and
// stalled the kernels behind them for milliseconds.
__global__ void k_host_snapshot(volatile unsigned long long* host_dst, const unsigned long long* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) host_dst[i] = src[i];
  __threadfence_system();
}

// Integer-pipe yardstick for the SWG roofline: dependency-free streams of the two instructions the DP inner
// loop is made of -- VIADDMNMX (max(a + b, c)) and VIMNMX3 (max(a, b, c)) -- 16 independent chains per thread, enough
// warps per scheduler to cover the pipe latency.  The measured lane-operations per second replace the assumed
// "16 lanes per clock per scheduler" in the roofline's numerator.
template <int KIND>
__global__ void __launch_bounds__(256) k_int_peak(int iters, int k1, int k2, int* sink) {
  int v[16], w[16];
#pragma unroll
  for (int u = 0; u < 16; u++) { v[u] = (int)threadIdx.x * (u + 1) - k2; w[u] = k2 - u * (int)blockIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (KIND == 0) v[u] = __viaddmax_s32(v[u], k1, w[u]);
      else v[u] = __vimax3_s32(v[u], v[(u + 1) & 15], w[u]);  // (neighbour of the previous iteration: 16 independent ops per iteration)
    }
  }
  int acc = 0;
#pragma unroll
  for (int u = 0; u < 16; u++) acc ^= v[u];
  if (acc == 0x7fffffff) sink[0] = acc;  // keeps the chains alive
}

// ---------------------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------------------
#define CU_CHECK(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return tg_fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  tg_status ensure(size_t n) {
    if (n <= cap) return TG_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = want;
    return TG_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  tg_status ensure(size_t n) {
    if (n <= cap) return TG_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    cap = want;
    return TG_OK;
  }
  // grow while keeping the first `keep` bytes (results of earlier chunks)
  tg_status ensure_keep(size_t n, size_t keep) {
    if (n <= cap) return TG_OK;
    void* old = p;
    size_t want = n + n / 2 + 256;
    void* q = nullptr;
    cudaError_t e = cudaMallocHost(&q, want);
    if (e != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    if (old && keep) memcpy(q, old, keep);
    if (old) cudaFreeHost(old);
    p = q; cap = want;
    return TG_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// A k-mer table depends only on (index, min_seed_len): contexts of one index with the same k share one table (two
// contexts on a GPU used to hold 4.3 GB each for chr21).  Reference counted; freed with its last context.
struct TgKmerTable {
  uint32_t k = 0;
  TgSlot* slots = nullptr;
  uint64_t n_slots = 0;
  int refs = 0;
};
struct tg_index {
  int device = 0;
  void* d_blob = nullptr;
  bool owns = false;
  TgBlobHeader hdr;
  TgIndexDev dev;
  mutable std::recursive_mutex table_mu;  // (recursive: a failing tg_ctx_create destroys its context while it holds the lock)
  mutable std::vector<TgKmerTable> tables;
};

// Two helper threads per context (started with the first large batch, parked on a condition variable in between) that
// check the read offsets of the later input chunks while the calling thread checks the first chunk and issues its copy:
// one core streams 32 MB of offsets (4 M reads) in 3.4 ms, during which the GPU had nothing to do.
struct OffsCheck {
  struct Job {
    const uint64_t* o = nullptr; uint32_t m = 0; uint64_t wide = 0; uint32_t longest = 0; bool pending = false;
    // second kind of job: entries [0, n_fix) of a fix-up list (first index, read | count << 32) applied to the host result
    const unsigned long long* fix = nullptr; uint64_t n_fix = 0; void* first = nullptr; uint32_t* count = nullptr; bool compact = false;
  };
  static void apply_fix(const Job& j) {
    for (uint64_t i = 0; i < j.n_fix; i++) {
      const uint32_t r = (uint32_t)(j.fix[2 * i + 1] & 0xFFFFFFFFull);
      if (j.compact) ((uint32_t*)j.first)[r] = (uint32_t)j.fix[2 * i];
      else ((uint64_t*)j.first)[r] = j.fix[2 * i];
      j.count[r] = (uint32_t)(j.fix[2 * i + 1] >> 32);
    }
  }
  std::thread th[2];
  std::mutex mu;
  std::condition_variable cv, cv_done;
  Job job[2];
  bool started = false, stop = false;
  static void scan(Job& j) {
    uint64_t wide = 0;  // non-zero when some length is >= 1024 or negative (offsets decreasing)
    uint32_t longest = 0;
    const uint64_t* o = j.o;
    for (uint32_t i = 0; i < j.m; i++) {
      const uint64_t d = o[i + 1] - o[i];
      wide |= d >> 10;
      longest = (uint32_t)d > longest ? (uint32_t)d : longest;
    }
    j.wide = wide; j.longest = longest;
  }
  void run(int k) {
    std::unique_lock<std::mutex> l(mu);
    for (;;) {
      cv.wait(l, [&] { return stop || job[k].pending; });
      if (stop) return;
      l.unlock();
      if (job[k].fix) apply_fix(job[k]); else scan(job[k]);
      l.lock();
      job[k].pending = false;
      cv_done.notify_all();
    }
  }
  void submit(const uint64_t* o, uint32_t m) {  // offsets o[0 .. m]: split between the two helpers
    {
      std::lock_guard<std::mutex> l(mu);
      if (!started) { started = true; th[0] = std::thread([this] { run(0); }); th[1] = std::thread([this] { run(1); }); }
      const uint32_t h = m / 2;
      job[0].fix = nullptr; job[1].fix = nullptr;
      job[0].o = o; job[0].m = h; job[0].pending = true;
      job[1].o = o + h; job[1].m = m - h; job[1].pending = true;
    }
    cv.notify_all();
  }
  // the late reads' first / count (a few per cent of a batch, scattered): thirds of the list on the helpers and the caller
  void fix_up(const unsigned long long* list, uint64_t n, void* first, uint32_t* count, bool compact) {
    Job mine;
    mine.fix = list; mine.n_fix = n; mine.first = first; mine.count = count; mine.compact = compact;
    if (!started || n < 4096) { apply_fix(mine); return; }
    const uint64_t a = n / 3, b = 2 * n / 3;
    {
      std::lock_guard<std::mutex> l(mu);
      for (int k = 0; k < 2; k++) { job[k].first = first; job[k].count = count; job[k].compact = compact; job[k].pending = true; }
      job[0].fix = list + 2 * a; job[0].n_fix = b - a;
      job[1].fix = list + 2 * b; job[1].n_fix = n - b;
    }
    cv.notify_all();
    mine.n_fix = a;
    apply_fix(mine);
    uint64_t w; uint32_t l;
    wait(w, l);
  }
  void wait(uint64_t& wide, uint32_t& longest) {
    std::unique_lock<std::mutex> l(mu);
    cv_done.wait(l, [&] { return !job[0].pending && !job[1].pending; });
    wide = job[0].wide | job[1].wide;
    longest = std::max(job[0].longest, job[1].longest);
  }
  ~OffsCheck() {
    if (!started) return;
    { std::lock_guard<std::mutex> l(mu); stop = true; }
    cv.notify_all();
    th[0].join(); th[1].join();
  }
};

struct tg_ctx {
  const tg_index* ix = nullptr;
  OffsCheck offs_check;
  tg_opts opts;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};  // the band-class kernels of a round's DP are spread over these
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_dp0[TG_MAX_ROUNDS] = {}, ev_dp1[TG_MAX_ROUNDS] = {};  // DP section of every round (timing)
  int rounds_run = 0;
  float last_dp_ms = 0.f;
  float round_dp_ms[TG_MAX_ROUNDS] = {};  // DP section of every round of the last batch (debug)
  unsigned long long* h_active = nullptr;  // pinned
  float last_seed_ms = 0.f, last_extend_ms = 0.f;
  int exact_cells = 0;  // 1: run every column the reference runs (swg_cells == reference count)
  int n_sms = 0;
  TgSlot* slots = nullptr;
  uint64_t n_slots = 0;
  DevCounters* d_ctr = nullptr;
  DevCounters* h_ctr = nullptr;  // pinned
  // inputs
  DevBuf d_bases, d_offs;
  // seeding
  DevBuf d_seeds, d_seed_first, d_seed_count, d_probe, d_queue, d_queue2;
  uint64_t seed_cap = 0;
  // extension
  DevBuf d_cands, d_arena, d_order, d_aln_first, d_aln_count, d_alns, d_ops;
  uint64_t alns_cap = 0, ops_cap = 0;
  uint32_t scratch_warps = 0;
  int dpt_occ[TG_DPT_NCLS] = {};  // resident CTAs per SM of k_round_dpt<cls> (queried once)
  // round pipeline scratch
  DevBuf r_state, r_hits, r_ires, r_cands, r_hops, r_fin, r_rp, r_tasks, r_ops, r_complex, r_sorted, r_dpt_trace, r_late, r_perm, r_ikey, r_ibins;
  PinBuf h_late;
  int early_rows = 0;  // host-buffer path: first/count went to the host after round 1; the last pass sends a fix-up list
  uint64_t round_task_cap = 0, round_ops_cap = 0, item_cap = 0, hops_cap = 0;
  int use_rounds = 1;
  int item_sort = 2;  // rounds whose items are handed out in locus order (TG_ITEM_SORT; 0 = off)
  uint32_t tiny_batch = 256;     // below this many reads every extension runs on the warp kernel (TG_TINY_BATCH)
  uint32_t small_batch = 32768;  // below this many reads a batch is latency-bound: no early output, no item sort (TG_SMALL_BATCH)
  // chunked host-buffer path: results of chunk k start at these pool positions / read row
  unsigned long long base_alns = 0, base_ops = 0;
  uint32_t out_row0 = 0;
  uint32_t pool_reads = 0;  // reads the output pools are sized for (the whole batch when chunking)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  int early_out = 0;                 // host-buffer path: records of reads finished after round 1 go to the host early
  unsigned long long* h_snap = nullptr;  // pinned {alns_used, ops_used} after the early final pass
  unsigned long long early_alns = 0, early_ops = 0;  // what has been sent to the host already
  uint32_t in_chunks = 0, in_chunk_reads = 0;        // host-buffer path: seeding follows the input copies chunk by chunk
  std::vector<cudaEvent_t> ev_in;
  uint32_t chunk_reads = 0;  // 0: automatic (an eighth of the batch, at least 262,144 reads)
  uint64_t n_launches = 0;  // kernels launched by the last batch call
  // host results
  PinBuf h_first, h_count, h_alns, h_ops, h_seeds, h_seed_first, h_seed_count;
  PinBuf h2_first, h2_count, h2_alns, h2_ops;  // second result set (tg_ctx_set_result_buffers(ctx, 2)): swapped in before every call
  int result_sets = 1;
  // where the records of the current host-buffer call go: the context's own pinned buffers (grown on demand) or a
  // segment of a caller-owned result (tg_multi_align_batch: fixed capacity, the call reports what it needs instead)
  struct HostOut {
    uint8_t* first = nullptr;   // u64[n] (wide) or u32[n] (compact)
    uint32_t* count = nullptr;
    uint8_t* alns = nullptr;    // tg_aln[] or tg_aln_c[]
    uint32_t* ops = nullptr;
    size_t alns_cap = 0, ops_cap = 0;  // elements
    bool external = false;
  } ho;
  int compact = 0;                                  // records leave as tg_aln_c, firsts as u32
  unsigned long long first_base = 0, ops_base = 0;  // compact: added to every first index / operation offset on the device
  bool ho_overflow = false;                         // external segment too small: need_alns / need_ops say how much the shard needs
  uint64_t need_alns = 0, need_ops = 0;
  double last_wall_ms = 0.0;
  // swg batch
  DevBuf s_x, s_xo, s_y, s_yo, s_bw, s_xd, s_score, s_xe, s_ye, s_toff, s_tlen, s_ops, s_ypk, s_ysym;
};

namespace {

tg_status adopt_blob(tg_index* ix, const void* d_blob, size_t nbytes) {
  TgBlobHeader h;
  if (nbytes < sizeof(h)) return tg_fail(TG_ERR_INVALID, "index blob too small");
  CU_CHECK(cudaMemcpy(&h, d_blob, sizeof(h), cudaMemcpyDeviceToHost));
  if (const char* why = tg_blob_check_header(h, nbytes, true))
    return tg_fail(TG_ERR_INVALID, std::string("not a usable thermite_gpu index blob: ") + why);
  ix->hdr = h;
  const uint8_t* b = (const uint8_t*)d_blob;
  TgIndexDev& d = ix->dev;
  d.text4 = (const uint64_t*)(b + h.off_text4);
  d.sa = (const uint32_t*)(b + h.off_sa);
  d.refs = (const TgRef*)(b + h.off_refs);
  d.exon_nodes = (const TgTreeNode*)(b + h.off_exon_nodes);
  d.gene_nodes = (const TgTreeNode*)(b + h.off_gene_nodes);
  d.exon_stab = (const TgStab*)(b + h.off_exon_stab);
  d.gene_stab = (const TgStab*)(b + h.off_gene_stab);
  d.n_exon_stab = (uint32_t)h.n_exon_nodes; d.n_gene_stab = (uint32_t)h.n_gene_nodes;
  d.exon_maxlen = (uint32_t)h.exon_maxlen; d.gene_maxlen = (uint32_t)h.gene_maxlen;
  d.tx_seq_off = (const uint64_t*)(b + h.off_tx_seq_off);
  d.tx_exon_off = (const uint32_t*)(b + h.off_tx_exon_off);
  d.te_start = (const uint32_t*)(b + h.off_te_start);
  d.te_end = (const uint32_t*)(b + h.off_te_end);
  d.txseq4 = (const uint64_t*)(b + h.off_txseq4);
  d.text_len = h.text_len;
  d.n_refs = (uint32_t)h.n_refs;
  d.n_txs = (uint32_t)h.n_txs;
  d.exon_root = (int32_t)h.exon_root;
  d.gene_root = (int32_t)h.gene_root;
  return TG_OK;
}

uint32_t band_for(const tg_opts& o, uint32_t L) {  // src/aligner.rs:130-138
  float prod = o.min_aln_score_percent * (float)L;
  int32_t s = (int32_t)prod;
  if (o.min_aln_score > s) s = o.min_aln_score;
  if (s < 0) return 0;
  return L > (uint32_t)s ? L - (uint32_t)s : 0;
}

tg_status check_flags(int flags) {
  if (flags & TG_FLAG_READ_CAP)
    return tg_fail(TG_ERR_CAPACITY, "a read accepted more than TG_MAX_ALNS_PER_READ alignments");
  if (flags & TG_FLAG_ARENA) return tg_fail(TG_ERR_CAPACITY, "per-read operation arena exhausted");
  if (flags & TG_FLAG_YLEN) return tg_fail(TG_ERR_INTERNAL, "an alignment starts on another Ref than its seed hit");
  return TG_OK;
}

}  // namespace

extern "C" {

tg_status tg_index_create(const tg_index_host* hix, int device, tg_index** out) {
  TG_GUARD_BEGIN
  if (!hix || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return tg_fail(TG_ERR_CUDA, "no CUDA device: libthermite_gpu has no CPU fallback");
  CU_CHECK(cudaSetDevice(device));
  auto* ix = new tg_index();
  ix->device = device;
  size_t nb = hix->hdr()->device_bytes;  // (validated when the host index was created / loaded: <= blob size)
  cudaError_t e = cudaMalloc(&ix->d_blob, nb);
  if (e != cudaSuccess) { delete ix; return tg_fail(TG_ERR_CUDA, std::string("cudaMalloc(index): ") + cudaGetErrorString(e)); }
  ix->owns = true;
  e = cudaMemcpy(ix->d_blob, hix->blob.data(), nb, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(ix->d_blob); delete ix; return tg_fail(TG_ERR_CUDA, std::string("cudaMemcpy(index): ") + cudaGetErrorString(e)); }
  tg_status st = adopt_blob(ix, ix->d_blob, nb);
  if (st != TG_OK) { cudaFree(ix->d_blob); delete ix; return st; }
  *out = ix;
  return TG_OK;
  TG_GUARD_END
}

tg_status tg_index_create_from_device_blob(const void* device_blob, size_t nbytes, int device, tg_index** out) {
  TG_GUARD_BEGIN
  if (!device_blob || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  CU_CHECK(cudaSetDevice(device));
  auto* ix = new tg_index();
  ix->device = device;
  ix->d_blob = const_cast<void*>(device_blob);
  ix->owns = false;
  tg_status st = adopt_blob(ix, device_blob, nbytes);
  if (st != TG_OK) { delete ix; return st; }
  *out = ix;
  return TG_OK;
  TG_GUARD_END
}

void tg_index_destroy(tg_index* ix) {
  if (!ix) return;
  if (ix->owns && ix->d_blob) { cudaSetDevice(ix->device); cudaFree(ix->d_blob); }
  delete ix;
}

void tg_ctx_destroy(tg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->ix->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->d_bases, &c->d_offs, &c->d_seeds, &c->d_seed_first, &c->d_seed_count, &c->d_probe, &c->d_queue, &c->d_queue2, &c->d_cands, &c->d_arena, &c->d_order, &c->r_state, &c->r_hits, &c->r_ires, &c->r_cands, &c->r_hops, &c->r_fin, &c->r_rp, &c->r_tasks, &c->r_ops, &c->r_complex, &c->r_sorted, &c->r_dpt_trace, &c->r_late, &c->r_perm, &c->r_ikey, &c->r_ibins,
                    &c->d_aln_first, &c->d_aln_count, &c->d_alns, &c->d_ops, &c->s_x, &c->s_xo, &c->s_y, &c->s_yo, &c->s_bw,
                    &c->s_xd, &c->s_score, &c->s_xe, &c->s_ye, &c->s_toff, &c->s_tlen, &c->s_ops, &c->s_ypk, &c->s_ysym})
    b->release();
  for (PinBuf* b : {&c->h_first, &c->h_count, &c->h_alns, &c->h_ops, &c->h_seeds, &c->h_seed_first, &c->h_seed_count, &c->h_late,
                    &c->h2_first, &c->h2_count, &c->h2_alns, &c->h2_ops}) b->release();
  if (c->slots && c->ix) {
    std::lock_guard<std::recursive_mutex> l(c->ix->table_mu);
    for (size_t i = 0; i < c->ix->tables.size(); i++) {
      TgKmerTable& t = c->ix->tables[i];
      if (t.slots != c->slots) continue;
      if (--t.refs == 0) { cudaFree(t.slots); c->ix->tables.erase(c->ix->tables.begin() + (long)i); }
      break;
    }
  }
  if (c->d_ctr) cudaFree(c->d_ctr);
  if (c->h_ctr) cudaFreeHost(c->h_ctr);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev2) cudaEventDestroy(c->ev2);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int i = 0; i < TG_MAX_ROUNDS; i++) { if (c->ev_dp0[i]) cudaEventDestroy(c->ev_dp0[i]); if (c->ev_dp1[i]) cudaEventDestroy(c->ev_dp1[i]); }
  for (int i = 0; i < 4; i++) { if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]); if (c->side[i]) cudaStreamDestroy(c->side[i]); }
  if (c->h_active) cudaFreeHost(c->h_active);
  if (c->h_snap) cudaFreeHost(c->h_snap);
  for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// ThermiteAligner::align_read from many threads (src/wrapper.rs:20-27, :72): the micro-batcher of host_batcher.cpp over
// tg_align_batch of this context
tg_status tg_batcher_create(tg_ctx* ctx, uint32_t max_batch_reads, uint32_t max_wait_us, tg_batcher** out) {
  TG_GUARD_BEGIN
  if (!ctx || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  return tg_batcher_create_backend(
      [](void* user, const uint8_t* bases, const uint64_t* offs, uint32_t n, tg_result* res) {
        return tg_align_batch((tg_ctx*)user, bases, offs, n, res);
      },
      ctx, max_batch_reads, max_wait_us, out);
  TG_GUARD_END
}

tg_status tg_ctx_create(const tg_index* ix, const tg_opts* opts, tg_ctx** out) {
  TG_GUARD_BEGIN
  if (!ix || !opts || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  if (opts->min_seed_len < 1 || opts->min_seed_len > TG_MAX_SEED_LEN)
    return tg_fail(TG_ERR_INVALID, "min_seed_len must be in [1, TG_MAX_SEED_LEN]");
  if (!(opts->min_aln_score_percent >= 0.0f && opts->min_aln_score_percent <= 1.0f))
    return tg_fail(TG_ERR_INVALID, "Min alignment score percent must be between 0.0 and 1.0!");  // src/main.rs:46-49
  CU_CHECK(cudaSetDevice(ix->device));
  auto* c = new tg_ctx();
  if (const char* e = getenv("TG_SMALL_BATCH")) c->small_batch = (uint32_t)atol(e);
  if (const char* e = getenv("TG_TINY_BATCH")) c->tiny_batch = (uint32_t)atol(e);  // tests: 0 = every batch takes the large-batch path
  if (const char* e = getenv("TG_ITEM_SORT")) c->item_sort = atoi(e);  // experiments: rounds with locus-ordered items (0 = off)
  c->ix = ix;
  c->opts = *opts;
  auto fail = [&](tg_status s) { tg_ctx_destroy(c); return s; };
#define CTX_CHECK(call)                                                                                        \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess) return fail(tg_fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
  } while (0)
  CTX_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CTX_CHECK(cudaEventCreate(&c->ev0));
  CTX_CHECK(cudaEventCreate(&c->ev1));
  CTX_CHECK(cudaEventCreate(&c->ev2));
  CTX_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 4; i++) {
    CTX_CHECK(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
    CTX_CHECK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
  }
  CTX_CHECK(cudaMallocHost(&c->h_active, sizeof(unsigned long long)));
  CTX_CHECK(cudaMallocHost(&c->h_snap, 2 * sizeof(unsigned long long)));
  CTX_CHECK(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  CTX_CHECK(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  for (int i = 0; i < TG_MAX_ROUNDS; i++) { CTX_CHECK(cudaEventCreate(&c->ev_dp0[i])); CTX_CHECK(cudaEventCreate(&c->ev_dp1[i])); }
  CTX_CHECK(cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, ix->device));
  CTX_CHECK(cudaMalloc(&c->d_ctr, sizeof(DevCounters)));
  CTX_CHECK(cudaMallocHost(&c->h_ctr, sizeof(DevCounters)));
  CTX_CHECK(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  // k-mer table for this min_seed_len: shared with the other contexts of this index, else built on the device from the
  // suffix array (sized from the number of distinct k-mers: <= 50 % load)
  {
    std::lock_guard<std::recursive_mutex> l(ix->table_mu);
    for (TgKmerTable& t : ix->tables)
      if (t.k == opts->min_seed_len) { c->slots = t.slots; c->n_slots = t.n_slots; t.refs++; break; }
    if (!c->slots) {
      const uint64_t T = ix->dev.text_len;
      const int threads = 256;
      const int blocks = (int)std::min<uint64_t>((T + threads - 1) / threads, (uint64_t)c->n_sms * 32);
      k_kmer_count<<<blocks, threads, 0, c->stream>>>(ix->dev.text4, T, ix->dev.sa, opts->min_seed_len, c->d_ctr);
      CTX_CHECK(cudaGetLastError());
      CTX_CHECK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
      CTX_CHECK(cudaStreamSynchronize(c->stream));
      uint64_t groups = c->h_ctr->kmer_groups;
      uint64_t n_slots = 1024;
      while (n_slots < 2 * groups + 2) n_slots <<= 1;
      TgSlot* slots = nullptr;
      CTX_CHECK(cudaMalloc(&slots, n_slots * sizeof(TgSlot)));
      c->slots = slots; c->n_slots = n_slots;
      ix->tables.push_back(TgKmerTable{opts->min_seed_len, slots, n_slots, 1});  // (from here on tg_ctx_destroy releases it)
      CTX_CHECK(cudaMemsetAsync(c->slots, 0, n_slots * sizeof(TgSlot), c->stream));
      k_kmer_insert<<<blocks, threads, 0, c->stream>>>(ix->dev.text4, T, ix->dev.sa, opts->min_seed_len, c->slots, n_slots - 1);
      CTX_CHECK(cudaGetLastError());
      CTX_CHECK(cudaStreamSynchronize(c->stream));
    }
  }
#undef CTX_CHECK
  *out = c;
  return TG_OK;
  TG_GUARD_END
}

tg_status tg_ctx_set_result_buffers(tg_ctx* ctx, int n) {
  if (!ctx || (n != 1 && n != 2)) return tg_fail(TG_ERR_INVALID, "result buffers: 1 or 2");
  ctx->result_sets = n;
  return TG_OK;
}
void tg_ctx_set_chunk_reads(tg_ctx* ctx, uint32_t reads) {
  if (ctx) ctx->chunk_reads = reads == 0 ? 0 : (reads < 1024 ? 1024 : reads);
}
int tg_ctx_device(const tg_ctx* ctx) { return ctx ? ctx->ix->device : -1; }
void* tg_ctx_stream(tg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void tg_ctx_last_kernel_ms(const tg_ctx* ctx, float* seed_ms, float* extend_ms) {
  if (seed_ms) *seed_ms = ctx ? ctx->last_seed_ms : 0.f;
  if (extend_ms) *extend_ms = ctx ? ctx->last_extend_ms : 0.f;
}
void tg_ctx_set_round_pipeline(tg_ctx* ctx, int on) {
  if (ctx) ctx->use_rounds = on ? 1 : 0;
}
void tg_ctx_set_exact_cell_count(tg_ctx* ctx, int on) {
  if (ctx) ctx->exact_cells = on ? 1 : 0;
}
void tg_ctx_debug_phases(const tg_ctx* ctx, uint64_t* out16) {
  for (int k = 0; k < 16; k++) out16[k] = ctx ? ctx->h_ctr->phase[k] : 0;
}
// debug: out[0..3] = items, hops words, complex reads, fin words; then round_end[16], round_tasks[16], round_ops[16]
void tg_ctx_debug_rounds(const tg_ctx* ctx, uint64_t* out52) {
  const DevCounters* h = ctx->h_ctr;
  out52[0] = h->items_used; out52[1] = h->hops_used; out52[2] = h->n_complex; out52[3] = h->fin_used;
  for (int r = 0; r < TG_MAX_ROUNDS; r++) {
    out52[4 + r] = h->round_active[r]; out52[4 + TG_MAX_ROUNDS + r] = h->round_tasks[r]; out52[4 + 2 * TG_MAX_ROUNDS + r] = h->round_ops[r];
  }
}
void tg_ctx_debug_classes(const tg_ctx* ctx, uint32_t* out) {  // [TG_MAX_ROUNDS][TG_DPT_NCLS]
  memcpy(out, ctx->h_ctr->round_cls, sizeof(ctx->h_ctr->round_cls));
}
float tg_ctx_last_dp_ms(const tg_ctx* ctx) { return ctx ? ctx->last_dp_ms : 0.f; }
// debug: rounds run by the last batch and the device time of each round's DP section
int tg_ctx_debug_round_dp_ms(const tg_ctx* ctx, float* out16) {
  for (int r = 0; r < TG_MAX_ROUNDS; r++) out16[r] = r < ctx->rounds_run ? ctx->round_dp_ms[r] : 0.f;
  return ctx->rounds_run;
}
// Random 16-B gather rate over the context's own k-mer table (GB/s of 32-B sectors), best of `reps`.
tg_status tg_bench_random_gather(tg_ctx* c, uint64_t n_loads, int reps, double* sector_gbs, float* best_ms) {
  TG_GUARD_BEGIN
  if (!c || !sector_gbs) return tg_fail(TG_ERR_INVALID, "null argument");
  CU_CHECK(cudaSetDevice(c->ix->device));
  uint32_t* sink = nullptr;
  CU_CHECK(cudaMalloc(&sink, 64));
  const uint32_t per_thread = 16;
  const uint64_t threads = (n_loads + per_thread - 1) / per_thread;
  const int blocks = (int)((threads + 255) / 256);
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
    CU_CHECK(cudaEventRecord(c->ev0, c->stream));
    k_random_gather<<<blocks, 256, 0, c->stream>>>((const uint4*)c->slots, c->n_slots - 1, per_thread, sink);
    CU_CHECK(cudaEventRecord(c->ev1, c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CU_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(sink);
  *sector_gbs = 32.0 * (double)(threads * per_thread) / (best / 1e3) / 1e9;
  if (best_ms) *best_ms = best;
  return TG_OK;
  TG_GUARD_END
}
// ALU-pipe yardstick: lane-operations per second of dependency-free VIADDMNMX and VIMNMX3 streams (best of `reps`).
tg_status tg_bench_int_peak(tg_ctx* c, int reps, double* viaddmnmx_lane_ops, double* vimnmx3_lane_ops) {
  TG_GUARD_BEGIN
  if (!c || !viaddmnmx_lane_ops || !vimnmx3_lane_ops) return tg_fail(TG_ERR_INVALID, "null argument");
  CU_CHECK(cudaSetDevice(c->ix->device));
  int* sink = nullptr;
  CU_CHECK(cudaMalloc(&sink, 64));
  const int iters = 4096, blocks = c->n_sms * 8;
  for (int kind = 0; kind < 2; kind++) {
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
      CU_CHECK(cudaEventRecord(c->ev0, c->stream));
      if (kind == 0) k_int_peak<0><<<blocks, 256, 0, c->stream>>>(iters, -1, 12345, sink);
      else k_int_peak<1><<<blocks, 256, 0, c->stream>>>(iters, -1, 12345, sink);
      CU_CHECK(cudaEventRecord(c->ev1, c->stream));
      CU_CHECK(cudaStreamSynchronize(c->stream));
      float ms = 0.f;
      CU_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
      if (r > 0 && ms < best) best = ms;
    }
    const double ops = (double)iters * 16.0 * 256.0 * blocks / (best / 1e3);
    if (kind == 0) *viaddmnmx_lane_ops = ops; else *vimnmx3_lane_ops = ops;
  }
  cudaFree(sink);
  return TG_OK;
  TG_GUARD_END
}
uint64_t tg_ctx_last_kernel_launches(const tg_ctx* ctx) { return ctx ? ctx->n_launches : 0; }
uint64_t tg_ctx_kmer_table_bytes(const tg_ctx* ctx) { return ctx ? ctx->n_slots * sizeof(TgSlot) : 0; }

}  // extern "C"

namespace {

// compact output: the device pools hold tg_aln_c records / u32 firsts (same buffers, smaller elements)
void set_compact_out(tg_ctx* c, TgAlignOut& out) {
  if (!c->compact) return;
  out.alns_c = (tg_aln_c*)c->d_alns.p;
  out.first32 = (uint32_t*)c->d_aln_first.p + c->out_row0;
  out.first_base = c->first_base; out.ops_base = c->ops_base;
}

// seeding of the reads [r0, r0 + nk) of a batch of n reads (the whole batch when nk == n)
tg_status launch_seed(tg_ctx* c, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n, uint32_t maxL, uint32_t r0 = 0,
                      uint32_t nk = 0xFFFFFFFFu) {
  if (nk == 0xFFFFFFFFu) nk = n;
  SeedParams p;
  p.bases = d_bases; p.offs = d_offs + r0; p.n_reads = nk; p.k = c->opts.min_seed_len; p.max_len = maxL;
  p.max_q = maxL >= p.k ? maxL - p.k + 1 : 1;
  p.rp_words = maxL / 16 + 4;
  tg_status st;
  if ((st = c->r_rp.ensure((size_t)n * p.rp_words * 8)) != TG_OK) return st;
  if ((st = c->d_probe.ensure((size_t)n * p.max_q * sizeof(TgSeedHit))) != TG_OK) return st;
  p.slots = c->slots; p.slot_mask = c->n_slots - 1; p.text4 = c->ix->dev.text4; p.sa = c->ix->dev.sa;
  p.rp = (uint64_t*)c->r_rp.p + (size_t)r0 * p.rp_words; p.hits = (TgSeedHit*)c->d_probe.p + (size_t)r0 * p.max_q; p.wave = 0;
  p.out.pool = (tg_seed*)c->d_seeds.p; p.out.pool_used = &c->d_ctr->seed_used; p.out.pool_cap = c->seed_cap;
  p.out.read_first = (uint64_t*)c->d_seed_first.p + r0; p.out.read_count = (uint32_t*)c->d_seed_count.p + r0;
  p.out.flags = &c->d_ctr->flags; p.out.n_smems = &c->d_ctr->n_smems;
  p.ctr = c->d_ctr;
  if ((st = c->d_queue2.ensure((size_t)nk * p.max_q * 8 + 64)) != TG_OK) return st;
  if ((st = c->d_queue.ensure((size_t)nk * p.max_q * 8 + 64)) != TG_OK) return st;
  p.queue = (uint64_t*)c->d_queue.p; p.queue2 = (uint64_t*)c->d_queue2.p;
  const int grid = c->n_sms * 8;
  k_pack_reads<<<grid, 256, 0, c->stream>>>(p);
  for (int wave = 0; wave < (p.max_q > 1 ? 3 : 1); wave++) {
    p.wave = wave;
    if (r0 != 0) CU_CHECK(cudaMemsetAsync(&c->d_ctr->probe_n[wave][0], 0, 16, c->stream));  // a later chunk of the batch
    if (wave == 2) { k_probe_brackets<<<grid, 256, 0, c->stream>>>(p); c->n_launches++; }
    k_probe_light<<<grid, 256, 0, c->stream>>>(p);
    k_probe_heavy<<<grid, 256, 0, c->stream>>>(p);
    c->n_launches += 2;
  }
  c->n_launches -= 1;
  k_seed_select<<<(int)std::min<uint64_t>(((uint64_t)nk + 127) / 128, (uint64_t)c->n_sms * 16), 128, 0, c->stream>>>(p);
  c->n_launches += 3;
  CU_CHECK(cudaGetLastError());
  return TG_OK;
}

tg_status launch_extend(tg_ctx* c, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n, uint32_t maxL,
                        const uint32_t* read_list = nullptr) {
  uint32_t max_bw = band_for(c->opts, maxL);
  uint32_t max_xlen = maxL > 0 ? maxL - 1 : 0;
  uint32_t max_cols = max_xlen + max_bw + 1;
  uint32_t trace_bytes = max_cols * (uint32_t)tg_trace_bytes_per_col((int)max_xlen, 32);
  uint32_t ops_cap = 2 * maxL + max_bw + 16;
  ExtSmemLayout lay = ext_smem_layout(maxL, max_cols, trace_bytes, ops_cap);
  int wpc = TG_WARPS_PER_CTA;
  while (wpc > 1 && lay.total * wpc > 100 * 1024) wpc >>= 1;  // long reads: fewer warps per CTA
  size_t smem = lay.total * wpc;
  if (smem > 227 * 1024) return tg_fail(TG_ERR_CAPACITY, "reads too long for the extension kernel's shared memory");
  const int rcls = tg_swg_rows_class((int)max_xlen, 32);
  void (*kern)(ExtParams) = rcls <= 3 ? k_extend<3> : rcls <= 6 ? k_extend<6> : k_extend<16>;
  CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem));
  if (occ < 1) return tg_fail(TG_ERR_CAPACITY, "extension kernel does not fit in shared memory");
  int blocks = (int)std::min<uint64_t>((uint64_t)c->n_sms * occ, ((uint64_t)n + wpc - 1) / wpc);
  if (blocks < 1) blocks = 1;
  uint32_t warps = (uint32_t)blocks * wpc;
  const uint32_t arena_cap = 65536;
  if (warps > c->scratch_warps) {
    tg_status st = c->d_cands.ensure((size_t)warps * TG_MAX_ALNS_PER_READ * sizeof(TgCand));
    if (st != TG_OK) return st;
    st = c->d_arena.ensure((size_t)warps * arena_cap * 4);
    if (st != TG_OK) return st;
    st = c->d_order.ensure((size_t)warps * 2 * TG_MAX_ALNS_PER_READ * 2);
    if (st != TG_OK) return st;
    c->scratch_warps = warps;
  }
  ExtParams p;
  p.P.ix = c->ix->dev; p.P.opts = c->opts;
  p.bases = d_bases; p.offs = d_offs; p.n_reads = n; p.max_len = maxL; p.max_cols = max_cols; p.trace_bytes = trace_bytes;
  p.ops_cap = ops_cap;
  p.seeds = (const tg_seed*)c->d_seeds.p; p.read_seed_first = (const uint64_t*)c->d_seed_first.p;
  p.read_seed_count = (const uint32_t*)c->d_seed_count.p;
  p.cands = (TgCand*)c->d_cands.p; p.arena = (uint32_t*)c->d_arena.p; p.arena_cap = arena_cap; p.order = (uint16_t*)c->d_order.p;
  p.bound_stop = c->exact_cells ? 0 : 1;
  p.read_list = read_list;
  p.out.read_aln_first = (uint64_t*)c->d_aln_first.p + c->out_row0; p.out.read_aln_count = (uint32_t*)c->d_aln_count.p + c->out_row0;
  p.out.alns = (tg_aln*)c->d_alns.p; p.out.ops = (uint32_t*)c->d_ops.p;
  p.out.alns_used = &c->d_ctr->alns_used; p.out.ops_used = &c->d_ctr->ops_used;
  p.out.alns_cap = c->alns_cap; p.out.ops_cap = c->ops_cap; p.out.flags = &c->d_ctr->flags;
  set_compact_out(c, p.out);
  p.ctr = c->d_ctr;
  kern<<<blocks, wpc * 32, smem, c->stream>>>(p);
  c->n_launches++;
  CU_CHECK(cudaGetLastError());
  return TG_OK;
}

// Geometry and scratch of the thread-per-extension kernels: one grid per class group, trace regions for `max_cols`
// columns, the sorted task index.
typedef void (*DptKernel)(RoundParams);
static DptKernel dpt_kernel(int cls) {
  switch (cls) {
    case 1: return k_round_dpt<1>;
    case 2: return k_round_dpt<2>;
    case 3: return k_round_dpt<3>;
    case 4: return k_round_dpt<4>;
    case 5: return k_round_dpt<5>;
    case 6: return k_round_dpt<6>;
    case 7: return k_round_dpt<7>;
    case 8: return k_round_dpt<8>;
    case 9: return k_round_dpt<9>;
    case 10: return k_round_dpt<10>;
    default: return k_round_dpt<11>;
  }
}
tg_status dpt_geometry(tg_ctx* c, RoundParams& p, uint32_t max_cols, uint64_t task_cap, int grid[TG_DPT_NCLS]) {
  tg_status st;
  size_t off = 0;
  grid[0] = 0; p.dpt_trace_off[0] = 0; p.dpt_trace_words[0] = 0;
  for (int g = 1; g < TG_DPT_NCLS; g++) {
    if (c->dpt_occ[g] == 0) {
      int o = 0;
      CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dpt_kernel(g), 128, 0));
      if (o < 1) return tg_fail(TG_ERR_INTERNAL, "thread DP kernel does not fit");
      c->dpt_occ[g] = o;
    }
    grid[g] = c->n_sms * c->dpt_occ[g];
    p.dpt_trace_off[g] = off;
    p.dpt_trace_words[g] = (size_t)max_cols * ((2 * tg_dpt_wb(g) + 31) / 32) * 32;  // per warp: [column][word][lane]
    off += (size_t)grid[g] * 4 * p.dpt_trace_words[g];
  }
  if ((st = c->r_dpt_trace.ensure(off * 4)) != TG_OK) return st;
  if ((st = c->r_sorted.ensure(task_cap * 4)) != TG_OK) return st;
  p.dpt_trace = (uint32_t*)c->r_dpt_trace.p;
  p.dpt_one = 1; p.dpt_k128 = 128;
  p.sorted = (uint32_t*)c->r_sorted.p;
  return TG_OK;
}

// sort the tasks of round p.round by band class and run the thread-per-extension kernels (one per class, widest first) and
// the warp-cooperative kernel (`warp_launch`, for what the thread kernels do not take) concurrently on the side streams
template <class WarpLaunch>
tg_status launch_dpt(tg_ctx* c, RoundParams& p, const int grid[TG_DPT_NCLS], WarpLaunch&& warp_launch) {
  k_round_hist<<<c->n_sms * 4, 256, 0, c->stream>>>(p);
  k_round_binscan<<<1, TG_BINSCAN_THREADS, 0, c->stream>>>(p);
  k_round_scatter<<<c->n_sms * 4, 256, 0, c->stream>>>(p);
  if (p.all_warp) {  // a tiny batch: one launch on the main stream, no fork / join
    warp_launch(c->stream);
    c->n_launches += 4;
    return TG_OK;
  }
  CU_CHECK(cudaEventRecord(c->ev_fork, c->stream));
  for (int i = 0; i < 4; i++) CU_CHECK(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
  warp_launch(c->side[3]);
  cudaStream_t lanes[4] = {c->side[0], c->side[1], c->side[2], c->stream};
  for (int cls = TG_DPT_NCLS - 1, k = 0; cls >= 1; cls--, k++) dpt_kernel(cls)<<<grid[cls], 128, 0, lanes[k & 3]>>>(p);
  for (int i = 0; i < 4; i++) {
    CU_CHECK(cudaEventRecord(c->ev_join[i], c->side[i]));
    CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
  }
  c->n_launches += 4 + (TG_DPT_NCLS - 1);
  return TG_OK;
}

tg_status launch_rounds(tg_ctx* c, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n, uint32_t maxL) {
  tg_status st;
  const uint32_t rp_words = maxL / 16 + 4;
  if (c->round_task_cap < (uint64_t)n * 4 + 65536) c->round_task_cap = (uint64_t)n * 4 + 65536;
  if (c->round_ops_cap < (uint64_t)n * 16 + 65536) c->round_ops_cap = (uint64_t)n * 16 + 65536;
  if (c->item_cap < (uint64_t)n * 3 + 65536) c->item_cap = (uint64_t)n * 3 + 65536;
  if (c->hops_cap < (uint64_t)n * 64 + (1u << 20)) c->hops_cap = (uint64_t)n * 64 + (1u << 20);
  if (c->item_cap > 0xFFFFFFF0ull) return tg_fail(TG_ERR_CAPACITY, "too many seed hits in one batch: split the batch");
  if ((st = c->r_state.ensure((size_t)n * sizeof(TgReadState))) != TG_OK) return st;
  if ((st = c->r_hits.ensure(c->item_cap * sizeof(TgHit))) != TG_OK) return st;
  if ((st = c->r_ires.ensure(c->item_cap * sizeof(TgItemRes))) != TG_OK) return st;
  if ((st = c->r_cands.ensure(c->item_cap * sizeof(TgCand))) != TG_OK) return st;
  if ((st = c->r_hops.ensure(c->hops_cap * 4)) != TG_OK) return st;
  if ((st = c->r_fin.ensure(c->item_cap * 12)) != TG_OK) return st;
  if ((st = c->r_tasks.ensure(c->round_task_cap * sizeof(TgTask))) != TG_OK) return st;
  if ((st = c->r_ops.ensure(c->round_ops_cap * 4)) != TG_OK) return st;
  if ((st = c->r_complex.ensure((size_t)n * 4 + 16)) != TG_OK) return st;
  if ((st = c->r_perm.ensure(c->item_cap * 4)) != TG_OK) return st;
  if ((st = c->r_ikey.ensure(c->item_cap * 4)) != TG_OK) return st;
  if ((st = c->r_ibins.ensure((size_t)2 * TG_IB_N * 4)) != TG_OK) return st;
  const uint32_t max_bw = band_for(c->opts, maxL);
  const uint32_t max_xlen = maxL > 0 ? maxL - 1 : 0;
  const uint32_t max_cols = max_xlen + max_bw + 1;
  const uint32_t trace_bytes = (max_cols + 1) * (uint32_t)tg_trace_bytes_per_col((int)max_xlen, 32);
  const uint32_t ops_words = max_xlen + max_cols + 8;
  RoundParams p;
  p.P.ix = c->ix->dev; p.P.opts = c->opts;
  p.bases = d_bases; p.offs = d_offs; p.n_reads = n; p.max_len = maxL; p.rp_words = rp_words;
  p.seeds = (const tg_seed*)c->d_seeds.p; p.read_seed_first = (const uint64_t*)c->d_seed_first.p;
  p.read_seed_count = (const uint32_t*)c->d_seed_count.p;
  p.st = (TgReadState*)c->r_state.p; p.rp = (uint64_t*)c->r_rp.p;
  p.hits = (TgHit*)c->r_hits.p; p.ires = (TgItemRes*)c->r_ires.p; p.cands = (TgCand*)c->r_cands.p; p.item_cap = c->item_cap;
  p.hp.w = (uint32_t*)c->r_hops.p; p.hp.used = &c->d_ctr->hops_used; p.hp.cap = c->hops_cap;
  p.fin = (uint32_t*)c->r_fin.p; p.fin_cap = c->item_cap * 3;
  p.tasks = (TgTask*)c->r_tasks.p; p.task_cap = c->round_task_cap;
  p.ops_pool = (uint32_t*)c->r_ops.p; p.ops_cap = c->round_ops_cap;
  p.complex_list = (uint32_t*)c->r_complex.p; p.round = 0; p.early = 0;
  p.late = nullptr;
  p.perm = nullptr; p.ikey = (uint32_t*)c->r_ikey.p; p.ibins = (uint32_t*)c->r_ibins.p;
  p.ib_shift = 0;
  while (((uint64_t)TG_IB_N << p.ib_shift) < c->ix->dev.text_len) p.ib_shift++;
  if (c->item_sort > 0) CU_CHECK(cudaMemsetAsync(c->r_ibins.p, 0, (size_t)TG_IB_N * 4, c->stream));
  c->early_rows = 0;
  if (c->early_out) {
    if ((st = c->r_late.ensure((size_t)n * 16 + 64)) != TG_OK) return st;
    p.late = (ulonglong2*)c->r_late.p;
  }
  p.max_xlen = max_xlen; p.max_cols = max_cols; p.trace_bytes = trace_bytes; p.ops_words = ops_words;
  p.bound_stop = c->exact_cells ? 0 : 1;
  p.out.read_aln_first = (uint64_t*)c->d_aln_first.p + c->out_row0; p.out.read_aln_count = (uint32_t*)c->d_aln_count.p + c->out_row0;
  p.out.alns = (tg_aln*)c->d_alns.p; p.out.ops = (uint32_t*)c->d_ops.p;
  p.out.alns_used = &c->d_ctr->alns_used; p.out.ops_used = &c->d_ctr->ops_used;
  p.out.alns_cap = c->alns_cap; p.out.ops_cap = c->ops_cap; p.out.flags = &c->d_ctr->flags;
  set_compact_out(c, p.out);
  p.ctr = c->d_ctr;
  // extension kernel geometry
  int wpc = TG_WARPS_PER_CTA;
  const size_t per_warp = swg_smem_per_warp(max_xlen, max_cols, trace_bytes, ops_words);
  while (wpc > 1 && per_warp * wpc > 100 * 1024) wpc >>= 1;
  const size_t smem = per_warp * wpc;
  if (smem > 227 * 1024) return tg_fail(TG_ERR_CAPACITY, "reads too long for the extension kernel's shared memory");
  const int rcls = tg_swg_rows_class((int)max_xlen, 32);
  void (*kdp)(RoundParams) = rcls <= 3 ? k_round_dp<3> : rcls <= 6 ? k_round_dp<6> : k_round_dp<16>;
  CU_CHECK(cudaFuncSetAttribute(kdp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kdp, wpc * 32, smem));
  if (occ < 1) return tg_fail(TG_ERR_CAPACITY, "extension kernel does not fit in shared memory");
  const int dp_blocks = c->n_sms * occ;
  const int tblocks = (int)std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)c->n_sms * 16);
  const int iblocks = c->n_sms * 16;
  // thread-per-extension kernel: geometry, trace scratch, sorted task list
  int dpt_grid[TG_DPT_NCLS];
  if ((st = dpt_geometry(c, p, std::min<uint32_t>(max_xlen, TG_DPT_MAX_X) + max_bw + 1, c->round_task_cap, dpt_grid)) != TG_OK) return st;
  // a tiny batch cannot fill even one class of the thread kernels: all of its extensions run on the warp-cooperative
  // kernel and the eleven per-class launches (with their fork / join across streams) are not made at all
  p.all_warp = n < c->tiny_batch ? 1 : 0;
  k_round_init<<<tblocks, 128, 0, c->stream>>>(p);
  c->n_launches++;
  // a small batch (per-read callers behind tg_batcher) is all launch latency: no item sort, and the host looks for "nothing
  // left" one round earlier (98 % of the reads are finished after round 1; a round is 13 launches, a check one sync)
  const bool small = n < c->small_batch;
  const uint32_t check_from = n <= 2 ? 0u : small ? 1u : 2u;
  for (uint32_t r = 0; r < TG_MAX_ROUNDS; r++) {
    p.round = r;
    k_round_plan<<<tblocks, 128, 0, c->stream>>>(p);
    p.perm = nullptr;
    if ((int)r < c->item_sort && !small) {  // the big rounds: hand the items to prep / post in locus order
      p.perm = (uint32_t*)c->r_perm.p;
      k_round_ikey<<<c->n_sms * 8, 256, 0, c->stream>>>(p);
      k_round_iscan<<<1, 1024, 0, c->stream>>>(p);
      k_round_iscatter<<<c->n_sms * 8, 256, 0, c->stream>>>(p);
      c->n_launches += 3;
    }
    k_round_prep<<<iblocks, 128, 0, c->stream>>>(p);
    CU_CHECK(cudaEventRecord(c->ev_dp0[r], c->stream));
    if ((st = launch_dpt(c, p, dpt_grid, [&](cudaStream_t s2) { kdp<<<dp_blocks, wpc * 32, smem, s2>>>(p); })) != TG_OK) return st;
    CU_CHECK(cudaEventRecord(c->ev_dp1[r], c->stream));
    c->rounds_run = (int)r + 1;
    k_round_post<<<iblocks, 128, 0, c->stream>>>(p);
    k_round_scan<<<tblocks, 128, 0, c->stream>>>(p);
    c->n_launches += 4;
    if (r <= 1 && c->early_out && (c->early_out > 1 || r == 1)) {
      // ~55 % of the reads are finished after round 0 and ~98 % after round 1: write their records now and move them to
      // the host while the next rounds run (the records of round 0 travel under round 1, the heaviest one)
      p.early = 1;
      k_round_final<<<tblocks, 128, 0, c->stream>>>(p);
      p.early = 0;
      c->n_launches++;
      k_host_snapshot<<<1, 32, 0, c->stream>>>(c->h_snap, &c->d_ctr->alns_used, 2);
      CU_CHECK(cudaStreamSynchronize(c->stream));
      const unsigned long long a1 = std::min<unsigned long long>(c->h_snap[0], c->alns_cap), o1 = std::min<unsigned long long>(c->h_snap[1], c->ops_cap);
      const unsigned long long a0 = c->early_alns, o0 = c->early_ops;
      const size_t rec = c->compact ? sizeof(tg_aln_c) : sizeof(tg_aln), fsz = c->compact ? 4 : 8;
      bool room = true;
      if (!c->ho.external) {
        // room for the whole batch, estimated from what is known (grown again at the end if a batch needs more)
        const size_t want_a = (size_t)std::max<unsigned long long>(a1 + a1 / 16, (unsigned long long)n + n / 8) + 65536;
        const size_t want_o = (size_t)std::max<unsigned long long>(o1 + o1 / 16, r == 0 ? 3 * o1 : 0ull) + 262144;
        if (want_a * rec > c->h_alns.cap || want_o * 4 > c->h_ops.cap) {
          CU_CHECK(cudaStreamSynchronize(c->copy_out));  // what was sent so far has landed before the buffers move
          if ((st = c->h_alns.ensure_keep(want_a * rec, (size_t)a0 * rec)) != TG_OK) return st;
          if ((st = c->h_ops.ensure_keep(want_o * 4, (size_t)o0 * 4)) != TG_OK) return st;
        }
        c->ho.alns = (uint8_t*)c->h_alns.p; c->ho.ops = (uint32_t*)c->h_ops.p;
        c->ho.alns_cap = c->h_alns.cap / rec; c->ho.ops_cap = c->h_ops.cap / 4;
      } else if (a1 > c->ho.alns_cap || o1 > c->ho.ops_cap) {
        room = false;  // the caller's segment is too small: the call will end with the sizes it needs
        c->ho_overflow = true;
      }
      if (room && !c->ho_overflow) {
        if (a1 > a0)
          CU_CHECK(cudaMemcpyAsync(c->ho.alns + a0 * rec, (const uint8_t*)c->d_alns.p + a0 * rec, (size_t)(a1 - a0) * rec, cudaMemcpyDeviceToHost, c->copy_out));
        if (o1 > o0)
          CU_CHECK(cudaMemcpyAsync(c->ho.ops + o0, (uint32_t*)c->d_ops.p + o0, (size_t)(o1 - o0) * 4, cudaMemcpyDeviceToHost, c->copy_out));
        c->early_alns = a1; c->early_ops = o1;
        if (r == 1) {  // first/count of every read that is finished by now; the others follow as a fix-up list
          CU_CHECK(cudaMemcpyAsync(c->ho.first, c->d_aln_first.p, (size_t)n * fsz, cudaMemcpyDeviceToHost, c->copy_out));
          CU_CHECK(cudaMemcpyAsync(c->ho.count, c->d_aln_count.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_out));
          c->early_rows = 1;
        }
      }
    }
    // late rounds are short: a host check for "nothing left" costs less than launching the remaining empty rounds
    if (r >= check_from && r + 1 < TG_MAX_ROUNDS) {
      k_host_snapshot<<<1, 32, 0, c->stream>>>(c->h_active, &c->d_ctr->round_active[r], 1);
      CU_CHECK(cudaStreamSynchronize(c->stream));
      if (*c->h_active == 0) break;
    }
  }
  k_round_final<<<tblocks, 128, 0, c->stream>>>(p);
  c->n_launches++;
  CU_CHECK(cudaGetLastError());
  // whatever the rounds could not finish (too many hits / transcripts per seed) runs on the single-warp path
  return launch_extend(c, d_bases, d_offs, n, maxL, (const uint32_t*)c->r_complex.p);
}

tg_status ensure_pools(tg_ctx* c, uint32_t n) {
  const uint32_t np = std::max(n, c->pool_reads);  // output pools hold the whole batch when it is processed in chunks
  if (c->seed_cap < (uint64_t)n * 4 + 4096) c->seed_cap = (uint64_t)n * 4 + 4096;
  if (c->alns_cap < (uint64_t)np * 2 + 4096) c->alns_cap = (uint64_t)np * 2 + 4096;
  if (c->ops_cap < (uint64_t)np * 24 + 65536) c->ops_cap = (uint64_t)np * 24 + 65536;
  tg_status st;
  if ((st = c->d_seeds.ensure(c->seed_cap * sizeof(tg_seed))) != TG_OK) return st;
  if ((st = c->d_seed_first.ensure((size_t)n * 8 + 8)) != TG_OK) return st;
  if ((st = c->d_seed_count.ensure((size_t)n * 4 + 4)) != TG_OK) return st;
  if ((st = c->d_aln_first.ensure((size_t)np * 8 + 8)) != TG_OK) return st;
  if ((st = c->d_aln_count.ensure((size_t)np * 4 + 4)) != TG_OK) return st;
  if ((st = c->d_alns.ensure(c->alns_cap * sizeof(tg_aln))) != TG_OK) return st;
  if ((st = c->d_ops.ensure(c->ops_cap * 4)) != TG_OK) return st;
  return TG_OK;
}

// seeds (+ optionally extension) with pool-overflow retry; leaves the counters in c->h_ctr
tg_status run_pipeline(tg_ctx* c, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n, uint32_t maxL, bool extend) {
  for (int attempt = 0; attempt < 8; attempt++) {
    tg_status st = ensure_pools(c, n);
    if (st != TG_OK) return st;
    CU_CHECK(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
    if (c->base_alns || c->base_ops) {  // a later chunk of a batch: its records continue the pools (alns_used, ops_used adjacent)
      static_assert(offsetof(DevCounters, ops_used) == offsetof(DevCounters, alns_used) + 8, "counter layout");
      const unsigned long long bases2[2] = {c->base_alns, c->base_ops};
      CU_CHECK(cudaMemcpyAsync(&c->d_ctr->alns_used, bases2, 16, cudaMemcpyHostToDevice, c->stream));
    }
    CU_CHECK(cudaEventRecord(c->ev0, c->stream));
    c->early_alns = 0; c->early_ops = 0;
    if (c->in_chunks > 1 && attempt == 0) {
      for (uint32_t k = 0; k < c->in_chunks; k++) {  // seed chunk k as soon as its bases have arrived
        const uint32_t r0 = k * c->in_chunk_reads, nk = std::min(n, r0 + c->in_chunk_reads) - r0;
        CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
        if ((st = launch_seed(c, d_bases, d_offs, n, maxL, r0, nk)) != TG_OK) return st;
      }
    } else {
      if (c->in_chunks && attempt == 0) CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_in[0], 0));
      if ((st = launch_seed(c, d_bases, d_offs, n, maxL)) != TG_OK) return st;
    }
    CU_CHECK(cudaEventRecord(c->ev1, c->stream));
    if (extend && (st = (c->use_rounds ? launch_rounds(c, d_bases, d_offs, n, maxL) : launch_extend(c, d_bases, d_offs, n, maxL))) != TG_OK) return st;
    CU_CHECK(cudaEventRecord(c->ev2, c->stream));
    CU_CHECK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream));
    CU_CHECK(cudaEventElapsedTime(&c->last_seed_ms, c->ev0, c->ev1));
    CU_CHECK(cudaEventElapsedTime(&c->last_extend_ms, c->ev1, c->ev2));
    c->last_dp_ms = 0.f;
    if (extend && c->use_rounds)
      for (int r = 0; r < c->rounds_run; r++) {
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, c->ev_dp0[r], c->ev_dp1[r]));
        c->last_dp_ms += ms;
        c->round_dp_ms[r] = ms;
      }
    int f = c->h_ctr->flags;
    if (f & (TG_FLAG_SEED_POOL | TG_FLAG_ALN_POOL | TG_FLAG_OPS_POOL | TG_FLAG_TASK_POOL | TG_FLAG_ITEM_POOL | TG_FLAG_HOPS_POOL)) {  // grow the pool that overflowed and redo the batch
      if (f & TG_FLAG_TASK_POOL) c->round_task_cap *= 2;
      if (f & TG_FLAG_ITEM_POOL) c->item_cap = std::max<uint64_t>(c->item_cap * 2, c->h_ctr->items_used + 65536);
      if (f & TG_FLAG_HOPS_POOL) c->hops_cap = std::max<uint64_t>(c->hops_cap * 2, c->h_ctr->hops_used + 65536);
      if (f & TG_FLAG_SEED_POOL) c->seed_cap = std::max<uint64_t>(c->seed_cap * 2, c->h_ctr->seed_used + 4096);
      if (f & TG_FLAG_ALN_POOL) c->alns_cap = std::max<uint64_t>(c->alns_cap * 2, c->h_ctr->alns_used + 4096);
      if (f & TG_FLAG_OPS_POOL) {
        c->ops_cap = std::max<uint64_t>(c->ops_cap * 2, c->h_ctr->ops_used + 4096);
        c->round_ops_cap *= 2;
      }
      continue;
    }
    return check_flags(f);
  }
  return tg_fail(TG_ERR_CAPACITY, "result pools kept overflowing");
}

tg_status check_batch_args(tg_ctx* ctx, const void* a, const void* b, const void* out) {
  if (!ctx || !out || !a || !b) return tg_fail(TG_ERR_INVALID, "null argument");
  cudaError_t e = cudaSetDevice(ctx->ix->device);
  if (e != cudaSuccess) return tg_fail(TG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  return TG_OK;
}

tg_status upload_reads(tg_ctx* c, const uint8_t* bases, const uint64_t* offs, uint32_t n, uint32_t* maxL_out) {
  uint32_t maxL = 0;
  for (uint32_t r = 0; r < n; r++) {
    if (offs[r + 1] < offs[r]) return tg_fail(TG_ERR_INVALID, "read offsets must be non-decreasing");
    uint64_t L = offs[r + 1] - offs[r];
    if (L > TG_MAX_READ_LEN) return tg_fail(TG_ERR_INVALID, "read longer than TG_MAX_READ_LEN");
    if (L > maxL) maxL = (uint32_t)L;
  }
  uint64_t total = offs[n];
  tg_status st;
  if ((st = c->d_bases.ensure(total + 64)) != TG_OK) return st;
  if ((st = c->d_offs.ensure((size_t)(n + 1) * 8)) != TG_OK) return st;
  if (total) CU_CHECK(cudaMemcpyAsync(c->d_bases.p, bases, total, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->d_offs.p, offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  *maxL_out = maxL;
  return TG_OK;
}

void fill_result(tg_ctx* c, uint32_t n, tg_result* out) {
  out->n_reads = n;
  out->n_alns = c->h_ctr->alns_used;
  out->n_ops = c->h_ctr->ops_used;
  out->swg_cells = c->h_ctr->cells;
  out->swg_extensions = c->h_ctr->n_ext;
  out->seed_hits = c->h_ctr->hits;
  out->n_smems = c->h_ctr->n_smems;
}

}  // namespace

extern "C" {

tg_status tg_align_batch_device(tg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n_reads,
                                uint64_t total_bases, uint32_t max_read_len, tg_result* out) {
  TG_GUARD_BEGIN
  tg_status st = check_batch_args(ctx, d_bases, d_offs, out);
  if (st != TG_OK) return st;
  (void)total_bases;
  if (max_read_len > TG_MAX_READ_LEN) return tg_fail(TG_ERR_INVALID, "read longer than TG_MAX_READ_LEN");
  memset(out, 0, sizeof(*out));
  if (n_reads == 0) return TG_OK;
  ctx->n_launches = 0;
  ctx->compact = 0; ctx->first_base = 0; ctx->ops_base = 0;
  if ((st = run_pipeline(ctx, d_bases, d_offs, n_reads, std::max(max_read_len, 1u), true)) != TG_OK) return st;
  fill_result(ctx, n_reads, out);
  out->read_aln_first = (const uint64_t*)ctx->d_aln_first.p;
  out->read_aln_count = (const uint32_t*)ctx->d_aln_count.p;
  out->alns = (const tg_aln*)ctx->d_alns.p;
  out->ops = (const uint32_t*)ctx->d_ops.p;
  return TG_OK;
  TG_GUARD_END
}

// tg_align_batch / tg_align_batch_compact / one shard of tg_multi_align_batch.  offs[0] need not be 0 (a shard of a larger
// batch): reads are bases[offs[r], offs[r+1]).  ext: caller-owned result segment (compact only), else the context's buffers.
static tg_status align_host(tg_ctx* c, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, bool compact,
                            const tg_ctx::HostOut* ext, unsigned long long first_base, unsigned long long ops_base) {
  tg_status st;
  const bool dbg = getenv("TG_DEBUG_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_in = now();
  const size_t rec = compact ? sizeof(tg_aln_c) : sizeof(tg_aln), fsz = compact ? 4 : 8;
  c->compact = compact ? 1 : 0;
  c->first_base = first_base; c->ops_base = ops_base;
  c->ho_overflow = false; c->need_alns = 0; c->need_ops = 0;
  if (c->result_sets == 2 && !ext) {  // the previous call's result stays where it is; this call fills the other set
    std::swap(c->h_first, c->h2_first); std::swap(c->h_count, c->h2_count);
    std::swap(c->h_alns, c->h2_alns); std::swap(c->h_ops, c->h2_ops);
  }
  // Copies overlap kernels at both ends of the call (three streams): the bases arrive in chunks and every chunk is
  // seeded as soon as it has landed; the records of the reads that are finished after round 1 (~98 %) travel to the
  // host while the late rounds run.  One pass over the whole batch, one contiguous result.
  // measured on B200: ~4-8 chunks hide the input copy best; smaller chunks lose more in the seeding kernels than they hide
  const uint32_t want = c->chunk_reads ? c->chunk_reads : std::max<uint32_t>(262144u, (n_reads + 7) / 8);
  const uint32_t chunk = n_reads >= 2 * (uint64_t)want ? want : n_reads;
  const uint32_t n_chunks = (n_reads + chunk - 1) / chunk;
  const uint64_t base0 = offs[0];
  if (offs[n_reads] < base0) return tg_fail(TG_ERR_INVALID, "read offsets must be non-decreasing");
  const uint64_t total = offs[n_reads] - base0;
  if ((st = c->d_bases.ensure(total + 64)) != TG_OK) return st;
  if ((st = c->d_offs.ensure((size_t)(n_reads + 1) * 8)) != TG_OK) return st;
  while (c->ev_in.size() < n_chunks) {
    cudaEvent_t e;
    CU_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_in.push_back(e);
  }
  CU_CHECK(cudaMemcpyAsync(c->d_offs.p, offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, c->copy_in));
  uint32_t maxL = 1;
  // Offsets are checked before the bytes they describe are touched: one streaming pass (8 B per read), the first chunk by
  // this thread, the others meanwhile by the context's two helpers (large batches only), so that the first copy is on its
  // way after an eighth of the pass and the kernels can be launched after a third of it.
  auto fail_offsets = [&](const uint64_t* o, uint32_t m) {
    bool decreasing = false;
    for (uint32_t i = 0; i < m; i++) decreasing |= o[i + 1] < o[i];
    cudaStreamSynchronize(c->copy_in);
    return tg_fail(TG_ERR_INVALID, decreasing ? "read offsets must be non-decreasing" : "read longer than TG_MAX_READ_LEN");
  };
  const bool helpers = n_chunks >= 2 && n_reads >= (1u << 20);
  struct HelperGuard {  // the helpers read the caller's offsets: never leave this call while they are still at it
    OffsCheck* oc;
    ~HelperGuard() { if (oc) { uint64_t w; uint32_t l; oc->wait(w, l); } }
  } helper_guard{helpers ? &c->offs_check : nullptr};
  if (helpers) c->offs_check.submit(offs + chunk, n_reads - chunk);
  for (uint32_t k = 0; k < n_chunks; k++) {
    const uint32_t r0 = k * chunk, r1 = std::min(n_reads, r0 + chunk);
    if (k == 0 || !helpers) {
      OffsCheck::Job j;
      j.o = offs + r0; j.m = r1 - r0;
      OffsCheck::scan(j);
      if (j.wide || j.longest > TG_MAX_READ_LEN) {
        if (helpers) { uint64_t w; uint32_t l; c->offs_check.wait(w, l); }
        return fail_offsets(offs + r0, r1 - r0);
      }
      maxL = std::max<uint32_t>(maxL, j.longest);
    } else if (k == 1) {  // everything behind the first chunk was checked by the helpers
      uint64_t w; uint32_t l;
      c->offs_check.wait(w, l);
      if (w || l > TG_MAX_READ_LEN) return fail_offsets(offs + chunk, n_reads - chunk);
      maxL = std::max<uint32_t>(maxL, l);
    }
    const uint64_t b0 = offs[r0], b1 = offs[r1];
    if (b1 > b0) CU_CHECK(cudaMemcpyAsync((uint8_t*)c->d_bases.p + (b0 - base0), bases + b0, b1 - b0, cudaMemcpyHostToDevice, c->copy_in));
    CU_CHECK(cudaEventRecord(c->ev_in[k], c->copy_in));
  }
  // maxL of the whole batch is needed before the first kernel: the loop above has run over every read by now
  if (ext) c->ho = *ext;
  else {
    if ((st = c->h_first.ensure((size_t)n_reads * 8)) != TG_OK) return st;
    if ((st = c->h_count.ensure((size_t)n_reads * 4)) != TG_OK) return st;
    c->ho.first = (uint8_t*)c->h_first.p; c->ho.count = (uint32_t*)c->h_count.p;
    c->ho.alns = (uint8_t*)c->h_alns.p; c->ho.ops = (uint32_t*)c->h_ops.p;
    c->ho.alns_cap = c->h_alns.cap / rec; c->ho.ops_cap = c->h_ops.cap / 4;
    c->ho.external = false;
  }
  const double t_issued = now();
  c->n_launches = 0;
  c->in_chunks = n_chunks; c->in_chunk_reads = chunk;
  {
    const char* em = getenv("TG_EARLY_MODE");  // experiments: 0 no early output, 1 after round 1 only, 2 (default) after rounds 0 and 1
    c->early_out = c->use_rounds && n_reads >= c->small_batch ? (em ? atoi(em) : 2) : 0;  // (small batch: nothing to hide)
  }
  // the kernels address the reads through the caller's offsets: hand them the device buffer shifted by offs[0]
  st = run_pipeline(c, (const uint8_t*)c->d_bases.p - base0, (const uint64_t*)c->d_offs.p, n_reads, maxL, true);
  c->in_chunks = 0; c->early_out = 0;
  const double t_pipe = now();
  if (st != TG_OK) { cudaStreamSynchronize(c->copy_out); return st; }
  const uint64_t n_alns = c->h_ctr->alns_used, n_ops = c->h_ctr->ops_used;
  const uint64_t ea = std::min<uint64_t>(c->early_alns, n_alns), eo = std::min<uint64_t>(c->early_ops, n_ops);
  if (!c->ho.external) {
    if (n_alns * rec + 16 > c->h_alns.cap || n_ops * 4 + 16 > c->h_ops.cap) {
      CU_CHECK(cudaStreamSynchronize(c->copy_out));  // the early part has landed before the buffers move
      if ((st = c->h_alns.ensure_keep((size_t)n_alns * rec + 16, ea * rec)) != TG_OK) return st;
      if ((st = c->h_ops.ensure_keep((size_t)n_ops * 4 + 16, eo * 4)) != TG_OK) return st;
    }
    c->ho.alns = (uint8_t*)c->h_alns.p; c->ho.ops = (uint32_t*)c->h_ops.p;
  } else if (c->ho_overflow || n_alns > c->ho.alns_cap || n_ops > c->ho.ops_cap) {
    c->ho_overflow = true; c->need_alns = n_alns; c->need_ops = n_ops;
    c->early_rows = 0;
    CU_CHECK(cudaStreamSynchronize(c->copy_out));
    return TG_OK;  // (the caller looks at ho_overflow)
  }
  // first/count: already on their way for the reads finished after round 1; the last pass lists the rest (a few per cent).
  // Reads redone by the single-warp kernel are not in that list: then everything is copied again.
  const uint64_t n_late = c->h_ctr->n_late;
  const bool fixup = c->early_rows && c->h_ctr->n_complex == 0;
  c->early_rows = 0;
  if (fixup) {
    if ((st = c->h_late.ensure((size_t)n_late * 16 + 16)) != TG_OK) return st;
    if (n_late) CU_CHECK(cudaMemcpyAsync(c->h_late.p, c->r_late.p, (size_t)n_late * 16, cudaMemcpyDeviceToHost, c->stream));
  } else {
    CU_CHECK(cudaStreamSynchronize(c->copy_out));  // (an early copy of the same arrays must not land after this one)
    CU_CHECK(cudaMemcpyAsync(c->ho.first, c->d_aln_first.p, (size_t)n_reads * fsz, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(cudaMemcpyAsync(c->ho.count, c->d_aln_count.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if (n_alns > ea)
    CU_CHECK(cudaMemcpyAsync(c->ho.alns + ea * rec, (const uint8_t*)c->d_alns.p + ea * rec, (size_t)(n_alns - ea) * rec,
                             cudaMemcpyDeviceToHost, c->stream));
  if (n_ops > eo)
    CU_CHECK(cudaMemcpyAsync(c->ho.ops + eo, (uint32_t*)c->d_ops.p + eo, (size_t)(n_ops - eo) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  const double t_sync1 = now();
  CU_CHECK(cudaStreamSynchronize(c->copy_out));
  const double t_sync2 = now();
  if (fixup) c->offs_check.fix_up((const unsigned long long*)c->h_late.p, n_late, c->ho.first, c->ho.count, compact);
  c->last_wall_ms = now() - t_in;
  if (dbg)
    fprintf(stderr, "tg_align_batch tail: late copies %.2f ms, early copies still in flight %.2f ms, fix-up of %llu reads %.2f ms\n",
            t_sync1 - t_pipe, t_sync2 - t_sync1, (unsigned long long)n_late, now() - t_sync2);
  if (dbg)
    fprintf(stderr, "tg_align_batch[dev %d]: %u reads, issue inputs %.2f ms, pipeline %.2f ms (seed %.2f + extend %.2f on the device), tail %.2f ms; "
            "H2D %.1f MB, D2H %.1f MB (%.1f MB early)\n", c->ix->device, n_reads,
            t_issued - t_in, t_pipe - t_issued, c->last_seed_ms, c->last_extend_ms, now() - t_pipe,
            (total + 8.0 * (n_reads + 1)) / 1e6, (n_alns * rec + n_ops * 4.0 + n_reads * (fsz + 4.0)) / 1e6, (ea * rec + eo * 4.0) / 1e6);
  return TG_OK;
}

tg_status tg_align_batch(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result* out) {
  TG_GUARD_BEGIN
  tg_status st = check_batch_args(ctx, offs, offs, out);
  if (st != TG_OK) return st;
  memset(out, 0, sizeof(*out));
  if (n_reads == 0) return TG_OK;
  if (!bases && offs[n_reads] > offs[0]) return tg_fail(TG_ERR_INVALID, "null argument");
  if ((st = align_host(ctx, bases, offs, n_reads, false, nullptr, 0, 0)) != TG_OK) return st;
  fill_result(ctx, n_reads, out);
  out->read_aln_first = (const uint64_t*)ctx->ho.first;
  out->read_aln_count = ctx->ho.count;
  out->alns = (const tg_aln*)ctx->ho.alns;
  out->ops = ctx->ho.ops;
  return TG_OK;
  TG_GUARD_END
}

static void fill_result_c(const tg_ctx* c, uint32_t n, tg_result_c* out) {
  out->n_reads = n; out->n_segments = 1;
  out->n_alns = c->h_ctr->alns_used; out->n_ops = c->h_ctr->ops_used;
  out->alns_extent = out->n_alns; out->ops_extent = out->n_ops;
  out->swg_cells = c->h_ctr->cells; out->swg_extensions = c->h_ctr->n_ext;
  out->seed_hits = c->h_ctr->hits; out->n_smems = c->h_ctr->n_smems;
}

tg_status tg_align_batch_compact(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result_c* out) {
  TG_GUARD_BEGIN
  tg_status st = check_batch_args(ctx, offs, offs, out);
  if (st != TG_OK) return st;
  memset(out, 0, sizeof(*out));
  if (n_reads == 0) return TG_OK;
  if (!bases && offs[n_reads] > offs[0]) return tg_fail(TG_ERR_INVALID, "null argument");
  if ((st = align_host(ctx, bases, offs, n_reads, true, nullptr, 0, 0)) != TG_OK) return st;
  fill_result_c(ctx, n_reads, out);
  out->read_aln_first = (const uint32_t*)ctx->ho.first;
  out->read_aln_count = ctx->ho.count;
  out->alns = (const tg_aln_c*)ctx->ho.alns;
  out->ops = ctx->ho.ops;
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"

tg_status tg_ctx_align_segment(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, const TgHostSegment& seg,
                               TgShardStat* stat) {
  TG_GUARD_BEGIN
  memset(stat, 0, sizeof(*stat));
  if (n_reads == 0) return TG_OK;
  tg_status st = check_batch_args(ctx, offs, offs, stat);
  if (st != TG_OK) return st;
  tg_ctx::HostOut ho;
  ho.first = (uint8_t*)seg.first; ho.count = seg.count; ho.alns = (uint8_t*)seg.alns; ho.ops = seg.ops;
  ho.alns_cap = seg.alns_cap; ho.ops_cap = seg.ops_cap; ho.external = true;
  if ((st = align_host(ctx, bases, offs, n_reads, true, &ho, seg.first_base, seg.ops_base)) != TG_OK) return st;
  stat->overflow = ctx->ho_overflow;
  stat->need_alns = ctx->need_alns; stat->need_ops = ctx->need_ops;
  stat->n_alns = ctx->h_ctr->alns_used; stat->n_ops = ctx->h_ctr->ops_used;
  stat->swg_cells = ctx->h_ctr->cells; stat->swg_extensions = ctx->h_ctr->n_ext;
  stat->seed_hits = ctx->h_ctr->hits; stat->n_smems = ctx->h_ctr->n_smems;
  stat->wall_ms = ctx->last_wall_ms; stat->seed_ms = ctx->last_seed_ms; stat->extend_ms = ctx->last_extend_ms; stat->dp_ms = ctx->last_dp_ms;
  return TG_OK;
  TG_GUARD_END
}

extern "C" {

tg_status tg_seed_batch(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_seed_result* out) {
  TG_GUARD_BEGIN
  tg_status st = check_batch_args(ctx, offs, offs, out);
  if (st != TG_OK) return st;
  memset(out, 0, sizeof(*out));
  if (n_reads == 0) return TG_OK;
  if (!bases && offs[n_reads] > 0) return tg_fail(TG_ERR_INVALID, "null argument");
  uint32_t maxL = 0;
  if ((st = upload_reads(ctx, bases, offs, n_reads, &maxL)) != TG_OK) return st;
  ctx->n_launches = 0;
  if ((st = run_pipeline(ctx, (const uint8_t*)ctx->d_bases.p, (const uint64_t*)ctx->d_offs.p, n_reads, std::max(maxL, 1u), false)) != TG_OK)
    return st;
  uint64_t ns = ctx->h_ctr->seed_used;
  if ((st = ctx->h_seed_first.ensure((size_t)n_reads * 8)) != TG_OK) return st;
  if ((st = ctx->h_seed_count.ensure((size_t)n_reads * 4)) != TG_OK) return st;
  if ((st = ctx->h_seeds.ensure((size_t)ns * sizeof(tg_seed) + 16)) != TG_OK) return st;
  CU_CHECK(cudaMemcpyAsync(ctx->h_seed_first.p, ctx->d_seed_first.p, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(cudaMemcpyAsync(ctx->h_seed_count.p, ctx->d_seed_count.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (ns) CU_CHECK(cudaMemcpyAsync(ctx->h_seeds.p, ctx->d_seeds.p, (size_t)ns * sizeof(tg_seed), cudaMemcpyDeviceToHost, ctx->stream));
  CU_CHECK(cudaStreamSynchronize(ctx->stream));
  out->n_reads = n_reads;
  out->n_seeds = ns;
  out->read_seed_first = (const uint64_t*)ctx->h_seed_first.p;
  out->read_seed_count = (const uint32_t*)ctx->h_seed_count.p;
  out->seeds = (const tg_seed*)ctx->h_seeds.p;
  return TG_OK;
  TG_GUARD_END
}

tg_status tg_swg_extend_batch(tg_ctx* c, const uint8_t* xs, const uint64_t* xoff, const uint8_t* ys, const uint64_t* yoff,
                              uint32_t n, const uint32_t* band_width, const int32_t* x_drop, int32_t* score, uint32_t* xend,
                              uint32_t* yend, uint64_t* ops_off, uint32_t* ops, uint64_t ops_cap, uint64_t* cells,
                              float* kernel_ms) {
  TG_GUARD_BEGIN
  if (!c || !xoff || !yoff || !band_width || !x_drop || !score || !xend || !yend || !ops_off)
    return tg_fail(TG_ERR_INVALID, "null argument");
  CU_CHECK(cudaSetDevice(c->ix->device));
  if (cells) *cells = 0;
  if (kernel_ms) *kernel_ms = 0.f;
  if (n == 0) { ops_off[0] = 0; return TG_OK; }
  uint32_t max_xlen = 0, max_cols = 1, dpt_cols = 1, dpt_x = 1, n_dpt = 0;
  uint64_t worst_ops = 0, ysym_total = 0;
  std::vector<uint64_t> ysym(n);
  for (uint32_t t = 0; t < n; t++) {
    uint64_t xl = xoff[t + 1] - xoff[t], yl = yoff[t + 1] - yoff[t];
    if (xl > TG_MAX_READ_LEN) return tg_fail(TG_ERR_INVALID, "x longer than TG_MAX_READ_LEN");
    if (band_width[t] > 2 * TG_MAX_READ_LEN) return tg_fail(TG_ERR_INVALID, "band width too large");
    if (x_drop[t] < (int32_t)band_width[t])
      return tg_fail(TG_ERR_INVALID, "x_drop < band_width: the reference reads stale columns or panics there (unsupported)");
    uint64_t cols = std::min<uint64_t>(yl, xl + band_width[t]);
    max_xlen = std::max<uint32_t>(max_xlen, (uint32_t)xl);
    max_cols = std::max<uint32_t>(max_cols, (uint32_t)cols);
    worst_ops += xl + cols + 2;
    ysym[t] = ysym_total;
    if (xl && yl && tg_dpt_class((int)xl, (int)band_width[t], x_drop[t]) > 0) {  // may run on the thread kernels
      dpt_cols = std::max<uint32_t>(dpt_cols, (uint32_t)cols);
      dpt_x = std::max<uint32_t>(dpt_x, (uint32_t)xl);
      ysym_total += ((cols + 15) / 16 + 2) * 16;
      n_dpt++;
    }
  }
  uint32_t trace_bytes = (max_cols + 1) * (uint32_t)tg_trace_bytes_per_col((int)max_xlen, 32);
  uint32_t ops_words = max_xlen + max_cols + 8;
  int wpc = TG_WARPS_PER_CTA;
  while (wpc > 1 && swg_smem_per_warp(max_xlen, max_cols, trace_bytes, ops_words) * wpc > 100 * 1024) wpc >>= 1;
  size_t smem = swg_smem_per_warp(max_xlen, max_cols, trace_bytes, ops_words) * wpc;
  if (smem > 227 * 1024) return tg_fail(TG_ERR_CAPACITY, "sequences too long for the SWG kernel's shared memory");
  tg_status st;
  uint64_t nx = xoff[n], ny = yoff[n];
#define ENS(buf, bytes) if ((st = c->buf.ensure(bytes)) != TG_OK) return st
  ENS(s_x, nx + 64); ENS(s_xo, (size_t)(n + 1) * 8); ENS(s_y, ny + 64); ENS(s_yo, (size_t)(n + 1) * 8);
  ENS(s_bw, (size_t)n * 4); ENS(s_xd, (size_t)n * 4); ENS(s_score, (size_t)n * 4); ENS(s_xe, (size_t)n * 4);
  ENS(s_ye, (size_t)n * 4); ENS(s_toff, (size_t)n * 8); ENS(s_tlen, (size_t)n * 4); ENS(s_ops, worst_ops * 4 + 64);
  const uint32_t rp_words = dpt_x / 16 + 4;
  ENS(r_tasks, (size_t)n * sizeof(TgTask)); ENS(r_rp, (size_t)n * rp_words * 8); ENS(s_ypk, (ysym_total / 16 + 4) * 8);
  ENS(s_ysym, (size_t)n * 8); ENS(r_ops, worst_ops * 4 + 64);
#undef ENS
  CU_CHECK(cudaMemcpyAsync(c->s_ysym.p, ysym.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  if (nx) CU_CHECK(cudaMemcpyAsync(c->s_x.p, xs, nx, cudaMemcpyHostToDevice, c->stream));
  if (ny) CU_CHECK(cudaMemcpyAsync(c->s_y.p, ys, ny, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->s_xo.p, xoff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->s_yo.p, yoff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->s_bw.p, band_width, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->s_xd.p, x_drop, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemsetAsync(c->d_ctr, 0, sizeof(DevCounters), c->stream));
  const int rcls = tg_swg_rows_class((int)max_xlen, 32);
  void (*kern)(SwgParams) = rcls <= 3 ? k_swg_batch<3> : rcls <= 6 ? k_swg_batch<6> : k_swg_batch<16>;
  CU_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem));
  if (occ < 1) return tg_fail(TG_ERR_CAPACITY, "SWG kernel does not fit in shared memory");
  int blocks = (int)std::min<uint64_t>((uint64_t)c->n_sms * occ, ((uint64_t)n + wpc - 1) / wpc);
  SwgParams p;
  p.xs = (const uint8_t*)c->s_x.p; p.xoff = (const uint64_t*)c->s_xo.p; p.ys = (const uint8_t*)c->s_y.p; p.yoff = (const uint64_t*)c->s_yo.p;
  p.n = n; p.bw = (const uint32_t*)c->s_bw.p; p.x_drop = (const int32_t*)c->s_xd.p;
  p.score = (int32_t*)c->s_score.p; p.xend = (uint32_t*)c->s_xe.p; p.yend = (uint32_t*)c->s_ye.p;
  p.task_off = (uint64_t*)c->s_toff.p; p.task_len = (uint32_t*)c->s_tlen.p; p.ops = (uint32_t*)c->s_ops.p; p.ops_cap = worst_ops;
  p.max_xlen = max_xlen; p.max_cols = max_cols; p.trace_bytes = trace_bytes; p.ops_words = ops_words; p.ctr = c->d_ctr;
  p.bound_stop = c->exact_cells ? 0 : 1;
  p.list = nullptr;
  p.tasks = (TgTask*)c->r_tasks.p; p.xpk = (uint64_t*)c->r_rp.p; p.ypk = (uint64_t*)c->s_ypk.p;
  p.ysym_off = (const uint64_t*)c->s_ysym.p; p.rp_words = rp_words; p.dp_ops = (const uint32_t*)c->r_ops.p;
  c->n_launches = 0;
  bool dp_timed = false;
  CU_CHECK(cudaEventRecord(c->ev0, c->stream));
  if (n_dpt == 0 || !c->use_rounds) {
    kern<<<blocks, wpc * 32, smem, c->stream>>>(p);
    c->n_launches++;
  } else {
    // pairs of A/C/G/N/T symbols that fit the register-band kernels: round-pipeline tasks on k_round_dpt<1..11>; the rest
    // (other bytes, very long x, very wide bands, empty inputs) on the warp kernel with raw bytes
    RoundParams rp{};
    rp.P.ix = c->ix->dev; rp.P.ix.text4 = p.ypk; rp.P.opts = c->opts;
    rp.n_reads = n; rp.rp_words = rp_words; rp.rp = p.xpk;
    rp.tasks = p.tasks; rp.task_cap = n; rp.ops_pool = (uint32_t*)c->r_ops.p; rp.ops_cap = worst_ops; rp.round = 0;
    rp.bound_stop = p.bound_stop; rp.ctr = c->d_ctr;
    int grid[TG_DPT_NCLS];
    if ((st = dpt_geometry(c, rp, dpt_cols + 1, n, grid)) != TG_OK) return st;
    const unsigned long long n_tasks = n;
    CU_CHECK(cudaMemcpyAsync(&c->d_ctr->round_tasks[0], &n_tasks, sizeof(n_tasks), cudaMemcpyHostToDevice, c->stream));
    const int tb = (int)std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)c->n_sms * 16);
    k_swg_prepare<<<tb, 128, 0, c->stream>>>(p);
    p.list = rp.sorted;
    CU_CHECK(cudaEventRecord(c->ev_dp0[0], c->stream));  // the DP kernels alone (tg_ctx_last_dp_ms): without pack / sort / collect
    if ((st = launch_dpt(c, rp, grid, [&](cudaStream_t s2) { kern<<<blocks, wpc * 32, smem, s2>>>(p); })) != TG_OK) return st;
    CU_CHECK(cudaEventRecord(c->ev_dp1[0], c->stream));
    dp_timed = true;
    k_swg_collect<<<tb, 128, 0, c->stream>>>(p);
    c->n_launches += 2;
  }
  CU_CHECK(cudaGetLastError());
  CU_CHECK(cudaEventRecord(c->ev1, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaMemcpyAsync(score, c->s_score.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaMemcpyAsync(xend, c->s_xe.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaMemcpyAsync(yend, c->s_ye.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  if (c->h_ctr->flags) return tg_fail(TG_ERR_INTERNAL, "SWG ops pool overflow");
  if (kernel_ms) CU_CHECK(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
  if (dp_timed) CU_CHECK(cudaEventElapsedTime(&c->last_dp_ms, c->ev_dp0[0], c->ev_dp1[0]));
  else CU_CHECK(cudaEventElapsedTime(&c->last_dp_ms, c->ev0, c->ev1));
  if (cells) *cells = c->h_ctr->cells;
  // reorder the ops pool into task order on the host
  uint64_t used = c->h_ctr->swg_ops_used;
  std::vector<uint64_t> toff(n);
  std::vector<uint32_t> tlen(n), pool(used);
  CU_CHECK(cudaMemcpy(toff.data(), c->s_toff.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  CU_CHECK(cudaMemcpy(tlen.data(), c->s_tlen.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (used) CU_CHECK(cudaMemcpy(pool.data(), c->s_ops.p, used * 4, cudaMemcpyDeviceToHost));
  uint64_t o = 0;
  for (uint32_t t = 0; t < n; t++) {
    ops_off[t] = o;
    if (ops && o + tlen[t] <= ops_cap) memcpy(ops + o, pool.data() + toff[t], (size_t)tlen[t] * 4);
    o += tlen[t];
  }
  ops_off[n] = o;
  if (ops && o > ops_cap) return tg_fail(TG_ERR_CAPACITY, "ops buffer too small");
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"
