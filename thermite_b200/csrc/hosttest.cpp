// hosttest.cpp -- TEST-ONLY host build of tg_core.h (libtg_hosttest.so).
//
// There is no GPU in the development container, so the warp-cooperative code of tg_core.h is also compiled
// by g++ and driven by an emulated warp: either 1 lane (plain serial) or 32 lanes on 32 threads that meet
// at barriers for every shuffle / vote.  `tests/` uses this to check the kernel LOGIC against the oracle
// on CPU.  The product (thermite_b200/, libthermite_gpu.so) never loads this library: there is no CPU
// fallback in the product path.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include <random>

#include "tg_rounds.h"
#include "tg_dpt.h"

namespace {

struct HostWarp1 {
  static constexpr int LANES = 1;
  int lane() const { return 0; }
  int shfl_up(int v, int) { return v; }
  int shfl(int v, int) { return v; }
  unsigned long long shfl64(unsigned long long v, int) { return v; }
  bool any(bool p) { return p; }
  unsigned long long sum64(unsigned long long v) { return v; }
  uint32_t ballot(bool p) { return p ? 1u : 0u; }
  uint32_t reduce_min_u32(uint32_t v) { return v; }
  int reduce_max_i32(int v) { return v; }
  void sync() {}
  void sync_global() {}
  unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
  void atomic_or(int* p, int v) { *p |= v; }
};

struct Barrier {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    int g = gen.load();
    if (count.fetch_add(1) + 1 == n) {
      count.store(0);
      gen.fetch_add(1);
    } else {
      while (gen.load() == g) std::this_thread::yield();
    }
  }
};
struct Warp32Shared {
  Barrier bar{32};
  int xi[32];
  unsigned long long xl[32];
};
struct HostWarp32 {
  static constexpr int LANES = 32;
  int l;
  Warp32Shared* sh;
  int lane() const { return l; }
  int shfl_up(int v, int d) {
    sh->xi[l] = v;
    sh->bar.wait();
    int r = l >= d ? sh->xi[l - d] : v;
    sh->bar.wait();
    return r;
  }
  int shfl(int v, int src) {
    sh->xi[l] = v;
    sh->bar.wait();
    int r = sh->xi[src & 31];
    sh->bar.wait();
    return r;
  }
  unsigned long long shfl64(unsigned long long v, int src) {
    sh->xl[l] = v;
    sh->bar.wait();
    unsigned long long r = sh->xl[src & 31];
    sh->bar.wait();
    return r;
  }
  bool any(bool p) {
    sh->xi[l] = p;
    sh->bar.wait();
    bool r = false;
    for (int i = 0; i < 32; i++) r |= sh->xi[i] != 0;
    sh->bar.wait();
    return r;
  }
  uint32_t ballot(bool p) {
    sh->xi[l] = p;
    sh->bar.wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (sh->xi[i] ? 1u : 0u) << i;
    sh->bar.wait();
    return r;
  }
  int reduce_max_i32(int v) {
    sh->xi[l] = v;
    sh->bar.wait();
    int r = sh->xi[0];
    for (int i = 1; i < 32; i++) r = std::max(r, sh->xi[i]);
    sh->bar.wait();
    return r;
  }
  uint32_t reduce_min_u32(uint32_t v) {
    sh->xl[l] = v;
    sh->bar.wait();
    uint32_t r = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) r = std::min<uint32_t>(r, (uint32_t)sh->xl[i]);
    sh->bar.wait();
    return r;
  }
  unsigned long long sum64(unsigned long long v) {
    sh->xl[l] = v;
    sh->bar.wait();
    unsigned long long r = 0;
    for (int i = 0; i < 32; i++) r += sh->xl[i];
    sh->bar.wait();
    return r;
  }
  void sync() { sh->bar.wait(); }
  void sync_global() { sh->bar.wait(); }
  unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
  void atomic_or(int* p, int v) { *p |= v; }
};

TgIndexDev view_of(const tg_index_host* ix) {
  const TgBlobHeader* h = ix->hdr();
  const uint8_t* b = ix->blob.data();
  TgIndexDev d;
  d.text4 = (const uint64_t*)(b + h->off_text4);
  d.sa = (const uint32_t*)(b + h->off_sa);
  d.refs = (const TgRef*)(b + h->off_refs);
  d.exon_nodes = (const TgTreeNode*)(b + h->off_exon_nodes);
  d.gene_nodes = (const TgTreeNode*)(b + h->off_gene_nodes);
  d.exon_stab = (const TgStab*)(b + h->off_exon_stab);
  d.gene_stab = (const TgStab*)(b + h->off_gene_stab);
  d.n_exon_stab = (uint32_t)h->n_exon_nodes; d.n_gene_stab = (uint32_t)h->n_gene_nodes;
  d.exon_maxlen = (uint32_t)h->exon_maxlen; d.gene_maxlen = (uint32_t)h->gene_maxlen;
  d.tx_seq_off = (const uint64_t*)(b + h->off_tx_seq_off);
  d.tx_exon_off = (const uint32_t*)(b + h->off_tx_exon_off);
  d.te_start = (const uint32_t*)(b + h->off_te_start);
  d.te_end = (const uint32_t*)(b + h->off_te_end);
  d.txseq4 = (const uint64_t*)(b + h->off_txseq4);
  d.text_len = h->text_len;
  d.n_refs = (uint32_t)h->n_refs;
  d.n_txs = (uint32_t)h->n_txs;
  d.exon_root = (int32_t)h->exon_root;
  d.gene_root = (int32_t)h->gene_root;
  return d;
}

struct HostCtx {
  const tg_index_host* ix;
  TgIndexDev dev;
  tg_opts opts;
  std::vector<TgSlot> slots;
  uint64_t slot_mask;
};

struct WarpBuffers {
  std::vector<uint8_t> rd, xs, ys, trace;
  std::vector<uint32_t> a, b, c, t;
  std::vector<int32_t> stack;
  std::vector<uint64_t> rp, rp2;
  std::vector<TgSeedHit> hits;
  std::vector<tg_seed> sm;
  std::vector<uint16_t> grp;
  std::vector<TgCand> cands;
  std::vector<uint32_t> arena;
  TgWarpMem mem(uint32_t maxL, uint32_t max_bw, int lanes) {
    uint32_t max_cols = maxL + max_bw + 2;
    uint32_t cap = 2 * maxL + max_bw + 16;
    rd.assign(maxL + 16, 0); xs.assign(maxL + 16, 0); ys.assign(max_cols + 16, 0);
    size_t tb = (size_t)max_cols * tg_trace_bytes_per_col((int)maxL, lanes);
    trace.assign(tb + 64, 0);
    a.assign(cap, 0); b.assign(cap, 0); c.assign(cap, 0); t.assign(cap, 0);
    stack.assign(TG_TREE_STACK, 0);
    rp2.assign(maxL / 16 + 4, 0);
    return TgWarpMem{rd.data(), xs.data(), ys.data(), trace.data(), a.data(), b.data(), c.data(), t.data(), stack.data(), cap, rp2.data(), false};
  }
  TgSeedMem seed_mem(uint32_t maxL) {
    rp.assign(maxL / 16 + 4, 0); hits.assign(maxL + 1, TgSeedHit{0, 0, 0}); sm.assign(maxL + 1, tg_seed{}); grp.assign(maxL + 1, 0);
    return TgSeedMem{rp.data(), hits.data(), sm.data()};
  }
};

uint32_t max_bw_for(const tg_opts& o, uint32_t maxL) {
  (void)o;
  return maxL;  // bw = L - min_aln_score <= L
}

template <class F>
void run_lanes(int lanes, F&& f) {
  if (lanes == 1) {
    HostWarp1 w;
    f(w);
    return;
  }
  Warp32Shared sh;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++) th.emplace_back([&, l]() { HostWarp32 w{l, &sh}; f(w); });
  for (auto& t : th) t.join();
}

// ---- tg_dpt.h on the host: one "thread" (a pair of extensions), strides of 1 -----------------------------------------
struct DptScratch {
  std::vector<uint32_t> msk, tr;
  DptScratch() : msk(64, 0), tr((size_t)(TG_DPT_MAX_X + 2 * TG_MAX_READ_LEN + 2) * tg_dpt_twp(TG_DPT_MAX_WB), 0) {}
  TgDptMem mem() { return TgDptMem{msk.data(), 1, tr.data(), 1}; }
};
struct DptHalf {  // one extension of a pair: y stream, columns, x-drop; results
  TgDptY ys;
  int ncols, x_drop;
  std::vector<uint32_t> ops;
};
struct DptHalfX {  // where the shortcut finds x: packed sequence, offset, side
  const uint64_t* xseq = nullptr;
  uint32_t xoff = 0;
  int side = 0;
  uint64_t y0 = 0;
};
template <int WB>
void dpt_two(const TgDptMem& m, DptHalf& A, DptHalf& B, const DptHalfX& XA, const DptHalfX& XB, int xlen, int bw, bool bound_stop,
             TgDpt2Result& res) {
  tg_dpt2_fill<WB>(m, A.ys, B.ys, xlen, bw, A.ncols, B.ncols, A.x_drop, B.x_drop, bound_stop, res);
  // operations the way dpt_pair (thermite_gpu.cu) produces them: gapless shortcut where it applies, else traceback
  bool diag[2];
  uint64_t lo[2], hi[2];
  for (int h = 0; h < 2; h++) {
    const DptHalfX& X = h == 0 ? XA : XB;
    DptHalf& H = h == 0 ? A : B;
    diag[h] = tg_dpt_diag_mask(X.xseq, X.xoff, xlen, X.side, H.ys.seq, X.y0, res.xend[h], res.yend[h], res.score[h], lo[h], hi[h]);
    if (diag[h]) {
      const uint32_t n = tg_dpt_diag_emit(lo[h], hi[h], res.xend[h], xlen, h, [](int, uint32_t, uint32_t, uint32_t) {});
      H.ops.assign(n, 0);
      tg_dpt_diag_emit(lo[h], hi[h], res.xend[h], xlen, h, [&](int, uint32_t i, uint32_t kind, uint32_t run) { H.ops[i] = kind | (run << 3); });
    }
  }
  if (!diag[0] || !diag[1]) {
    uint32_t nA = 0, nB = 0;
    tg_dpt2_traceback<WB>(m, A.ys, B.ys, xlen, bw, res, !diag[0], !diag[1], nA, nB, [](int, uint32_t, uint32_t, uint32_t) {});
    if (!diag[0]) A.ops.assign(nA, 0);
    if (!diag[1]) B.ops.assign(nB, 0);
    tg_dpt2_traceback<WB>(m, A.ys, B.ys, xlen, bw, res, !diag[0], !diag[1], nA, nB,
                          [&](int h, uint32_t i, uint32_t kind, uint32_t run) { (h == 0 ? A : B).ops[i] = kind | (run << 3); });
  }
}
void dpt_dispatch(int cls, const TgDptMem& m, DptHalf& A, DptHalf& B, const DptHalfX& XA, const DptHalfX& XB, int xlen, int bw,
                  bool bound_stop, TgDpt2Result& res) {
  switch (cls) {
    case 1: dpt_two<4>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 2: dpt_two<8>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 3: dpt_two<16>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 4: dpt_two<24>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 5: dpt_two<32>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 6: dpt_two<40>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 7: dpt_two<48>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 8: dpt_two<56>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 9: dpt_two<64>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    case 10: dpt_two<72>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
    default: dpt_two<80>(m, A, B, XA, XB, xlen, bw, bound_stop, res); break;
  }
}
// Round tasks the way k_round_dpt runs them: sorted by (class, tg_dpt_subkey), neighbours with equal (xlen, band width)
// share a "thread", the others run paired with themselves.  Tasks of class 0 are left alone (returned as not done).
void dpt_tasks_host(const TgIndexDev& ix, const uint64_t* rp_base, uint32_t rp_words, TgTask* tasks, size_t n_tasks, bool bound_stop,
                    uint32_t* ops_pool, unsigned long long* ops_ctr, std::vector<uint8_t>& done) {
  done.assign(n_tasks, 0);
  std::vector<std::pair<uint32_t, uint32_t>> order;  // (bin, task)
  for (size_t t = 0; t < n_tasks; t++) {
    const int cls = tg_dpt_class((int)tasks[t].xlen, (int)tasks[t].bw, tasks[t].x_drop);
    if (cls == 0) continue;
    order.emplace_back((uint32_t)cls * TG_DPT_CBINS + tg_dpt_subkey((int)tasks[t].xlen, (int)tasks[t].bw), (uint32_t)t);
  }
  std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  DptScratch sc;
  auto run = [&](uint32_t i0, uint32_t i1, bool commit1) {
    TgTask& t0 = tasks[i0];
    TgTask& t1 = tasks[i1];
    const int xlen = (int)t0.xlen, bw = (int)t0.bw;
    const int nc0 = (int)t0.ylen < xlen + bw ? (int)t0.ylen : xlen + bw, nc1 = (int)t1.ylen < xlen + bw ? (int)t1.ylen : xlen + bw;
    TgDptMem m = sc.mem();
    DptHalf A, B;
    A.ys.init(tg_seq_of(ix, t0.seqsel), t0.y0, nc0, (int)t0.side); A.ncols = nc0; A.x_drop = t0.x_drop;
    B.ys.init(tg_seq_of(ix, t1.seqsel), t1.y0, nc1, (int)t1.side); B.ncols = nc1; B.x_drop = t1.x_drop;
    tg_dpt_profile(m, 0, rp_base + (size_t)t0.read * rp_words, t0.xoff, xlen, t0.side);
    tg_dpt_profile(m, 1, rp_base + (size_t)t1.read * rp_words, t1.xoff, xlen, t1.side);
    TgDpt2Result res;
    DptHalfX XA, XB;
    XA.xseq = rp_base + (size_t)t0.read * rp_words; XA.xoff = t0.xoff; XA.side = (int)t0.side; XA.y0 = t0.y0;
    XB.xseq = rp_base + (size_t)t1.read * rp_words; XB.xoff = t1.xoff; XB.side = (int)t1.side; XB.y0 = t1.y0;
    dpt_dispatch(tg_dpt_class(xlen, bw, t0.x_drop), m, A, B, XA, XB, xlen, bw, bound_stop, res);
    for (int h = 0; h < (commit1 ? 2 : 1); h++) {
      TgTask& t = h == 0 ? t0 : t1;
      const std::vector<uint32_t>& ops = h == 0 ? A.ops : B.ops;
      t.score = res.score[h]; t.xend = (uint32_t)res.xend[h]; t.yend = (uint32_t)res.yend[h]; t.cells = res.cells[h];
      t.ops_off = (uint32_t)*ops_ctr; t.ops_n = (uint32_t)ops.size();
      for (uint32_t w : ops) ops_pool[(*ops_ctr)++] = w;
      done[&t - tasks] = 1;
    }
  };
  for (size_t k = 0; k < order.size();) {
    // entries of one bin in pairs, the odd one alone -- the padded layout k_round_binscan builds
    size_t e = k;
    while (e < order.size() && order[e].first == order[k].first) e++;
    for (size_t u = k; u < e; u += 2) {
      const uint32_t ia = order[u].second;
      if (u + 1 < e) {
        const uint32_t ib = order[u + 1].second;
        const bool compat = tasks[ia].xlen == tasks[ib].xlen && tasks[ia].bw == tasks[ib].bw;
        if (compat) run(ia, ib, true);
        else { run(ia, ia, false); run(ib, ib, false); }
      } else run(ia, ia, false);
    }
    k = e;
  }
}

}  // namespace

extern "C" {

void* ht_ctx_create(const tg_index_host* ix, const tg_opts* opts) {
  auto* c = new HostCtx();
  c->ix = ix;
  c->dev = view_of(ix);
  c->opts = *opts;
  // k-mer table, same per-row logic as the device build kernel
  const uint32_t k = opts->min_seed_len;
  const uint64_t T = c->dev.text_len;
  uint64_t n_groups = 0;
  for (uint64_t r = 0; r < T; r++) {
    uint64_t w0, w1;
    if (tg_kmer_group_start(c->dev.text4, T, c->dev.sa, r, k, w0, w1)) n_groups++;
  }
  uint64_t n_slots = 1024;
  while (n_slots < 2 * n_groups + 2) n_slots <<= 1;
  c->slots.assign(n_slots, TgSlot{0, 0, 0, 0});
  c->slot_mask = n_slots - 1;
  for (uint64_t r = 0; r < T; r++) {
    uint64_t w0, w1;
    if (!tg_kmer_group_start(c->dev.text4, T, c->dev.sa, r, k, w0, w1)) continue;
    uint32_t cnt = tg_kmer_group_count(c->dev.text4, T, c->dev.sa, r, k, w0, w1);
    uint64_t h = tg_hash_kmer(w0, w1);
    uint64_t idx = h & c->slot_mask;
    while (c->slots[idx].tag != 0) idx = (idx + 1) & c->slot_mask;
    c->slots[idx] = TgSlot{tg_tag_of(h), cnt == 1 ? c->dev.sa[r] : (uint32_t)r, cnt, 0};
  }
  return c;
}
void ht_ctx_destroy(void* c) { delete (HostCtx*)c; }

// seeds out: caller-provided arrays sized generously; returns total seeds or -1
long long ht_seed_batch(void* cp, const uint8_t* bases, const uint64_t* offs, uint32_t n, int lanes, tg_seed* pool,
                        uint64_t pool_cap, uint64_t* read_first, uint32_t* read_count) {
  HostCtx* c = (HostCtx*)cp;
  uint32_t maxL = 1;
  for (uint32_t r = 0; r < n; r++) maxL = std::max<uint32_t>(maxL, (uint32_t)(offs[r + 1] - offs[r]));
  if (maxL > TG_MAX_READ_LEN) return -1;
  WarpBuffers wb;
  TgSeedMem sm = wb.seed_mem(maxL);
  unsigned long long used = 0, n_smems = 0;
  int flags = 0;
  TgSeedOut out{pool, &used, pool_cap, read_first, read_count, &flags, &n_smems};
  for (uint32_t r = 0; r < n; r++) {
    uint32_t L = (uint32_t)(offs[r + 1] - offs[r]);
    if (lanes == 0) {
      // the device pipeline: k_pack_reads -> k_seed_probe (offset 0 first, the rest unless the whole read matched) -> k_seed_select
      const uint32_t k = c->opts.min_seed_len, max_q = maxL >= k ? maxL - k + 1 : 1;
      std::vector<uint64_t> rp(maxL / 16 + 4);
      for (uint32_t wi = 0; wi < rp.size(); wi++) rp[wi] = wi < L / 16 + 3 ? tg_pack_word(bases, offs[r], L, wi) : ~0ull;
      std::vector<TgSeedHit> hits(max_q, TgSeedHit{0xDEADu, 0xDEADu, 0xDEADu});  // poison: skipped offsets must never be read
      if (L >= k) {
        const uint32_t q_last = L - k;
        tg_seed_offset(rp.data(), L, 0, k, c->slots.data(), c->slot_mask, c->dev.text4, c->dev.sa, hits[0]);  // wave 0
        if (hits[0].e != L) {
          for (uint32_t j = 0; j < (max_q + TG_PROBE_STRIDE - 1) / TG_PROBE_STRIDE; j++) {  // wave 1
            const uint32_t q = tg_probe_sample(j, q_last);
            if (q != 0xFFFFFFFFu) tg_seed_offset(rp.data(), L, q, k, c->slots.data(), c->slot_mask, c->dev.text4, c->dev.sa, hits[q]);
          }
          for (uint32_t q = 1; q <= q_last; q++) {  // wave 2
            if (tg_probe_is_sample(q, q_last) || tg_probe_bracketed(hits.data(), q, q_last)) continue;
            tg_seed_offset(rp.data(), L, q, k, c->slots.data(), c->slot_mask, c->dev.text4, c->dev.sa, hits[q]);
          }
        }
      }
      HostWarp1 w1;
      tg_seed_select_read(w1, hits.data(), L, k, out, r);
      continue;
    }
    run_lanes(lanes, [&](auto& w) {
      tg_seed_read(w, sm, bases, offs[r], L, c->opts.min_seed_len, c->slots.data(), c->slot_mask, c->dev.text4,
                   c->dev.sa, out, r);
    });
  }
  return flags ? -1 : (long long)used;
}

struct HtResult {
  std::vector<uint64_t> first;
  std::vector<uint32_t> count;
  std::vector<tg_aln> alns;
  std::vector<uint32_t> ops;
  unsigned long long n_alns = 0, n_ops = 0;
  TgCounters ctr{0, 0, 0};
  int flags = 0;
  unsigned long long items = 0, rounds = 0;  // round pipeline: evaluated (read, hit) items incl. discarded ones
};

// compact-record mode of the NEXT ht_align_batch_mode call: records are written as tg_aln_c with the first indices /
// operation offsets rebased by (first_base, ops_base) -- what one shard of tg_multi_align_batch does on the device -- and
// expanded back with tg_aln_expand before the result is handed out
static bool g_ht_compact = false;
static unsigned long long g_ht_first_base = 0, g_ht_ops_base = 0;
void ht_set_compact(int on, unsigned long long first_base, unsigned long long ops_base) {
  g_ht_compact = on != 0; g_ht_first_base = first_base; g_ht_ops_base = ops_base;
}
void* ht_align_batch_mode(void* cp, const uint8_t* bases, const uint64_t* offs, uint32_t n, int lanes, int bound_stop, int rounds);
void* ht_align_batch(void* cp, const uint8_t* bases, const uint64_t* offs, uint32_t n, int lanes, int bound_stop) {
  return ht_align_batch_mode(cp, bases, offs, n, lanes, bound_stop, 0);
}

// rounds = 1: the batch goes through the speculative round pipeline exactly as the device drives it (tg_rounds.h:
// plan / prep / task_run / post / scan per round over ALL reads, then final), and reads that leave it are redone on
// the single-warp path.
void* ht_align_batch_mode(void* cp, const uint8_t* bases, const uint64_t* offs, uint32_t n, int lanes, int bound_stop, int rounds) {
  HostCtx* c = (HostCtx*)cp;
  uint32_t maxL = 1;
  for (uint32_t r = 0; r < n; r++) maxL = std::max<uint32_t>(maxL, (uint32_t)(offs[r + 1] - offs[r]));
  if (maxL > TG_MAX_READ_LEN) return nullptr;
  auto* res = new HtResult();
  res->first.assign(n, 0);
  res->count.assign(n, 0);
  res->alns.resize((size_t)n * 8 + 1024);
  res->ops.resize((size_t)n * 64 + 65536);
  WarpBuffers wb;
  TgSeedMem sm = wb.seed_mem(maxL);
  TgWarpMem wm = wb.mem(maxL, max_bw_for(c->opts, maxL), lanes);
  wm.bound_stop = bound_stop != 0;
  wb.cands.resize(TG_MAX_ALNS_PER_READ);
  wb.arena.resize(1 << 16);
  std::vector<uint16_t> order(2 * TG_MAX_ALNS_PER_READ);
  TgWarpScratch sc{wb.cands.data(), wb.arena.data(), (uint32_t)wb.arena.size(), order.data()};
  TgAlignParams P{c->dev, c->opts};
  TgAlignOut out{res->first.data(), res->count.data(), res->alns.data(), res->ops.data(), &res->n_alns, &res->n_ops,
                 res->alns.size(), res->ops.size(), &res->flags};
  const bool compact = g_ht_compact;
  std::vector<tg_aln_c> calns(compact ? res->alns.size() : 0);
  std::vector<uint32_t> cfirst(compact ? n : 0);
  if (compact) { out.alns_c = calns.data(); out.first32 = cfirst.data(); out.first_base = g_ht_first_base; out.ops_base = g_ht_ops_base; }
  g_ht_compact = false;
  std::vector<uint8_t> done(n, 0);
  HostWarp1 w1;
  if (rounds) {
    // seeds of every read
    std::vector<tg_seed> pool((size_t)n * 8 + 1024);
    std::vector<uint64_t> sfirst(n);
    std::vector<uint32_t> scount(n);
    unsigned long long used = 0, n_smems = 0;
    for (;;) {
      used = 0; n_smems = 0; res->flags = 0;
      TgSeedOut sout{pool.data(), &used, pool.size(), sfirst.data(), scount.data(), &res->flags, &n_smems};
      for (uint32_t r = 0; r < n; r++)
        tg_seed_read(w1, sm, bases, offs[r], (uint32_t)(offs[r + 1] - offs[r]), c->opts.min_seed_len, c->slots.data(), c->slot_mask,
                     c->dev.text4, c->dev.sa, sout, r);
      if (!(res->flags & TG_FLAG_SEED_POOL)) break;
      pool.resize(pool.size() * 2);
    }
    const uint32_t rp_words = maxL / 16 + 4;
    std::vector<uint64_t> rp((size_t)n * rp_words, 0);
    std::vector<TgReadState> st(n);
    uint64_t total_hits = 0;
    for (uint32_t r = 0; r < n; r++) {
      const uint32_t L = (uint32_t)(offs[r + 1] - offs[r]);
      if (!tg_read_state_init(st[r], L, c->opts, scount[r], pool.data() + sfirst[r])) st[r].status = TG_RS_COMPLEX;
      total_hits += st[r].n_hits;
      for (uint32_t wi = 0; wi < L / 16 + 3; wi++) {
        uint64_t word = 0;
        for (uint32_t t = 0; t < 16; t++) {
          uint32_t p = wi * 16 + t;
          word |= (uint64_t)(p < L ? tg_ascii_code(bases[offs[r] + p]) : (uint32_t)TG_C_PAD) << ((15 - t) * 4);
        }
        rp[(size_t)r * rp_words + wi] = word;
      }
    }
    const size_t item_cap = (size_t)total_hits * 3 + 1024;
    std::vector<TgHit> hits(item_cap);
    std::vector<TgItemRes> ires(item_cap);
    std::vector<TgCand> cands(item_cap);
    std::vector<uint32_t> hops(item_cap * 64 + 65536);
    unsigned long long hops_used = 0, items_used = 0;
    TgHopsPool hp{hops.data(), &hops_used, hops.size()};
    std::vector<TgTask> tasks;
    std::vector<uint32_t> dp_ops;
    unsigned long long next_lo = 0;  // items appended by scan belong to the next round
    for (uint32_t round = 0; round < TG_MAX_ROUNDS; round++) {
      const unsigned long long lo = next_lo;
      for (uint32_t r = 0; r < n; r++) {  // plan
        if (st[r].status != TG_RS_ACTIVE || st[r].planned) continue;
        uint32_t b = tg_plan_batch(st[r], round);
        if (items_used + b > item_cap) b = 0;
        st[r].batch_first = (uint32_t)items_used; st[r].batch_n = b;
        for (uint32_t i = 0; i < b; i++) {
          TgItemRes& ir = ires[items_used + i];
          ir.read = r; ir.hit = st[r].next_hit + i; ir.flags = 0; ir.prev_acc = TG_NONE; ir.state = tg_pack_state(st[r].bw, st[r].x_drop);
        }
        items_used += b;
      }
      const unsigned long long hi = items_used;
      if (hi == lo) break;
      res->rounds = round + 1;
      tasks.assign((size_t)(hi - lo) * 2 * TG_PMAX + 16, TgTask{});
      unsigned long long tctr = 0, octr = 0;
      for (unsigned long long it = lo; it < hi; it++) {  // prep
        const uint32_t r = ires[it].read;
        if (!tg_item_prep(w1, P, rp.data() + (size_t)r * rp_words, st[r], ires[it].state, pool.data() + sfirst[r], scount[r], r, ires[it].hit, hits[it],
                          tasks.data(), &tctr, tasks.size(), &res->flags))
          ires[it].flags = TG_IF_FAIL;
      }
      dp_ops.assign((size_t)tctr * (2 * maxL + 64) + 64, 0);
      std::vector<uint8_t> dpt_done(tctr, 0);
      if (rounds == 2) dpt_tasks_host(c->dev, rp.data(), rp_words, tasks.data(), (size_t)tctr, wm.bound_stop, dp_ops.data(), &octr, dpt_done);
      for (unsigned long long t = 0; t < tctr; t++) {  // extend
        if (dpt_done[t]) continue;
        run_lanes(lanes, [&](auto& w) {
          tg_task_run(w, c->dev, bases, offs, tasks[t], wm.xs, wm.ys, wm.trace, wm.opsT, dp_ops.data(), &octr, dp_ops.size(),
                      &res->flags, wm.bound_stop);
        });
      }
      for (unsigned long long it = lo; it < hi; it++) {  // post
        if (ires[it].flags & TG_IF_FAIL) continue;
        tg_item_post(w1, P, st[ires[it].read], hits[it], tasks.data(), dp_ops.data(), ires[it], cands[it], hp, &res->flags);
      }
      next_lo = items_used;
      for (uint32_t r = 0; r < n; r++) {  // scan
        if (st[r].status != TG_RS_ACTIVE || st[r].batch_n == 0) continue;
        if (!tg_scan_read(w1, c->opts, st[r], ires.data(), r, &items_used, item_cap, &res->flags)) st[r].status = TG_RS_COMPLEX;
      }
    }
    res->items = items_used;
    for (uint32_t r = 0; r < n; r++) {  // final
      if (st[r].status != TG_RS_DONE) continue;
      std::vector<uint32_t> f(3 * (size_t)st[r].n_acc + 3);
      tg_round_final<HostWarp1, uint32_t>(w1, P, st[r], cands.data(), ires.data(), hops.data(), f.data(), f.data() + st[r].n_acc,
                                          f.data() + 2 * (size_t)st[r].n_acc, out, r);
      res->ctr.cells += st[r].cells; res->ctr.n_ext += st[r].n_ext; res->ctr.hits += st[r].hits;
      done[r] = 1;
    }
  }
  std::vector<tg_seed> pool1(maxL + 8);
  std::vector<uint64_t> sfirst1(1);
  std::vector<uint32_t> scount1(1);
  for (uint32_t r = 0; r < n; r++) {
    if (done[r]) continue;
    uint32_t L = (uint32_t)(offs[r + 1] - offs[r]);
    unsigned long long used = 0, n_smems = 0;
    TgSeedOut sout{pool1.data(), &used, pool1.size(), sfirst1.data(), scount1.data(), &res->flags, &n_smems};
    std::vector<TgCounters> lane_ctr(32, TgCounters{0, 0, 0});
    run_lanes(lanes, [&](auto& w) {
      tg_seed_read(w, sm, bases, offs[r], L, c->opts.min_seed_len, c->slots.data(), c->slot_mask, c->dev.text4,
                   c->dev.sa, sout, 0);
      tg_align_read(w, wm, P, bases, offs[r], L, pool1.data() + sfirst1[0], scount1[0], sc, out, r, lane_ctr[w.lane()]);
    });
    for (auto& lc : lane_ctr) { res->ctr.cells += lc.cells; res->ctr.n_ext += lc.n_ext; res->ctr.hits += lc.hits; }
  }
  if (compact) {  // back to wide records, bases removed
    for (uint32_t r = 0; r < n; r++) {
      res->first[r] = cfirst[r] - out.first_base;
      for (uint32_t i = 0; i < res->count[r]; i++) {
        tg_aln_c cc = calns[res->first[r] + i];
        cc.ops_off -= (uint32_t)out.ops_base;
        if (tg_aln_expand(c->ix, &cc, (uint32_t)(offs[r + 1] - offs[r]), &res->alns[res->first[r] + i]) != TG_OK) res->flags |= 1 << 30;
      }
    }
  }
  return res;
}
void ht_result_info(void* rp, uint64_t* out /*n_alns,n_ops,cells,n_ext,hits,flags,items,rounds*/) {
  auto* r = (HtResult*)rp;
  out[0] = r->n_alns; out[1] = r->n_ops; out[2] = r->ctr.cells; out[3] = r->ctr.n_ext; out[4] = r->ctr.hits; out[5] = (uint64_t)r->flags; out[6] = r->items; out[7] = r->rounds;
}
void ht_result_copy(void* rp, uint64_t* first, uint32_t* count, tg_aln* alns, uint32_t* ops) {
  auto* r = (HtResult*)rp;
  memcpy(first, r->first.data(), r->first.size() * 8);
  memcpy(count, r->count.data(), r->count.size() * 4);
  memcpy(alns, r->alns.data(), r->n_alns * sizeof(tg_aln));
  memcpy(ops, r->ops.data(), r->n_ops * 4);
}
void ht_result_free(void* rp) { delete (HtResult*)rp; }

// SwgExtend::extend batch with raw byte comparison (like tg_swg_extend_batch)
long long ht_swg_extend_batch(const uint8_t* xs, const uint64_t* xoff, const uint8_t* ys, const uint64_t* yoff,
                              uint32_t n, const uint32_t* bw, const int32_t* x_drop, int lanes, int bound_stop, int32_t* score,
                              uint32_t* xend, uint32_t* yend, uint64_t* ops_off, uint32_t* ops, uint64_t ops_cap,
                              uint64_t* cells_out) {
  unsigned long long total = 0, cells = 0;
  // lanes == 0: the thread-per-pair code of tg_dpt.h.  Like the device, sort the eligible pairs by (class, sub-key) and
  // let neighbours with the same (xlen, band width) share a "thread"; the rest runs paired with itself.
  struct DptOut { bool done = false; int32_t score = 0; uint32_t xend = 0, yend = 0, cells = 0; std::vector<uint32_t> ops; };
  std::vector<DptOut> dpt_res;
  if (lanes == 0) {
    dpt_res.resize(n);
    std::vector<std::pair<uint32_t, uint32_t>> order;
    for (uint32_t t = 0; t < n; t++) {
      const int xlen = (int)(xoff[t + 1] - xoff[t]), ylen = (int)(yoff[t + 1] - yoff[t]);
      if (xlen > (int)TG_MAX_READ_LEN || x_drop[t] < (int32_t)bw[t]) return -1;
      const int dcls = tg_dpt_class(xlen, (int)bw[t], x_drop[t]);
      if (dcls > 0 && xlen > 0 && ylen > 0) order.emplace_back((uint32_t)dcls * TG_DPT_CBINS + tg_dpt_subkey(xlen, (int)bw[t]), t);
    }
    std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    DptScratch dsc;
    auto run = [&](uint32_t t0, uint32_t t1, bool commit1) {
      const uint32_t tt[2] = {t0, t1};
      const int xlen = (int)(xoff[t0 + 1] - xoff[t0]), bwv = (int)bw[t0];
      TgDptMem m = dsc.mem();
      DptHalf H[2];
      DptHalfX HX[2];
      std::vector<uint64_t> ypk[2], xpk[2];
      for (int h = 0; h < 2; h++) {
        const uint32_t t = tt[h];
        const int ylen = (int)(yoff[t + 1] - yoff[t]);
        const int ncols = ylen < xlen + bwv ? ylen : xlen + bwv;
        // pack y as 4-bit codes (ACGNT only), profile from x codes
        ypk[h].assign((size_t)ylen / 16 + 4, 0);
        for (int i = 0; i < ylen; i++) ypk[h][i >> 4] |= (uint64_t)tg_ascii_code(ys[yoff[t] + i]) << ((15 - (i & 15)) * 4);
        std::vector<uint8_t> xc(xlen);
        for (int i = 0; i < xlen; i++) xc[i] = (uint8_t)tg_ascii_code(xs[xoff[t] + i]);
        tg_dpt_profile_codes(m, h, xc.data(), xlen);
        xpk[h].assign((size_t)xlen / 16 + 4, 0);
        for (int i = 0; i < 16 * ((xlen + 15) / 16); i++) xpk[h][i >> 4] |= (uint64_t)(i < xlen ? xc[i] : (uint8_t)TG_C_PAD) << ((15 - (i & 15)) * 4);
        HX[h].xseq = xpk[h].data(); HX[h].xoff = 0; HX[h].side = 0; HX[h].y0 = 0;
        H[h].ys.init(ypk[h].data(), 0, ncols, 0); H[h].ncols = ncols; H[h].x_drop = x_drop[t];
      }
      TgDpt2Result dr;
      dpt_dispatch(tg_dpt_class(xlen, bwv, x_drop[t0]), m, H[0], H[1], HX[0], HX[1], xlen, bwv, bound_stop != 0, dr);
      for (int h = 0; h < (commit1 ? 2 : 1); h++) {
        DptOut& d = dpt_res[tt[h]];
        d.done = true; d.score = dr.score[h]; d.xend = (uint32_t)dr.xend[h]; d.yend = (uint32_t)dr.yend[h]; d.cells = dr.cells[h];
        d.ops = H[h].ops;
      }
    };
    for (size_t k = 0; k < order.size();) {
      size_t e = k;
      while (e < order.size() && order[e].first == order[k].first) e++;
      for (size_t u = k; u < e; u += 2) {
        const uint32_t ia = order[u].second;
        if (u + 1 < e) {
          const uint32_t ib = order[u + 1].second;
          const bool compat = xoff[ia + 1] - xoff[ia] == xoff[ib + 1] - xoff[ib] && bw[ia] == bw[ib];
          if (compat) run(ia, ib, true);
          else { run(ia, ia, false); run(ib, ib, false); }
        } else run(ia, ia, false);
      }
      k = e;
    }
  }
  for (uint32_t t = 0; t < n; t++) {
    int xlen = (int)(xoff[t + 1] - xoff[t]), ylen = (int)(yoff[t + 1] - yoff[t]);
    if (xlen > (int)TG_MAX_READ_LEN || x_drop[t] < (int32_t)bw[t]) return -1;
    int ncols = ylen < xlen + (int)bw[t] ? ylen : xlen + (int)bw[t];
    std::vector<uint8_t> trace((size_t)(ncols + 1) * tg_trace_bytes_per_col(xlen, lanes == 0 ? 1 : lanes) + 64);
    std::vector<uint32_t> buf((size_t)xlen + ncols + 8);
    TgOps o{buf.data(), 0};
    TgSwgResult res{0, 0, 0};
    std::vector<unsigned long long> lc(32, 0), le(32, 0);
    int ylen_c = ylen > xlen + (int)bw[t] ? xlen + (int)bw[t] + 1 : ylen;
    if (t < dpt_res.size() && dpt_res[t].done) {
      const DptOut& d = dpt_res[t];
      cells += d.cells;
      score[t] = d.score; xend[t] = d.xend; yend[t] = d.yend;
      ops_off[t] = total;
      for (uint32_t i = (uint32_t)d.ops.size(); i-- > 0;) {
        if (total < ops_cap) ops[total] = d.ops[i];
        total++;
      }
      continue;
    }
    run_lanes(lanes == 0 ? 1 : lanes, [&](auto& w) {
      TgSwgResult r{0, 0, 0};
      TgOps lo{buf.data(), 0};
      tg_swg_extend(w, xs + xoff[t], ys + yoff[t], xlen, ylen_c, (int)bw[t], x_drop[t], trace.data(), r, lo,
                    lc[w.lane()], le[w.lane()], bound_stop != 0);
      if (w.lane() == 0) { res = r; o.n = lo.n; }
    });
    for (auto v : lc) cells += v;
    score[t] = res.score; xend[t] = (uint32_t)res.xend; yend[t] = (uint32_t)res.yend;
    ops_off[t] = total;
    // buffer holds rev(operations): emit forward
    for (uint32_t i = o.n; i-- > 0;) {
      if (total < ops_cap) ops[total] = o.w[i];
      total++;
    }
  }
  ops_off[n] = total;
  if (cells_out) *cells_out = cells;
  return (long long)total;
}

}  // extern "C"

// ---- micro-batcher (host_batcher.cpp) over a stand-in batch aligner -----------------------------------------------------
// The stand-in derives a read's records from its bytes alone (count = len % 4, scores / coordinates / operation words
// from a hash), scatters them through the result pools in a scrambled order and serves `n_threads` callers that each
// submit `per_thread` reads -- half of them blocking, half in windows of tickets -- and check what comes back.
// Returns 0 when every read got exactly its own records; batches / largest batch are reported through the pointers.
namespace {
struct FakeBackend {
  std::vector<uint64_t> first;
  std::vector<uint32_t> count, ops;
  std::vector<tg_aln> alns;
  int fail_every = 0, calls = 0;
};
uint32_t fb_hash(const uint8_t* p, uint32_t n, uint32_t salt) {
  uint32_t h = 2166136261u ^ salt;
  for (uint32_t i = 0; i < n; i++) h = (h ^ p[i]) * 16777619u;
  return h;
}
void fb_expect(const uint8_t* r, uint32_t len, uint32_t a, tg_aln& rec, std::vector<uint32_t>& gx, std::vector<uint32_t>& tx) {
  memset(&rec, 0, sizeof(rec));
  rec.score = (int32_t)(fb_hash(r, len, a) & 0xFFFF);
  rec.ref_id = fb_hash(r, len, 100 + a) & 7;
  rec.xlen = len;
  rec.aln_type = (uint8_t)(fb_hash(r, len, 200 + a) % 3);
  rec.primary = a == 0;
  gx.resize(1 + fb_hash(r, len, 300 + a) % 5);
  for (size_t i = 0; i < gx.size(); i++) gx[i] = fb_hash(r, len, 400 + 16 * a + (uint32_t)i);
  tx.resize(rec.aln_type == 0 ? 1 + fb_hash(r, len, 500 + a) % 3 : 0);
  for (size_t i = 0; i < tx.size(); i++) tx[i] = fb_hash(r, len, 600 + 16 * a + (uint32_t)i);
  rec.ops_len = (uint32_t)gx.size();
  rec.tx_ops_len = (uint32_t)tx.size();
}
tg_status fb_align(void* user, const uint8_t* bases, const uint64_t* offs, uint32_t n, tg_result* out) {
  FakeBackend* fb = (FakeBackend*)user;
  if (fb->fail_every && ++fb->calls % fb->fail_every == 0) return tg_fail(TG_ERR_CAPACITY, "stand-in backend failure");
  fb->first.assign(n, 0); fb->count.assign(n, 0); fb->alns.clear(); fb->ops.assign(7, 0xDEADu);
  std::vector<uint32_t> gx, tx;
  for (uint32_t i = 0; i < n; i++) {
    const uint8_t* r = bases + offs[i];
    const uint32_t len = (uint32_t)(offs[i + 1] - offs[i]);
    fb->first[i] = fb->alns.size();
    fb->count[i] = len % 4;
    for (uint32_t a = 0; a < fb->count[i]; a++) {
      tg_aln rec;
      fb_expect(r, len, a, rec, gx, tx);
      // transcript operations first, a gap, then the genome operations: offsets are not monotonic
      rec.tx_ops_off = (uint32_t)fb->ops.size();
      fb->ops.insert(fb->ops.end(), tx.begin(), tx.end());
      fb->ops.push_back(0xBEEFu);
      rec.ops_off = (uint32_t)fb->ops.size();
      fb->ops.insert(fb->ops.end(), gx.begin(), gx.end());
      fb->alns.push_back(rec);
    }
  }
  memset(out, 0, sizeof(*out));
  out->n_reads = n; out->n_alns = fb->alns.size(); out->n_ops = fb->ops.size();
  out->read_aln_first = fb->first.data(); out->read_aln_count = fb->count.data();
  out->alns = fb->alns.data(); out->ops = fb->ops.data();
  return TG_OK;
}
bool fb_check(const std::vector<uint8_t>& read, const tg_read_alns& got) {
  const uint32_t len = (uint32_t)read.size();
  if (got.n_alns != len % 4) return false;
  std::vector<uint32_t> gx, tx;
  uint32_t total = 0;
  for (uint32_t a = 0; a < got.n_alns; a++) {
    tg_aln want;
    fb_expect(read.data(), len, a, want, gx, tx);
    const tg_aln& g = got.alns[a];
    if (g.score != want.score || g.ref_id != want.ref_id || g.xlen != len || g.aln_type != want.aln_type ||
        g.primary != want.primary || g.ops_len != gx.size() || g.tx_ops_len != tx.size())
      return false;
    if ((uint64_t)g.ops_off + g.ops_len > got.n_ops || (uint64_t)g.tx_ops_off + g.tx_ops_len > got.n_ops) return false;
    if (memcmp(got.ops + g.ops_off, gx.data(), gx.size() * 4) != 0) return false;
    if (!tx.empty() && memcmp(got.ops + g.tx_ops_off, tx.data(), tx.size() * 4) != 0) return false;
    total += g.ops_len + g.tx_ops_len;
  }
  return total == got.n_ops;
}
}  // namespace

extern "C" int ht_batcher_selftest(int n_threads, int per_thread, uint32_t max_batch, uint32_t max_wait_us, int fail_every,
                                   uint64_t* n_batches, uint32_t* largest, uint64_t* n_failed_reads) {
  FakeBackend fb;
  fb.fail_every = fail_every;
  tg_batcher* b = nullptr;
  if (tg_batcher_create_backend(fb_align, &fb, max_batch, max_wait_us, &b) != TG_OK) return -1;
  std::atomic<int> bad{0};
  std::atomic<uint64_t> failed{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([&, t] {
      std::mt19937 rng(1234 + t);
      auto make = [&] {
        std::vector<uint8_t> r(rng() % 120);
        for (auto& c : r) c = "ACGTN"[rng() % 5];
        return r;
      };
      for (int i = 0; i < per_thread;) {
        if (t % 2 == 0) {  // blocking calls
          std::vector<uint8_t> r = make();
          tg_read_alns got;
          tg_status st = tg_batcher_align_read(b, r.data(), (uint32_t)r.size(), &got);
          if (st != TG_OK) { failed++; if (!fail_every || got.alns) bad++; }
          else { if (!fb_check(r, got)) bad++; tg_read_alns_free(&got); }
          i++;
        } else {  // a window of tickets, waited for in reverse order
          const int win = std::min(per_thread - i, 1 + (int)(rng() % 40));
          std::vector<std::vector<uint8_t>> rs(win);
          std::vector<uint64_t> tk(win);
          for (int k = 0; k < win; k++) {
            rs[k] = make();
            std::vector<uint8_t> copy = rs[k];
            if (tg_batcher_submit(b, copy.data(), (uint32_t)copy.size(), &tk[k]) != TG_OK) bad++;
            std::fill(copy.begin(), copy.end(), 0);  // the batcher must have taken its own copy
          }
          for (int k = win - 1; k >= 0; k--) {
            tg_read_alns got;
            tg_status st = tg_batcher_wait(b, tk[k], &got);
            if (st != TG_OK) { failed++; if (!fail_every) bad++; }
            else { if (!fb_check(rs[k], got)) bad++; tg_read_alns_free(&got); }
          }
          tg_read_alns dummy;
          if (tg_batcher_wait(b, tk[0], &dummy) != TG_ERR_INVALID) bad++;  // a ticket is good for one wait
          i += win;
        }
      }
    });
  for (auto& x : th) x.join();
  uint64_t reads = 0;
  tg_batcher_stats(b, &reads, n_batches, largest);
  if (reads != (uint64_t)n_threads * per_thread) bad++;
  // results nobody waits for are freed by destroy (checked under valgrind/ASan runs; here: must not crash)
  uint64_t orphan = 0;
  const uint8_t r3[3] = {'A', 'C', 'G'};
  tg_batcher_submit(b, r3, 3, &orphan);
  tg_batcher_destroy(b);
  if (n_failed_reads) *n_failed_reads = failed.load();
  return bad.load();
}
