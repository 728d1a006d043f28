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

// ---- tg_dpt.h on the host: one "thread", strides of 1 --------------------------------------------------------------
struct DptScratch {
  std::vector<uint32_t> msk, tr;
  DptScratch() : msk(32, 0), tr((size_t)(TG_DPT_MAX_X + 2 * TG_MAX_READ_LEN + 2) * 5, 0) {}
  TgDptMem mem() { return TgDptMem{msk.data(), 1, tr.data(), 1, 1, 128}; }
};
template <int WB>
void dpt_one(const TgDptMem& m, TgDptY& ys, int xlen, int ncols, int bw, int x_drop, bool bound_stop, TgDptResult& res,
             std::vector<uint32_t>& ops) {
  tg_dpt_fill<WB>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res);
  uint32_t n = tg_dpt_traceback<WB>(m, ys, xlen, bw, res, [](uint32_t, uint32_t, uint32_t) {});
  ops.assign(n, 0);
  tg_dpt_traceback<WB>(m, ys, xlen, bw, res, [&](uint32_t i, uint32_t kind, uint32_t run) { ops[i] = kind | (run << 3); });
}
void dpt_dispatch(int cls, const TgDptMem& m, TgDptY& ys, int xlen, int ncols, int bw, int x_drop, bool bound_stop,
                  TgDptResult& res, std::vector<uint32_t>& ops) {
  switch (cls) {
    case 1: dpt_one<4>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 2: dpt_one<8>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 3: dpt_one<16>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 4: dpt_one<24>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 5: dpt_one<32>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 6: dpt_one<40>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 7: dpt_one<48>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 8: dpt_one<56>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 9: dpt_one<64>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    case 10: dpt_one<72>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
    default: dpt_one<80>(m, ys, xlen, ncols, bw, x_drop, bound_stop, res, ops); break;
  }
}
// one round task the way k_round_dpt runs it; false when the task is not eligible (class 0)
bool dpt_task_host(const TgIndexDev& ix, const uint64_t* rp, TgTask& t, DptScratch& sc, bool bound_stop, uint32_t* ops_pool,
                   unsigned long long* ops_ctr) {
  const int xlen = (int)t.xlen, bw = (int)t.bw, ylen = (int)t.ylen;
  const int cls = tg_dpt_class(xlen, bw, t.x_drop);
  if (cls == 0) return false;
  const int ncols = ylen < xlen + bw ? ylen : xlen + bw;
  TgDptMem m = sc.mem();
  TgDptY ys{tg_seq_of(ix, t.seqsel), t.y0, ncols, (int)t.side, 0, 0, 0, 0, -1};
  tg_dpt_profile(m, rp, t.xoff, xlen, t.side);
  TgDptResult res{0, 0, 0, 0};
  std::vector<uint32_t> ops;
  dpt_dispatch(cls, m, ys, xlen, ncols, bw, t.x_drop, bound_stop, res, ops);
  t.score = res.score; t.xend = (uint32_t)res.xend; t.yend = (uint32_t)res.yend; t.cells = res.cells;
  t.ops_off = (uint32_t)*ops_ctr; t.ops_n = (uint32_t)ops.size();
  for (uint32_t w : ops) ops_pool[(*ops_ctr)++] = w;
  return true;
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
      DptScratch dsc;
      for (unsigned long long t = 0; t < tctr; t++) {  // extend
        if (rounds == 2 && dpt_task_host(c->dev, rp.data() + (size_t)tasks[t].read * rp_words, tasks[t], dsc, wm.bound_stop, dp_ops.data(), &octr))
          continue;
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
    const int dcls = tg_dpt_class(xlen, (int)bw[t], x_drop[t]);
    if (lanes == 0 && dcls > 0 && xlen > 0 && ylen > 0) {
      // pack y as 4-bit codes (ACGNT only), profile from x codes
      std::vector<uint64_t> ypk((size_t)ylen / 16 + 4, 0);
      for (int i = 0; i < ylen; i++) ypk[i >> 4] |= (uint64_t)tg_ascii_code(ys[yoff[t] + i]) << ((15 - (i & 15)) * 4);
      std::vector<uint8_t> xc(xlen);
      for (int i = 0; i < xlen; i++) xc[i] = (uint8_t)tg_ascii_code(xs[xoff[t] + i]);
      DptScratch dsc;
      TgDptMem m = dsc.mem();
      tg_dpt_profile_codes(m, xc.data(), xlen);
      TgDptY yy{ypk.data(), 0, ncols, 0, 0, 0, 0, 0, -1};
      TgDptResult dr{0, 0, 0, 0};
      std::vector<uint32_t> dops;
      dpt_dispatch(dcls, m, yy, xlen, ncols, (int)bw[t], x_drop[t], bound_stop != 0, dr, dops);
      cells += dr.cells;
      score[t] = dr.score; xend[t] = (uint32_t)dr.xend; yend[t] = (uint32_t)dr.yend;
      ops_off[t] = total;
      for (uint32_t i = (uint32_t)dops.size(); i-- > 0;) {
        if (total < ops_cap) ops[total] = dops[i];
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

// ---- the device's SAM formatter (tg_textfmt.h) on the host: count -> exclusive scan -> write, like tg_paf.cu --------------
#include "tg_textfmt.h"
extern "C" int ht_format_sam(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                             const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                             char** out, size_t* out_len) {
  TgTextTables tb;
  tb.build(ix);
  const uint32_t n = res->n_reads;
  std::vector<unsigned long long> line_off(n + 1, 0);
  TgTextParams p;
  memset(&p, 0, sizeof(p));
  p.n_reads = n; p.bases = bases; p.offs = offs;
  p.aln_first = res->read_aln_first; p.aln_count = res->read_aln_count; p.alns = res->alns; p.ops = res->ops;
  p.names = names; p.name_offs = name_offs; p.quals = quals; p.qual_offs = qual_offs;
  p.ref_names = tb.ref_names.data(); p.ref_name_offs = tb.ref_name_offs.data();
  p.tx_ids = tb.tx_ids.data(); p.tx_id_offs = tb.tx_id_offs.data();
  p.gene_ids = tb.gene_ids.data(); p.gene_id_offs = tb.gene_id_offs.data();
  p.gene_names = tb.gene_names.data(); p.gene_name_offs = tb.gene_name_offs.data();
  p.tx_gene = tb.tx_gene.data();
  p.line_off = line_off.data();
  p.mapq[0] = 255; p.mapq[1] = 255; p.mapq[5] = 0;
  for (int k = 2; k <= 4; k++) p.mapq[k] = (uint32_t)std::lround(-10.0f * std::log10(1.0f - 1.0f / (float)k));
  unsigned long long total = 0;
  for (uint32_t r = 0; r < n; r++) { const unsigned long long b = tg_sam_read<false>(p, r); line_off[r] = total; total += b; }
  line_off[n] = total;
  char* text = (char*)malloc(total + 1);
  if (!text) return -1;
  memset(text, '#', total);
  text[total] = 0;
  p.text = text;
  for (uint32_t r = 0; r < n; r++) tg_sam_read<true>(p, r);
  *out = text; *out_len = (size_t)total;
  return 0;
}

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
