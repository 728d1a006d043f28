// oracle_capi.cpp -- C entry points of the CPU ORACLE for ctypes (TEST INFRASTRUCTURE ONLY).
// Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include <chrono>
#include <cstring>
#include <string>
#include <thread>

#include "thermite_oracle.hpp"

using namespace orc;

static thread_local std::string g_err;

// Flat record: field-for-field the same layout as `tg_aln` in include/thermite_gpu.h so tests can
// compare the two byte-wise.  (Declared separately on purpose: the oracle does not include product
// headers and the product does not include the oracle.)
struct orc_aln {
  uint64_t ystart, yend, ylen;
  uint64_t tx_ystart, tx_yend, tx_ylen;
  int32_t score;
  uint32_t ref_id;
  uint32_t xstart, xend, xlen;
  uint32_t tx_or_gene_idx;
  int32_t tx_score;
  uint32_t tx_xstart, tx_xend;
  uint32_t ops_off, ops_len;
  uint32_t tx_ops_off, tx_ops_len;
  uint8_t aln_type, primary, strand, pad;
};

// RLE: word = kind | run << 3.  Match/Subst/Del/Ins runs merge; every Xclip/Yclip is its own word
// whose run field is the clip length.
static void rle_append(std::vector<uint32_t>& out, const std::vector<Op>& ops) {
  size_t first = out.size();
  for (auto& op : ops) {
    if (op.kind <= Ins && out.size() > first && (out.back() & 7u) == op.kind) out.back() += 8u;
    else out.push_back((uint32_t)op.kind | (uint32_t)(op.n << 3));
  }
}

struct orc_result {
  std::vector<orc_aln> alns;
  std::vector<uint32_t> ops;
  std::vector<uint64_t> read_off;
  double seconds = 0;
};

static AlignOpts mk_opts(uint32_t k, float pct, int32_t min_score, uint32_t range, int intron) {
  AlignOpts o;
  o.min_seed_len = k;
  o.min_aln_score_percent = pct;
  o.min_aln_score = min_score;
  o.multimap_score_range = range;
  o.intron_mode = intron != 0;
  return o;
}

static void flatten(const std::vector<GenomeAlignment>& alns, orc_result& res) {
  for (auto& g : alns) {
    orc_aln a;
    std::memset(&a, 0, sizeof(a));
    a.score = g.gx_aln.score;
    a.ref_id = (uint32_t)g.ref_id;
    a.ystart = g.gx_aln.ystart; a.yend = g.gx_aln.yend; a.ylen = g.gx_aln.ylen;
    a.xstart = (uint32_t)g.gx_aln.xstart; a.xend = (uint32_t)g.gx_aln.xend; a.xlen = (uint32_t)g.gx_aln.xlen;
    a.aln_type = g.aln_type; a.primary = g.primary; a.strand = g.strand;
    a.ops_off = (uint32_t)res.ops.size();
    rle_append(res.ops, g.gx_aln.operations);
    a.ops_len = (uint32_t)res.ops.size() - a.ops_off;
    a.tx_or_gene_idx = 0xFFFFFFFFu;
    if (g.aln_type == Exonic) {
      a.tx_or_gene_idx = (uint32_t)g.tx_idx;
      a.tx_score = g.tx_aln.score;
      a.tx_ystart = g.tx_aln.ystart; a.tx_yend = g.tx_aln.yend; a.tx_ylen = g.tx_aln.ylen;
      a.tx_xstart = (uint32_t)g.tx_aln.xstart; a.tx_xend = (uint32_t)g.tx_aln.xend;
      a.tx_ops_off = (uint32_t)res.ops.size();
      rle_append(res.ops, g.tx_aln.operations);
      a.tx_ops_len = (uint32_t)res.ops.size() - a.tx_ops_off;
    } else if (g.aln_type == Intronic) {
      a.tx_or_gene_idx = (uint32_t)g.gene_idx;
    }
    res.alns.push_back(a);
  }
}

#define ORC_TRY try {
#define ORC_CATCH(ret)                         \
  }                                            \
  catch (const std::exception& e) {            \
    g_err = e.what();                          \
    return ret;                                \
  }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
size_t orc_sizeof_aln() { return sizeof(orc_aln); }

void* orc_index_create_files(const char* fasta, const char* gtf, uint32_t sa_rate, uint32_t occ_rate) {
  ORC_TRY
  return new Index(Index::create_from_files(fasta, gtf, sa_rate, occ_rate));
  ORC_CATCH(nullptr)
}
void* orc_index_create_mem(const char* fasta_text, size_t fasta_len, const char* gtf_text, size_t gtf_len,
                           uint32_t sa_rate, uint32_t occ_rate) {
  ORC_TRY
  return new Index(Index::create(parse_fasta(std::string(fasta_text, fasta_len)), std::string(gtf_text, gtf_len),
                                 sa_rate, occ_rate));
  ORC_CATCH(nullptr)
}
void orc_index_free(void* ix) { delete (Index*)ix; }

uint64_t orc_index_text_len(void* ix) { return ((Index*)ix)->text.size(); }
void orc_index_text(void* ix, uint8_t* out) { auto& t = ((Index*)ix)->text; std::memcpy(out, t.data(), t.size()); }
void orc_index_sa(void* ix, uint32_t* out) { auto& t = ((Index*)ix)->full_sa; std::memcpy(out, t.data(), t.size() * 4); }
uint32_t orc_index_n_refs(void* ix) { return (uint32_t)((Index*)ix)->refs.size(); }
uint32_t orc_index_n_txs(void* ix) { return (uint32_t)((Index*)ix)->txome.txs.size(); }
uint32_t orc_index_n_genes(void* ix) { return (uint32_t)((Index*)ix)->txome.genes.size(); }
// out[0..3] = start_idx, end_idx, len, strand ; returns name
const char* orc_index_ref(void* ix, uint32_t i, uint64_t* out) {
  const Ref& r = ((Index*)ix)->refs[i];
  out[0] = r.start_idx; out[1] = r.end_idx; out[2] = r.len; out[3] = r.strand;
  return r.name.c_str();
}
const char* orc_index_tx(void* ix, uint32_t i, uint64_t* out /*gene_idx, strand, n_exons, seq_len*/) {
  const Tx& t = ((Index*)ix)->txome.txs[i];
  out[0] = t.gene_idx; out[1] = t.strand; out[2] = t.exons.size(); out[3] = t.seq.size();
  return t.id.c_str();
}
void orc_index_tx_seq(void* ix, uint32_t i, uint8_t* out) {
  const Tx& t = ((Index*)ix)->txome.txs[i];
  std::memcpy(out, t.seq.data(), t.seq.size());
}
void orc_index_tx_exons(void* ix, uint32_t i, uint64_t* out /*pairs*/) {
  const Tx& t = ((Index*)ix)->txome.txs[i];
  for (size_t e = 0; e < t.exons.size(); e++) { out[2 * e] = t.exons[e].start; out[2 * e + 1] = t.exons[e].end; }
}
const char* orc_index_gene_id(void* ix, uint32_t i) { return ((Index*)ix)->txome.genes[i].id.c_str(); }
const char* orc_index_gene_name(void* ix, uint32_t i) { return ((Index*)ix)->txome.genes[i].name.c_str(); }

// tree: 0 = exon_to_tx, 1 = gene_intervals.  Returns the number of values written.
uint64_t orc_interval_find(void* ix, int tree, uint64_t s, uint64_t e, uint64_t* out, uint64_t cap) {
  auto& t = tree == 0 ? ((Index*)ix)->txome.exon_to_tx : ((Index*)ix)->txome.gene_intervals;
  auto v = t.find(s, e);
  for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return v.size();
}

void orc_counters(void* ix, uint64_t* out) {
  const Counters& c = ((Index*)ix)->counters;
  out[0] = c.swg_cells; out[1] = c.swg_calls; out[2] = c.occ_lookups; out[3] = c.fmd_ext; out[4] = c.sa_locates; out[5] = c.hits;
}
void orc_counters_reset(void* ix) { ((Index*)ix)->counters = Counters(); }

int orc_suffix_array(const uint8_t* text, uint64_t n, uint32_t* out) {
  ORC_TRY
  auto sa = suffix_array(std::vector<uint8_t>(text, text + n));
  std::memcpy(out, sa.data(), n * 4);
  return 0;
  ORC_CATCH(-1)
}

// mems out as (ref_idx, query_idx, len) u64 triples; returns count (may exceed cap)
int64_t orc_all_smems(void* ix, const uint8_t* read, uint64_t len, uint32_t k, int brute, uint64_t* out, uint64_t cap) {
  ORC_TRY
  std::vector<uint8_t> r(read, read + len);
  for (auto& c : r) if (c >= 'a' && c <= 'z') c -= 32;
  auto m = brute ? ((Index*)ix)->all_smems_brute(r.data(), len, k) : ((Index*)ix)->all_smems(r.data(), len, k);
  ((Index*)ix)->fold_tl();
  for (size_t i = 0; i < m.size() && i < cap; i++) { out[3 * i] = m[i].ref_idx; out[3 * i + 1] = m[i].query_idx; out[3 * i + 2] = m[i].len; }
  return (int64_t)m.size();
  ORC_CATCH(-1)
}

// One extension.  status: 0 ok, 1 = the reference panics on this input (quirk Q4 / assert).
int orc_swg_extend(const uint8_t* x, uint64_t xlen, const uint8_t* y, uint64_t ylen, uint64_t max_bw, uint64_t bw,
                   int32_t x_drop, int32_t* score, uint32_t* xend, uint32_t* yend, uint32_t* ops, uint32_t ops_cap,
                   uint32_t* n_ops, uint64_t* cells) {
  try {
    SwgExtend swg(max_bw, -1, -1, 1, -1);
    Alignment a = swg.extend(x, xlen, y, ylen, bw, x_drop);
    std::vector<uint32_t> r;
    rle_append(r, a.operations);
    *score = a.score; *xend = (uint32_t)a.xend; *yend = (uint32_t)a.yend; *n_ops = (uint32_t)r.size();
    if (cells) *cells = swg.cells;
    for (size_t i = 0; i < r.size() && i < ops_cap; i++) ops[i] = r[i];
    return 0;
  } catch (const ReferencePanic& e) {
    g_err = e.what();
    return 1;
  }
}

// Batch of independent extensions (fresh SwgExtend per task, max_bw = bw), outputs in the layout of
// tg_swg_extend_batch.  Returns total RLE words (ops written up to ops_cap), -1 on a reference panic.
int64_t orc_swg_extend_batch(const uint8_t* xs, const uint64_t* xoff, const uint8_t* ys, const uint64_t* yoff,
                             uint64_t n, const uint32_t* bw, const int32_t* x_drop, int32_t* score,
                             uint32_t* xend, uint32_t* yend, uint64_t* ops_off, uint32_t* ops, uint64_t ops_cap,
                             uint64_t* cells_total) {
  uint64_t total = 0, cells = 0;
  try {
    for (uint64_t t = 0; t < n; t++) {
      SwgExtend swg(bw[t], -1, -1, 1, -1);
      Alignment a = swg.extend(xs + xoff[t], xoff[t + 1] - xoff[t], ys + yoff[t], yoff[t + 1] - yoff[t], bw[t], x_drop[t]);
      cells += swg.cells;
      std::vector<uint32_t> r;
      rle_append(r, a.operations);
      score[t] = a.score; xend[t] = (uint32_t)a.xend; yend[t] = (uint32_t)a.yend;
      ops_off[t] = total;
      for (size_t i = 0; i < r.size(); i++) if (total + i < ops_cap) ops[total + i] = r[i];
      total += r.size();
    }
    ops_off[n] = total;
    if (cells_total) *cells_total = cells;
    return (int64_t)total;
  } catch (const ReferencePanic& e) {
    g_err = e.what();
    return -1;
  }
}

// The same on n_threads host threads (contiguous task ranges; results in task order).  Test / bench infrastructure for
// the config-5 microbench: 2^20 pairs per band width are too many for one core.
int64_t orc_swg_extend_batch_mt(const uint8_t* xs, const uint64_t* xoff, const uint8_t* ys, const uint64_t* yoff,
                                uint64_t n, const uint32_t* bw, const int32_t* x_drop, int32_t* score,
                                uint32_t* xend, uint32_t* yend, uint64_t* ops_off, uint32_t* ops, uint64_t ops_cap,
                                uint64_t* cells_total, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::vector<uint32_t>> words(n_threads);
  std::vector<uint64_t> cells(n_threads, 0);
  std::vector<std::string> errs(n_threads);
  std::vector<uint32_t> lens(n);
  auto work = [&](int t) {
    const uint64_t t0 = n * t / n_threads, t1 = n * (t + 1) / n_threads;
    try {
      for (uint64_t k = t0; k < t1; k++) {
        SwgExtend swg(bw[k], -1, -1, 1, -1);
        Alignment a = swg.extend(xs + xoff[k], xoff[k + 1] - xoff[k], ys + yoff[k], yoff[k + 1] - yoff[k], bw[k], x_drop[k]);
        cells[t] += swg.cells;
        const size_t before = words[t].size();
        rle_append(words[t], a.operations);
        lens[k] = (uint32_t)(words[t].size() - before);
        score[k] = a.score; xend[k] = (uint32_t)a.xend; yend[k] = (uint32_t)a.yend;
      }
    } catch (const ReferencePanic& e) {
      errs[t] = e.what();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (auto& e : errs)
    if (!e.empty()) { g_err = e; return -1; }
  uint64_t total = 0, c = 0;
  for (uint64_t k = 0; k < n; k++) { ops_off[k] = total; total += lens[k]; }
  ops_off[n] = total;
  uint64_t pos = 0;
  for (int t = 0; t < n_threads; t++) {
    c += cells[t];
    for (uint32_t w : words[t]) { if (pos < ops_cap) ops[pos] = w; pos++; }
  }
  if (cells_total) *cells_total = c;
  return (int64_t)total;
}

// src/aligner.rs:352-407 test hook (max_bw = bw)
int orc_extend_left_right(const uint8_t* ref_seq, uint64_t ref_len, uint64_t h_ref, uint64_t h_q, uint64_t h_len,
                          const uint8_t* read, uint64_t read_len, uint64_t max_bw, uint64_t bw, int32_t x_drop,
                          int64_t* out /*score,ystart,xstart,yend,xend,ylen,xlen*/, uint32_t* ops, uint32_t ops_cap, uint32_t* n_ops) {
  ORC_TRY
  SwgExtend swg(max_bw, -1, -1, 1, -1);
  Alignment a = extend_left_right(ref_seq, ref_len, Mem{h_ref, h_q, h_len}, read, read_len, swg, bw, x_drop);
  out[0] = a.score; out[1] = a.ystart; out[2] = a.xstart; out[3] = a.yend; out[4] = a.xend; out[5] = a.ylen; out[6] = a.xlen;
  std::vector<uint32_t> r;
  rle_append(r, a.operations);
  *n_ops = (uint32_t)r.size();
  for (size_t i = 0; i < r.size() && i < ops_cap; i++) ops[i] = r[i];
  return 0;
  ORC_CATCH(-1)
}

static Tx mk_tx(const uint64_t* exons, uint32_t n_exons) {
  Tx tx;
  tx.strand = true;
  tx.gene_idx = 0;
  for (uint32_t i = 0; i < n_exons; i++) tx.exons.push_back(Exon{exons[2 * i], exons[2 * i + 1], 0});
  return tx;
}
// src/txome.rs:82-103 test hook
int orc_lift_mem_to_tx(const uint64_t* exons, uint32_t n_exons, const uint64_t* mem_in, uint64_t* mem_out) {
  ORC_TRY
  Mem m = lift_mem_to_tx(Mem{mem_in[0], mem_in[1], mem_in[2]}, mk_tx(exons, n_exons));
  mem_out[0] = m.ref_idx; mem_out[1] = m.query_idx; mem_out[2] = m.len;
  return 0;
  ORC_CATCH(-1)
}
// src/txome.rs:110-160 test hook; ops are un-RLE'd (kind | n<<3 per op) in and out
int orc_lift_tx_to_gx(const uint64_t* exons, uint32_t n_exons, uint64_t ystart, uint64_t yend, const uint32_t* ops,
                      uint32_t n_ops, uint64_t* out_ystart_yend, uint32_t* out_ops, uint32_t cap, uint32_t* out_n) {
  ORC_TRY
  Alignment a;
  a.ystart = ystart; a.yend = yend;
  for (uint32_t i = 0; i < n_ops; i++) a.operations.push_back(Op{(uint8_t)(ops[i] & 7), ops[i] >> 3});
  Alignment g = lift_tx_to_gx(a, mk_tx(exons, n_exons));
  out_ystart_yend[0] = g.ystart; out_ystart_yend[1] = g.yend;
  *out_n = (uint32_t)g.operations.size();
  for (size_t i = 0; i < g.operations.size() && i < cap; i++) out_ops[i] = g.operations[i].kind | (uint32_t)(g.operations[i].n << 3);
  return 0;
  ORC_CATCH(-1)
}
// src/aligner.rs:317-349 test hook: rows of (name_id, strand, ystart, yend, score); returns kept row ids in order
int64_t orc_filter_overlapping(const int64_t* rows, uint32_t n, uint32_t* kept) {
  ORC_TRY
  std::vector<GenomeAlignment> v;
  for (uint32_t i = 0; i < n; i++) {
    GenomeAlignment g;
    g.ref_name = std::string(1, (char)('a' + rows[5 * i]));
    g.strand = rows[5 * i + 1] != 0;
    g.gx_aln.ystart = rows[5 * i + 2]; g.gx_aln.yend = rows[5 * i + 3]; g.gx_aln.score = (int32_t)rows[5 * i + 4];
    g.ref_id = i;  // carries the row id
    v.push_back(g);
  }
  auto r = filter_overlapping(v);
  for (size_t i = 0; i < r.size(); i++) kept[i] = (uint32_t)r[i].ref_id;
  return (int64_t)r.size();
  ORC_CATCH(-1)
}

uint32_t orc_multimapq(uint64_t n) { return multimapq(n); }

// Align a batch (reads concatenated, offs[n+1]).  n_threads > 1 shards reads contiguously over host
// threads (the reference itself is single-threaded).  Returns a result handle.
void* orc_align_batch(void* ixp, const uint8_t* bases, const uint64_t* offs, uint64_t n, uint32_t k, float pct,
                      int32_t min_score, uint32_t range, int intron, uint32_t n_threads) {
  ORC_TRY
  Index* ix = (Index*)ixp;
  AlignOpts o = mk_opts(k, pct, min_score, range, intron);
  auto* res = new orc_result();
  res->read_off.assign(n + 1, 0);
  if (n_threads < 1) n_threads = 1;
  std::vector<std::vector<std::vector<GenomeAlignment>>> per(n_threads);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](uint32_t t) {
    uint64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    per[t].reserve(hi - lo);
    for (uint64_t r = lo; r < hi; r++) per[t].push_back(align_read(*ix, bases + offs[r], offs[r + 1] - offs[r], o));
    ix->fold_tl();
  };
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t r = 0;
  for (auto& shard : per)
    for (auto& alns : shard) {
      res->read_off[r] = res->alns.size();
      flatten(alns, *res);
      r++;
    }
  res->read_off[n] = res->alns.size();
  return res;
  ORC_CATCH(nullptr)
}
uint64_t orc_result_n_alns(void* r) { return ((orc_result*)r)->alns.size(); }
uint64_t orc_result_n_ops(void* r) { return ((orc_result*)r)->ops.size(); }
double orc_result_seconds(void* r) { return ((orc_result*)r)->seconds; }
void orc_result_copy(void* rp, void* alns, uint32_t* ops, uint64_t* read_off) {
  auto* r = (orc_result*)rp;
  std::memcpy(alns, r->alns.data(), r->alns.size() * sizeof(orc_aln));
  std::memcpy(ops, r->ops.data(), r->ops.size() * 4);
  std::memcpy(read_off, r->read_off.data(), r->read_off.size() * 8);
}
void orc_result_free(void* r) { delete (orc_result*)r; }

// PAF (sam=0) or SAM (sam=1) text of a FASTQ given as text.  Caller frees with orc_free.
char* orc_align_fastq_text(void* ixp, const char* fastq, uint64_t len, uint32_t k, float pct, int32_t min_score,
                           uint32_t range, int intron, int sam, uint64_t* out_len) {
  ORC_TRY
  auto reads = parse_fastq(std::string(fastq, len));
  std::string s = align_fastq(*(Index*)ixp, reads, mk_opts(k, pct, min_score, range, intron), sam != 0);
  ((Index*)ixp)->fold_tl();
  char* out = (char*)malloc(s.size() + 1);
  std::memcpy(out, s.data(), s.size());
  out[s.size()] = 0;
  *out_len = s.size();
  return out;
  ORC_CATCH(nullptr)
}
void orc_free(void* p) { free(p); }

}  // extern "C"
