// tg_rounds.h -- the hit loop of align_read (reference src/aligner.rs:143-175) as a ROUND pipeline:
//
//   round r, for every read that still has hits:   prep (thread per read)  ->  extend (warp per task)  ->  post (thread per read)
//
// instead of one warp walking a read from start to finish (tg_align_read in tg_core.h, still used for the few
// "complex" reads: many hits, many transcripts per seed, many accepted alignments).  Why: the per-hit control
// code (window, interval stab, lifting, seed matching, choosing tx vs genome, filters) is branchy and serial; run by
// lane 0 of a warp it wastes 31/32 of every issue slot and thrashes the instruction cache, while the DP wants all
// lanes.  Splitting puts the control code into small thread-per-read kernels (32 reads per warp) and leaves a DP
// kernel whose warps all execute the same hot loop.  The serial dependency between hits of ONE read (band
// narrowing, running maximum) is preserved: a read advances by exactly one hit per round.
//
// Everything here is expressed with the same building blocks as tg_core.h (and therefore under the same parity
// tests): prep/post are tg_align_seed_hit cut at the SwgExtend calls.
#pragma once
#include "tg_core.h"

#define TG_CMAX 12            // transcript candidates tabulated per hit
#define TG_PMAX 6             // distinct extension problems per hit (problem 0 = genome)
#define TG_ACC_MAX 4          // accepted alignments kept per read on the round path
#define TG_ARENA_WORDS 320u   // RLE words of those alignments
#define TG_FAST_MAX_HITS 48u  // reads with more hits take the single-warp path

enum { TG_RS_DONE = 0, TG_RS_ACTIVE = 1, TG_RS_COMPLEX = 2 };

struct TgTask {  // one SwgExtend::extend call (src/swg.rs:31)
  uint32_t read;
  uint32_t xoff, xlen;  // side 0: x = read[xoff, xoff+xlen) ; side 1: x = reversed read[0, xlen)
  uint32_t ylen;        // clamped to xlen + bw + 1
  uint64_t y0;          // side 0: first y symbol ; side 1: one past it (y runs downwards)
  uint32_t bw;
  int32_t x_drop;
  uint8_t side, seqsel, pad0, pad1;  // seqsel 0: text4, 1: txseq4
  // results
  int32_t score;
  uint32_t xend, yend, cells, ops_off, ops_n;
};

struct TgProbE {
  uint64_t lo_abs, hi_abs, r_abs;
  uint32_t q, len;
  int32_t task_r, task_l;  // -1: trivial (empty x or empty y)
  uint8_t seqsel, pad[7];
};
struct TgCandE {
  uint64_t t0;
  uint32_t tx_idx, prob, tr, tlen;
};
struct TgHit {
  uint32_t ref_idx, q, len, ref_id;
  uint32_t n_cand, n_prob, bw;
  int32_t x_drop;
  TgProbE prob[TG_PMAX];
  TgCandE cand[TG_CMAX];
};
struct TgReadState {
  uint32_t L;
  int32_t min_aln, max_aln;
  uint32_t bw, x_drop;
  uint32_t si, rk;       // cursor: seed index, occurrence (counts down)
  uint32_t n_acc, arena_used;
  uint32_t status;
  uint32_t hits, n_ext;        // work counters of this read, added to the batch totals when it finishes on this path
  unsigned long long cells;
};

TG_HD const uint64_t* tg_seq_of(const TgIndexDev& ix, uint32_t seqsel) { return seqsel ? ix.txseq4 : ix.text4; }

// src/aligner.rs:130-138
TG_HD void tg_read_state_init(TgReadState& s, uint32_t L, const tg_opts& o, uint32_t n_seeds, const tg_seed* seeds) {
  float prod = o.min_aln_score_percent * (float)L;
  int32_t pct_score = (int32_t)prod;
  s.L = L;
  s.min_aln = pct_score > o.min_aln_score ? pct_score : o.min_aln_score;
  s.max_aln = s.min_aln;
  s.bw = (s.min_aln < 0) ? 0u : (L > (uint32_t)s.min_aln ? L - (uint32_t)s.min_aln : 0u);
  s.x_drop = s.bw;
  s.si = 0;
  s.rk = n_seeds ? seeds[0].count : 0;
  s.n_acc = 0; s.arena_used = 0;
  s.hits = 0; s.n_ext = 0; s.cells = 0;
  s.status = n_seeds ? TG_RS_ACTIVE : TG_RS_DONE;
}

// ---- prep: everything of align_seed_hit (src/aligner.rs:198-258) that happens before a SwgExtend call ------------
// Tabulates the genome problem and one candidate per transcript that the seed's exon stab yields (ALL of them: the
// reference's early `break` at a perfect transcript is applied in post), maps identical problems onto each other and
// emits one task per non-trivial extension.  Returns false when a table overflows (the read becomes "complex").
template <class W>
TG_HDN bool tg_round_prep(W& w, const TgAlignParams& P, const uint64_t* rp, TgReadState& st, const tg_seed* seeds,
                          uint32_t read, TgHit& hit, TgTask* tasks, unsigned long long* task_ctr, unsigned long long task_cap) {
  const TgIndexDev& ix = P.ix;
  const uint32_t L = st.L;
  const tg_seed sd = seeds[st.si];
  const uint32_t ref_idx = sd.direct ? sd.sa_lo : TG_LDG(ix.sa + sd.sa_lo + (st.rk - 1));
  const uint32_t q = sd.query_idx, len = sd.len, bw = st.bw;
  hit.ref_idx = ref_idx; hit.q = q; hit.len = len; hit.bw = bw; hit.x_drop = (int32_t)st.x_drop;
  hit.ref_id = tg_idx_to_ref(ix.refs, ix.n_refs, ref_idx);
  const TgRef aref = ix.refs[hit.ref_id];
  const uint64_t span = (uint64_t)L + bw;
  uint64_t seq_start = ref_idx > span ? ref_idx - span : 0;
  if (seq_start < aref.start_idx) seq_start = aref.start_idx;
  uint64_t seq_end = (uint64_t)ref_idx + len + L + bw;
  if (seq_end > (uint64_t)aref.end_idx - 1) seq_end = (uint64_t)aref.end_idx - 1;
  hit.n_prob = 1; hit.n_cand = 0;
  hit.prob[0].lo_abs = seq_start; hit.prob[0].hi_abs = seq_end; hit.prob[0].r_abs = ref_idx;
  hit.prob[0].q = q; hit.prob[0].len = len; hit.prob[0].seqsel = 0;
  const TgStabRange xr = tg_stab_begin<W>(w, ix.exon_stab, ix.n_exon_stab, ix.exon_maxlen, ref_idx, ref_idx + len);
  uint32_t next_rank = 0;
  for (;;) {
    uint32_t tx_idx = 0, xrank = 0;
    if (!tg_stab_next<W>(w, ix.exon_stab, xr, next_rank, xrank, tx_idx)) break;
    next_rank = xrank + 1;
    const uint32_t e0 = TG_LDG(ix.tx_exon_off + tx_idx), e1 = TG_LDG(ix.tx_exon_off + tx_idx + 1);
    const uint64_t t0 = TG_LDG(ix.tx_seq_off + tx_idx), t1 = TG_LDG(ix.tx_seq_off + tx_idx + 1);
    uint32_t tr = 0, tq = 0, tl = 0;
    if (!tg_lift_mem_to_tx(ix.te_start, ix.te_end, e0, e1, ref_idx, q, len, tr, tq, tl)) continue;
    tl += tg_match_fwd(rp, tq + tl, L, ix.txseq4, t0 + tr + tl, t1);
    {
      uint32_t back = tg_match_bwd(rp, tq, ix.txseq4, t0 + tr, t0);
      tr -= back; tq -= back; tl += back;
    }
    if (hit.n_cand >= TG_CMAX) return false;
    TgProblem pt{ix.txseq4, t0, t1, t0 + tr, tq, tl};
    uint32_t pi = hit.n_prob;
    for (uint32_t k = 0; k < hit.n_prob; k++) {
      const TgProbE& e = hit.prob[k];
      TgProblem pk{tg_seq_of(ix, e.seqsel), e.lo_abs, e.hi_abs, e.r_abs, e.q, e.len};
      if (tg_same_problem<W>(w, pt, pk, L, bw)) { pi = k; break; }
    }
    if (pi == hit.n_prob) {
      if (hit.n_prob >= TG_PMAX) return false;
      TgProbE& e = hit.prob[hit.n_prob++];
      e.lo_abs = t0; e.hi_abs = t1; e.r_abs = t0 + tr; e.q = tq; e.len = tl; e.seqsel = 1;
    }
    TgCandE& c = hit.cand[hit.n_cand++];
    c.t0 = t0; c.tx_idx = tx_idx; c.prob = pi; c.tr = tr; c.tlen = (uint32_t)(t1 - t0);
  }
  // tasks (the windows are those of extend_left_right, src/aligner.rs:360-375)
  uint32_t need = 0;
  for (uint32_t k = 0; k < hit.n_prob; k++) {
    TgProbE& e = hit.prob[k];
    TgProblem pk{nullptr, e.lo_abs, e.hi_abs, e.r_abs, e.q, e.len};
    uint32_t ncR, ncL;
    tg_problem_windows(pk, L, bw, ncR, ncL);
    e.task_r = ncR ? 0 : -1;
    e.task_l = ncL ? 0 : -1;
    need += (ncR ? 1u : 0u) + (ncL ? 1u : 0u);
  }
  unsigned long long base = need ? w.atomic_add(task_ctr, (unsigned long long)need) : 0;
  if (base + need > task_cap) return false;
  for (uint32_t k = 0; k < hit.n_prob; k++) {
    TgProbE& e = hit.prob[k];
    const uint32_t xr_len = L - (e.q + e.len);
    if (e.task_r == 0) {
      TgTask& t = tasks[base];
      e.task_r = (int32_t)base++;
      t.read = read; t.side = 0; t.seqsel = e.seqsel; t.xoff = e.q + e.len; t.xlen = xr_len;
      t.y0 = e.r_abs + e.len;
      uint64_t yl = e.hi_abs - t.y0;
      t.ylen = (uint32_t)(yl > (uint64_t)xr_len + bw ? (uint64_t)xr_len + bw + 1 : yl);
      t.bw = bw; t.x_drop = (int32_t)st.x_drop;
    }
    if (e.task_l == 0) {
      TgTask& t = tasks[base];
      e.task_l = (int32_t)base++;
      t.read = read; t.side = 1; t.seqsel = e.seqsel; t.xoff = 0; t.xlen = e.q;
      t.y0 = e.r_abs;
      uint64_t ys0 = (e.r_abs - e.lo_abs > span) ? e.r_abs - span : e.lo_abs;
      uint64_t yl = e.r_abs - ys0;
      t.ylen = (uint32_t)(yl > (uint64_t)e.q + bw ? (uint64_t)e.q + bw + 1 : yl);
      t.bw = bw; t.x_drop = (int32_t)st.x_drop;
    }
  }
  return true;
}

// ---- extend: one task, one warp (same device functions as the single-warp path) --------------------------------------
template <class W, int RMAX = 16>
TG_HDN void tg_task_run(W& w, const TgIndexDev& ix, const uint8_t* bases, const uint64_t* offs, TgTask& t, uint8_t* sx,
                        uint8_t* sy, uint8_t* trace, uint32_t* obuf, uint32_t* ops_pool, unsigned long long* ops_ctr,
                        unsigned long long ops_cap, int* flags, bool bound_stop) {
  const int lane = w.lane();
  const uint64_t* seq = tg_seq_of(ix, t.seqsel);
  const uint64_t roff = offs[t.read];
  const int xlen = (int)t.xlen, ylen = (int)t.ylen, bw = (int)t.bw;
  const int ncols = ylen < xlen + bw ? ylen : xlen + bw;
  if (t.side == 0) {
    for (int i = lane; i < xlen; i += W::LANES) sx[i] = (uint8_t)tg_ascii_code(TG_LDG(bases + roff + t.xoff + i));
    for (int i = lane; i < ncols; i += W::LANES) sy[i] = (uint8_t)tg_code_at(seq, t.y0 + (uint64_t)i);
  } else {
    for (int i = lane; i < xlen; i += W::LANES) sx[i] = (uint8_t)tg_ascii_code(TG_LDG(bases + roff + (uint64_t)(xlen - 1 - i)));
    for (int i = lane; i < ncols; i += W::LANES) sy[i] = (uint8_t)tg_code_at(seq, t.y0 - 1 - (uint64_t)i);
  }
  w.sync();
  TgSwgResult res{0, 0, 0};
  TgOps o{obuf, 0};
  unsigned long long cells = 0, n_ext = 0;
  tg_swg_extend<W, RMAX>(w, sx, sy, xlen, ylen, bw, t.x_drop, trace, res, o, cells, n_ext, bound_stop);
  cells = w.sum64(cells);
  unsigned long long dst = 0;
  if (lane == 0 && o.n) dst = w.atomic_add(ops_ctr, (unsigned long long)o.n);
  dst = w.shfl64(dst, 0);
  if (dst + o.n > ops_cap) {
    if (lane == 0) w.atomic_or(flags, TG_FLAG_OPS_POOL);
    o.n = 0;
  }
  for (uint32_t i = lane; i < o.n; i += W::LANES) ops_pool[dst + i] = obuf[i];  // generation order = rev(operations)
  if (lane == 0) {
    t.score = res.score; t.xend = (uint32_t)res.xend; t.yend = (uint32_t)res.yend; t.cells = (uint32_t)cells;
    t.ops_off = (uint32_t)dst; t.ops_n = o.n;
  }
  w.sync();
}

// ---- post: the rest of align_seed_hit (src/aligner.rs:240-314) + the filters of align_read (:146-174) ----------------
struct TgSideRes {
  int32_t score;
  uint32_t xend, yend, cells, n_ext, ops_off, ops_n, xclip;  // xclip: Xclip(xlen) of an empty-y extension
};
TG_HD TgSideRes tg_side_result(const TgTask* tasks, int32_t task, uint32_t xlen) {
  TgSideRes r{0, 0, 0, 0, 0, 0, 0, 0};
  if (task >= 0) {
    const TgTask& t = tasks[task];
    r.score = t.score; r.xend = t.xend; r.yend = t.yend; r.cells = t.cells; r.n_ext = 1; r.ops_off = t.ops_off; r.ops_n = t.ops_n;
  } else {
    r.xclip = xlen;  // src/swg.rs:39-55: empty x -> no ops; empty y -> [Xclip(xlen)]
  }
  return r;
}
TG_HD bool tg_ops_push_cap(TgOps& o, uint32_t cap, uint32_t kind, uint32_t run) {
  if (run == 0 && kind <= TG_OP_INS) return true;
  if (kind <= TG_OP_INS && o.n > 0 && (o.w[o.n - 1] & 7u) == kind) { o.w[o.n - 1] += run << 3; return true; }
  if (o.n >= cap) return false;
  o.w[o.n++] = kind | (run << 3);
  return true;
}
// rev(left.ops) ++ Match*len ++ right.ops (src/aligner.rs:388-394) from the stored (reversed) task operations
TG_HD bool tg_stitch_ops(const uint32_t* pool, const TgSideRes& L_, const TgSideRes& R_, uint32_t len, TgOps& out, uint32_t cap) {
  bool ok = true;
  if (L_.xclip) ok = ok && tg_ops_push_cap(out, cap, TG_OP_XCLIP, L_.xclip);
  for (uint32_t i = 0; i < L_.ops_n; i++) ok = ok && tg_ops_push_cap(out, cap, pool[L_.ops_off + i] & 7u, pool[L_.ops_off + i] >> 3);
  ok = ok && tg_ops_push_cap(out, cap, TG_OP_MATCH, len);
  for (uint32_t i = R_.ops_n; i-- > 0;) ok = ok && tg_ops_push_cap(out, cap, pool[R_.ops_off + i] & 7u, pool[R_.ops_off + i] >> 3);
  if (R_.xclip) ok = ok && tg_ops_push_cap(out, cap, TG_OP_XCLIP, R_.xclip);
  return ok;
}
// lift_tx_to_gx with a capacity check (same walk as tg_lift_tx_to_gx)
TG_HD bool tg_lift_tx_to_gx_cap(const uint32_t* te_start, const uint32_t* te_end, uint32_t e0, uint32_t e1, const TgOps& tx_ops,
                                uint32_t tx_ystart, uint32_t& g_ystart, uint32_t& g_yend, TgOps& out, uint32_t cap) {
  out.n = 0;
  uint32_t i = tx_ystart, exon_sum = 0, ex = e0;
  while (exon_sum + (TG_LDG(te_end + ex) - TG_LDG(te_start + ex)) <= i) {
    exon_sum += TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
    ex++;
  }
  g_ystart = TG_LDG(te_start + ex) + (i - exon_sum);
  bool ok = true;
  for (uint32_t k = 0; k < tx_ops.n && ok; k++) {
    uint32_t kind = tx_ops.w[k] & 7u, run = tx_ops.w[k] >> 3;
    bool consumes = kind == TG_OP_MATCH || kind == TG_OP_SUBST || kind == TG_OP_DEL;
    uint32_t units = (kind <= TG_OP_INS) ? run : 1u;
    while (units > 0 && ok) {
      uint32_t elen = TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
      if (ex + 1 < e1 && exon_sum + elen <= i) {
        exon_sum += elen;
        ex++;
        ok = ok && tg_ops_push_cap(out, cap, TG_OP_YCLIP, TG_LDG(te_start + ex) - TG_LDG(te_end + ex - 1));
        elen = TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
      }
      if (kind > TG_OP_INS) { ok = ok && tg_ops_push_cap(out, cap, kind, run); units = 0; }
      else if (!consumes) { ok = ok && tg_ops_push_cap(out, cap, kind, units); units = 0; }
      else {
        uint32_t room = (ex + 1 < e1) ? (exon_sum + elen - i) : units;
        uint32_t take = units < room ? units : room;
        if (take == 0) take = 1;
        ok = ok && tg_ops_push_cap(out, cap, kind, take);
        i += take;
        units -= take;
      }
    }
  }
  g_yend = TG_LDG(te_start + ex) + (i - exon_sum);
  return ok;
}

// Returns false when the read has to leave the round path (accepted-list or arena overflow).
template <class W>
TG_HDN bool tg_round_post(W& w, const TgAlignParams& P, TgReadState& st, const tg_seed* seeds, uint32_t n_seeds,
                          const TgHit& hit, const TgTask* tasks, const uint32_t* ops_pool, TgCand* acc, uint32_t* arena) {
  const TgIndexDev& ix = P.ix;
  const uint32_t L = st.L;
  struct { unsigned long long cells; uint32_t n_ext; } ctr{0, 0};
  // extend_left_right results per distinct problem (src/aligner.rs:377-406)
  TgAln pa[TG_PMAX];
  uint32_t pcells[TG_PMAX], pext[TG_PMAX];
  for (uint32_t k = 0; k < hit.n_prob; k++) {
    const TgProbE& e = hit.prob[k];
    TgSideRes R_ = tg_side_result(tasks, e.task_r, L - (e.q + e.len)), L_ = tg_side_result(tasks, e.task_l, e.q);
    pa[k].score = L_.score + (int32_t)e.len + R_.score;
    pa[k].ystart = (uint32_t)(e.r_abs - L_.yend);
    pa[k].yend = (uint32_t)(e.r_abs + e.len + R_.yend);
    pa[k].xstart = e.q - L_.xend;
    pa[k].xend = e.q + e.len + R_.xend;
    pcells[k] = R_.cells + L_.cells;
    pext[k] = R_.n_ext + L_.n_ext;
  }
  const TgAln gx = pa[0];
  ctr.cells += pcells[0]; ctr.n_ext += pext[0];
  bool have_tx = false;
  uint32_t best_c = 0;
  TgAln best{0, 0, 0, 0, 0};
  for (uint32_t c = 0; c < hit.n_cand; c++) {  // src/aligner.rs:237-258
    const TgCandE& ce = hit.cand[c];
    const TgProbE& e = hit.prob[ce.prob];
    TgAln ta = pa[ce.prob];
    // same offsets relative to the seed, expressed in this transcript's coordinates
    ta.ystart = ce.tr - (uint32_t)(e.r_abs - pa[ce.prob].ystart);
    ta.yend = ce.tr + (uint32_t)(pa[ce.prob].yend - e.r_abs);
    ctr.cells += pcells[ce.prob]; ctr.n_ext += pext[ce.prob];
    if (!have_tx || ta.score > best.score) { have_tx = true; best = ta; best_c = c; }
    if (ta.score >= (int32_t)L) break;
  }
  const bool exonic = have_tx && best.score >= gx.score;  // ties -> Exonic (:263)
  const int32_t s = exonic ? best.score : gx.score;
  const int32_t range = (int32_t)P.opts.multimap_score_range;
  bool keep = true;
  if (!P.opts.intron_mode && !exonic) keep = false;                                              // :146-151
  if (s < P.opts.min_aln_score || s < st.min_aln || s < st.max_aln - range) keep = false;         // :154-159
  if (keep) {
    if (st.n_acc >= TG_ACC_MAX) return false;
    const TgRef aref = ix.refs[hit.ref_id];
    TgCand c;
    tg_aln& a = c.a;
    a.ref_id = hit.ref_id; a.strand = (uint8_t)(aref.strand_rank & 1u); a.primary = 0; a.pad = 0; a.xlen = L;
    c.name_rank = aref.strand_rank >> 1;
    const uint32_t room = TG_ARENA_WORDS - st.arena_used;
    uint32_t ys, ye;
    TgOps gops{arena + st.arena_used, 0};
    uint32_t tx_n = 0;
    if (exonic) {
      const TgCandE& ce = hit.cand[best_c];
      const TgProbE& e = hit.prob[ce.prob];
      // transcript ops go to the upper half of the free arena, the lifted ops in front, then the tx ops are moved up
      const uint32_t half = room / 2;
      TgOps tops{arena + st.arena_used + half, 0};
      TgSideRes R_ = tg_side_result(tasks, e.task_r, L - (e.q + e.len)), L_ = tg_side_result(tasks, e.task_l, e.q);
      if (!tg_stitch_ops(ops_pool, L_, R_, e.len, tops, room - half)) return false;
      const uint32_t e0 = TG_LDG(ix.tx_exon_off + ce.tx_idx), e1 = TG_LDG(ix.tx_exon_off + ce.tx_idx + 1);
      if (!tg_lift_tx_to_gx_cap(ix.te_start, ix.te_end, e0, e1, tops, best.ystart, ys, ye, gops, half)) return false;
      for (uint32_t i = 0; i < tops.n; i++) arena[st.arena_used + gops.n + i] = tops.w[i];
      tx_n = tops.n;
      a.aln_type = TG_ALN_EXONIC;
      a.score = best.score; a.xstart = best.xstart; a.xend = best.xend;
      a.tx_or_gene_idx = ce.tx_idx;
      a.tx_score = best.score; a.tx_ystart = best.ystart; a.tx_yend = best.yend; a.tx_ylen = ce.tlen;
      a.tx_xstart = best.xstart; a.tx_xend = best.xend;
    } else {
      const TgProbE& e = hit.prob[0];
      TgSideRes R_ = tg_side_result(tasks, e.task_r, L - (e.q + e.len)), L_ = tg_side_result(tasks, e.task_l, e.q);
      if (!tg_stitch_ops(ops_pool, L_, R_, e.len, gops, room)) return false;
      uint32_t gene = 0, grank = 0;
      const TgStabRange gr = tg_stab_begin<W>(w, ix.gene_stab, ix.n_gene_stab, ix.gene_maxlen, gx.ystart, gx.yend);
      const bool found = tg_stab_next<W>(w, ix.gene_stab, gr, 0u, grank, gene);
      a.aln_type = found ? TG_ALN_INTRONIC : TG_ALN_INTERGENIC;
      a.tx_or_gene_idx = found ? gene : 0xFFFFFFFFu;
      a.score = gx.score; a.xstart = gx.xstart; a.xend = gx.xend;
      a.tx_score = 0; a.tx_ystart = 0; a.tx_yend = 0; a.tx_ylen = 0; a.tx_xstart = 0; a.tx_xend = 0;
      ys = gx.ystart; ye = gx.yend;
    }
    // concat_to_chr_aln (src/aligner.rs:429-449)
    const uint32_t rid2 = tg_idx_to_ref(ix.refs, ix.n_refs, ys);
    const TgRef r2 = ix.refs[rid2];
    if (r2.strand_rank & 1u) {
      a.ystart = ys - r2.start_idx;
      a.yend = ye - r2.start_idx;
    } else {
      a.ystart = (uint64_t)r2.len - (ye - r2.start_idx);
      a.yend = (uint64_t)r2.len - (ys - r2.start_idx);
      tg_ops_reverse(gops);
    }
    a.ylen = r2.len;
    a.ops_off = st.arena_used; a.ops_len = gops.n;
    a.tx_ops_off = st.arena_used + gops.n; a.tx_ops_len = tx_n;
    acc[st.n_acc++] = c;
    st.arena_used += gops.n + tx_n;
    // :162-172
    uint32_t lim = (s < 0) ? 0u : ((L + P.opts.multimap_score_range > (uint32_t)s) ? L + P.opts.multimap_score_range - (uint32_t)s : 0u);
    if (lim < st.bw) st.bw = lim;
    if (lim < st.x_drop) st.x_drop = lim;
    if (s > st.max_aln) st.max_aln = s;
  }
  st.hits++; st.cells += ctr.cells; st.n_ext += ctr.n_ext;
  // next hit: occurrences of a seed in descending SA rank, then the next seed (src/index.rs:236-253)
  if (st.rk > 1) st.rk--;
  else {
    st.si++;
    if (st.si >= n_seeds) st.status = TG_RS_DONE;
    else st.rk = seeds[st.si].count;
  }
  return true;
}

// ---- end of read (src/aligner.rs:177-187) for reads that finished on the round path ---------------------------------------
template <class W>
TG_HDN void tg_round_final(W& w, const TgAlignParams& P, const TgReadState& st, const TgCand* acc, const uint32_t* arena,
                           const TgAlignOut& out, uint32_t r) {
  uint16_t order[TG_ACC_MAX], tmp[TG_ACC_MAX];
  const uint32_t k = tg_finalize_read(acc, st.n_acc, st.max_aln, (int32_t)P.opts.multimap_score_range, order, tmp);
  unsigned long long words = 0;
  for (uint32_t i = 0; i < k; i++) words += acc[order[i]].a.ops_len + acc[order[i]].a.tx_ops_len;
  unsigned long long abase = 0, obase = 0;
  uint32_t kk = k;
  if (k > 0) {
    abase = w.atomic_add(out.alns_used, (unsigned long long)k);
    obase = w.atomic_add(out.ops_used, words);
    if (abase + k > out.alns_cap || obase + words > out.ops_cap || obase + words > 0xFFFFFFFFull) {
      w.atomic_or(out.flags, abase + k > out.alns_cap ? TG_FLAG_ALN_POOL : TG_FLAG_OPS_POOL);
      kk = 0;
    }
  }
  unsigned long long o = obase;
  for (uint32_t i = 0; i < kk; i++) {
    const TgCand& c = acc[order[i]];
    const uint32_t n1 = c.a.ops_len, n2 = c.a.tx_ops_len;
    for (uint32_t t = 0; t < n1 + n2; t++) out.ops[o + t] = arena[c.a.ops_off + t];
    tg_aln a = c.a;
    a.ops_off = (uint32_t)o;
    a.tx_ops_off = (uint32_t)(o + n1);
    if (a.aln_type != TG_ALN_EXONIC) { a.tx_ops_off = 0; a.tx_ops_len = 0; }
    a.primary = i == 0 ? 1 : 0;
    out.alns[abase + i] = a;
    o += n1 + n2;
  }
  out.read_aln_first[r] = abase;
  out.read_aln_count[r] = kk;
}
