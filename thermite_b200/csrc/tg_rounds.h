// tg_rounds.h -- the hit loop of align_read (reference src/aligner.rs:143-175) as a SPECULATIVE ROUND pipeline.
//
// The reference walks the hits of a read one after the other; every accepted hit may narrow (band_width, x_drop) for
// the hits behind it (:162-171) and raises the running maximum that later hits are filtered against (:154-159).  The
// DP result of a hit therefore depends on the hits before it only through (band_width, x_drop), and that pair changes
// rarely: typically once, at the first accepted hit.  So a round evaluates a whole BATCH of consecutive hits of every
// active read under the read's current (band_width, x_drop):
//
//   plan (thread/read)  ->  prep (thread/hit)  ->  extend (warp/task)  ->  post (thread/hit)  ->  scan (thread/read)
//
// and `scan` replays the reference's serial loop over the batch results: running maximum, accept / reject, narrowing.
// Every item carries the (band_width, x_drop) it was evaluated under and is consumed only if that is the state the
// read is really in when the scan reaches it.  When an accepted hit narrows the band, the hits of the batch BEHIND it
// are stale; scan submits them again at once, each under a PREDICTED state: the stale evaluations almost always have
// the scores the correct ones will have, so replaying accept / narrow over them tells which state every later hit will
// see.  A wrong prediction only costs another round.  Every consumed hit was evaluated with exactly the
// (band_width, x_drop) the serial loop would have used, so records and work counters are bit-identical; only wasted
// (discarded) work differs (bench: 2.80 M evaluations for 2.56 M hits, 5 rounds, 26 of 1 M reads left after round 2).
// Batch policy: 1 hit in round 0 (the first accept nearly always narrows the initial wide band), then all remaining
// hits once something was accepted, else 4, 16, 64 ... (tg_plan_batch).
//
// Accepted alignments stay where `post` wrote them (item arrays are append-only within a batch of reads) and are
// chained per read through TgItemRes::prev_acc; `final` applies src/aligner.rs:177-187.  Reads the tables cannot hold
// (more than TG_CMAX transcripts on one seed, more hits than the rounds can consume) take the single-warp kernel
// (tg_align_read in tg_core.h), which recomputes them from scratch.
//
// Everything here is expressed with the same building blocks as tg_core.h (and therefore under the same parity
// tests): prep/post are tg_align_seed_hit cut at the SwgExtend calls.
#pragma once
#include "tg_core.h"

#define TG_CMAX 12            // transcript candidates tabulated per hit
#define TG_PMAX 6             // distinct extension problems per hit (problem 0 = genome)
#define TG_MAX_ROUNDS 16
#define TG_BATCH_MAX 8192u    // hits of one read evaluated in one round
#define TG_ROUND_MAX_HITS (TG_BATCH_MAX * (TG_MAX_ROUNDS - 4))  // reads with more hits go straight to the single-warp path
#define TG_FINAL_SMALL 16u    // accepted alignments the thread-per-read finaliser sorts in local memory
#define TG_NONE 0xFFFFFFFFu

enum { TG_RS_DONE = 0, TG_RS_ACTIVE = 1, TG_RS_COMPLEX = 2, TG_RS_FINAL = 3 /* records already written */ };
enum { TG_IF_KEEP = 1, TG_IF_FAIL = 2 };
enum { TG_FLAG_TASK_POOL = 32, TG_FLAG_ITEM_POOL = 64, TG_FLAG_HOPS_POOL = 128 };  // (256: TG_FLAG_YLEN, tg_core.h)

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
  uint8_t seqsel, pad[3];
  uint32_t gkey;           // text position of the seed when both y windows are one contiguous piece of text (the genome
                           // problem; a transcript problem whose windows stay inside one exon), else TG_NONE
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
struct TgItemRes {  // one (read, hit) evaluation
  uint32_t read, hit;    // hit = flat index into the read's hit sequence (Index::all_smems order)
  int32_t score;         // score of the GenomeAlignment align_seed_hit returns
  uint32_t cells, n_ext; // work the reference does for this hit
  uint32_t flags;        // TG_IF_*
  uint32_t prev_acc;     // previous accepted item of the same read (TG_NONE: first)
  uint32_t state;        // (band_width | x_drop << 16) the item is evaluated under (tg_pack_state)
};
struct TgReadState {
  uint32_t L;
  int32_t min_aln, max_aln;
  uint32_t bw, x_drop;
  uint32_t n_hits, next_hit;        // flat hit cursor
  uint32_t batch_first, batch_n;    // items of the current round
  uint32_t n_acc, acc_head;         // accepted alignments (chained through TgItemRes::prev_acc, newest first)
  uint32_t planned;                 // the next batch was already submitted by scan (predicted states)
  uint32_t status;
  uint32_t hits, n_ext;             // work counters of the consumed hits
  unsigned long long cells;
};

TG_HD uint32_t tg_pack_state(uint32_t bw, uint32_t x_drop) { return (bw & 0xFFFFu) | (x_drop << 16); }
TG_HD const uint64_t* tg_seq_of(const TgIndexDev& ix, uint32_t seqsel) { return seqsel ? ix.txseq4 : ix.text4; }

// src/aligner.rs:130-138.  Returns false when the read has more hits than the round path handles.
TG_HD bool tg_read_state_init(TgReadState& s, uint32_t L, const tg_opts& o, uint32_t n_seeds, const tg_seed* seeds) {
  float prod = o.min_aln_score_percent * (float)L;
  int32_t pct_score = (int32_t)prod;
  s.L = L;
  s.min_aln = pct_score > o.min_aln_score ? pct_score : o.min_aln_score;
  s.max_aln = s.min_aln;
  s.bw = (s.min_aln < 0) ? 0u : (L > (uint32_t)s.min_aln ? L - (uint32_t)s.min_aln : 0u);
  s.x_drop = s.bw;
  unsigned long long hits = 0;
  for (uint32_t i = 0; i < n_seeds; i++) hits += seeds[i].count;
  s.next_hit = 0;
  s.batch_first = 0; s.batch_n = 0;
  s.n_acc = 0; s.acc_head = TG_NONE; s.planned = 0;
  s.hits = 0; s.n_ext = 0; s.cells = 0;
  s.status = hits ? TG_RS_ACTIVE : TG_RS_DONE;
  s.n_hits = hits > TG_ROUND_MAX_HITS ? 0u : (uint32_t)hits;
  return hits <= TG_ROUND_MAX_HITS;
}

// hits of the batch a read submits in round `round`
TG_HD uint32_t tg_plan_batch(const TgReadState& st, uint32_t round) {
  const uint32_t rem = st.n_hits - st.next_hit;
  uint32_t b = TG_BATCH_MAX;
  if (st.n_acc == 0) {  // nothing accepted yet: the state is still the initial one and the next accept will narrow it
    const uint32_t sh = 2 * round < 13 ? 2 * round : 13;
    b = 1u << sh;
  }
  return rem < b ? rem : b;
}

// flat hit index -> (seed, occurrence): seeds in order, occurrences of a seed in DESCENDING suffix-array rank
// (src/index.rs:236-253)
TG_HD void tg_hit_locate(const tg_seed* seeds, uint32_t n_seeds, uint32_t h, uint32_t& si, uint32_t& rk) {
  uint32_t s = 0;
  while (s + 1 < n_seeds && h >= seeds[s].count) { h -= seeds[s].count; s++; }
  si = s;
  rk = seeds[s].count - 1 - h;
}

// Do two transcripts run through the same pieces of text for `n` symbols, starting at offset so_a of exon se_a (resp.
// so_b of se_b) -- both the same text position -- and walking towards higher (dir = +1; e*_end = one past the last exon)
// or lower (dir = -1; e*_end = the first exon; the walk starts just below the offsets) transcript coordinates?
TG_HD bool tg_same_tx_pieces(const uint32_t* te_start, const uint32_t* te_end, uint32_t se_a, uint32_t so_a, uint32_t ea_end,
                             uint32_t se_b, uint32_t so_b, uint32_t eb_end, uint32_t n, int dir) {
  uint32_t ea = se_a, eb = se_b, oa = so_a, ob = so_b;
  while (n > 0) {
    const uint32_t sa_ = TG_LDG(te_start + ea), sb_ = TG_LDG(te_start + eb);
    uint32_t roomA, roomB;
    if (dir > 0) {
      if (sa_ + oa != sb_ + ob) return false;           // different text position
      roomA = TG_LDG(te_end + ea) - sa_ - oa;
      roomB = TG_LDG(te_end + eb) - sb_ - ob;
    } else {
      if (sa_ + oa != sb_ + ob) return false;
      roomA = oa; roomB = ob;                            // symbols below the offset inside the exon
    }
    const uint32_t takeA = roomA < n ? roomA : n, takeB = roomB < n ? roomB : n;
    if (takeA != takeB) return false;                    // one transcript leaves its exon earlier
    n -= takeA;
    if (n == 0) return true;
    if (dir > 0) {
      ea++; eb++;
      if (ea >= ea_end || eb >= eb_end) return false;    // (cannot happen: the windows lie inside the transcripts)
      oa = 0; ob = 0;
    } else {
      if (ea == ea_end || eb == eb_end) return false;
      ea--; eb--;
      oa = TG_LDG(te_end + ea) - TG_LDG(te_start + ea);
      ob = TG_LDG(te_end + eb) - TG_LDG(te_start + eb);
    }
  }
  return true;
}

// ---- prep: everything of align_seed_hit (src/aligner.rs:198-258) that happens before a SwgExtend call ------------
// Tabulates the genome problem and one candidate per transcript that the seed's exon stab yields (ALL of them: the
// reference's early `break` at a perfect transcript is applied in post), maps identical problems onto each other and
// emits one task per non-trivial extension.  Returns false when a table overflows (the read becomes "complex").
template <class W>
TG_HDN bool tg_item_prep(W& w, const TgAlignParams& P, const uint64_t* rp, const TgReadState& st, uint32_t state,
                         const tg_seed* seeds, uint32_t n_seeds, uint32_t read, uint32_t flat_hit, TgHit& hit, TgTask* tasks,
                         unsigned long long* task_ctr, unsigned long long task_cap, int* flags) {
  const TgIndexDev& ix = P.ix;
  const uint32_t L = st.L;
  const uint32_t s_bw = state & 0xFFFFu, s_xd = state >> 16;  // the state this item is evaluated under
  uint32_t si, rk;
  tg_hit_locate(seeds, n_seeds, flat_hit, si, rk);
  const tg_seed sd = seeds[si];
  const uint32_t ref_idx = sd.direct ? sd.sa_lo : TG_LDG(ix.sa + sd.sa_lo + rk);
  const uint32_t q = sd.query_idx, len = sd.len, bw = s_bw;
  hit.ref_idx = ref_idx; hit.q = q; hit.len = len; hit.bw = bw; hit.x_drop = (int32_t)s_xd;
  hit.ref_id = tg_idx_to_ref(ix.refs, ix.n_refs, ref_idx);
  const TgRef aref = ix.refs[hit.ref_id];
  const uint64_t span = (uint64_t)L + bw;
  uint64_t seq_start = ref_idx > span ? ref_idx - span : 0;
  if (seq_start < aref.start_idx) seq_start = aref.start_idx;
  uint64_t seq_end = (uint64_t)ref_idx + len + L + bw;
  if (seq_end > (uint64_t)aref.end_idx - 1) seq_end = (uint64_t)aref.end_idx - 1;
  // the problem table is read back many times while it is built: keep it in thread-local storage (L1) and store it to the
  // item once at the end, instead of reloading global memory after every store
  TgProbE prob[TG_PMAX];
  uint32_t n_prob = 1, n_cand = 0;
  prob[0].lo_abs = seq_start; prob[0].hi_abs = seq_end; prob[0].r_abs = ref_idx;
  prob[0].q = q; prob[0].len = len; prob[0].seqsel = 0; prob[0].gkey = ref_idx;
  prob[0].task_r = -1; prob[0].task_l = -1;
  // per problem: window lengths; for transcript problems the seed's text position, its exon / offset and the exon range
  uint32_t p_ncR[TG_PMAX], p_ncL[TG_PMAX], p_gpos[TG_PMAX], p_se[TG_PMAX], p_so[TG_PMAX], p_e0[TG_PMAX], p_e1[TG_PMAX];
  {
    TgProblem pg{nullptr, seq_start, seq_end, ref_idx, q, len};
    tg_problem_windows(pg, L, bw, p_ncR[0], p_ncL[0]);
    p_gpos[0] = ref_idx; p_se[0] = 0; p_so[0] = 0; p_e0[0] = 0; p_e1[0] = 0;
  }
  const TgStabRange xr = tg_stab_begin<W>(w, ix.exon_stab, ix.n_exon_stab, ix.exon_maxlen, ref_idx, ref_idx + len);
  // The intersecting exons in find() order.  One pass over the start-sorted range collects them, sorted by rank, into a
  // short list (a range holds every exon that starts within exon_maxlen before the seed: hundreds on a dense annotation,
  // and rescanning it for every next rank would cost range x candidates loads); only when more than TG_CMAX intersect
  // does the walk fall back to one scan per rank (it then ends in "complex" at the TG_CMAX + 1st candidate as before).
  uint32_t m_rank[TG_CMAX], m_tx[TG_CMAX], n_m = 0;
  bool listed = true;
  for (uint32_t i = xr.lo; i < xr.hi; i++) {
    const TgStab e = tg_stab_load(ix.exon_stab, i);
    if (!(e.start < xr.qe && xr.qs < e.end)) continue;
    if (n_m == TG_CMAX) { listed = false; break; }
    uint32_t k = n_m++;
    for (; k > 0 && m_rank[k - 1] > e.rank; k--) { m_rank[k] = m_rank[k - 1]; m_tx[k] = m_tx[k - 1]; }
    m_rank[k] = e.rank; m_tx[k] = e.data;
  }
  uint32_t next_rank = 0;
  for (uint32_t ci = 0;; ci++) {
    uint32_t tx_idx = 0, xrank = 0;
    if (listed) {
      if (ci >= n_m) break;
      tx_idx = m_tx[ci];
    } else {
      if (!tg_stab_next<W>(w, ix.exon_stab, xr, next_rank, xrank, tx_idx)) break;
      next_rank = xrank + 1;
    }
    const uint32_t e0 = TG_LDG(ix.tx_exon_off + tx_idx), e1 = TG_LDG(ix.tx_exon_off + tx_idx + 1);
    const uint64_t t0 = TG_LDG(ix.tx_seq_off + tx_idx), t1 = TG_LDG(ix.tx_seq_off + tx_idx + 1);
    uint32_t tr = 0, tq = 0, tl = 0;
    uint32_t lift_e = e0, lift_sum = 0;  // the exon holding the lifted seed start and its transcript offset
    if (!tg_lift_mem_to_tx(ix.te_start, ix.te_end, e0, e1, ref_idx, q, len, tr, tq, tl, &lift_e, &lift_sum)) continue;
    tl += tg_match_fwd(rp, tq + tl, L, ix.txseq4, t0 + tr + tl, t1);
    {
      uint32_t back = tg_match_bwd(rp, tq, ix.txseq4, t0 + tr, t0);
      tr -= back; tq -= back; tl += back;
    }
    if (n_cand >= TG_CMAX) return false;
    TgProblem pt{ix.txseq4, t0, t1, t0 + tr, tq, tl};
    // Identical DP problems are evaluated once, decided structurally (no symbol comparison): transcript sequences are
    // spliced text, so two problems whose seed starts at the same text position, with the same seed and window
    // lengths, and whose windows run through the SAME pieces of text are the same problem.  A transcript problem whose
    // windows stay inside ONE exon is a contiguous piece of text like the genome problem (gkey); windows that cross
    // junctions are compared exon by exon against the other junction-crossing problems.  A missed match merely costs a
    // duplicate evaluation with the same result.
    uint32_t ncRt, ncLt;
    tg_problem_windows(pt, L, bw, ncRt, ncLt);
    uint32_t gkey = TG_NONE, gpos = TG_NONE, se = e0, so = 0;
    {
      // the exon holding the (extended) seed start: the backward match can only have moved it towards the transcript
      // start, so walk back from the exon lift_mem_to_tx found instead of forward from the first one
      uint32_t e = lift_e, exon_sum = lift_sum;
      while (tr < exon_sum) { e--; exon_sum -= TG_LDG(ix.te_end + e) - TG_LDG(ix.te_start + e); }
      const uint32_t es = TG_LDG(ix.te_start + e), elen = TG_LDG(ix.te_end + e) - es;
      se = e; so = tr - exon_sum; gpos = es + so;
      if (so >= ncLt && (uint64_t)tr + tl + ncRt <= (uint64_t)exon_sum + elen) gkey = gpos;
    }
    uint32_t pi = n_prob;
    for (uint32_t k = 0; k < n_prob; k++) {
      const TgProbE& e = prob[k];
      if (e.q != tq || e.len != tl || p_ncR[k] != ncRt || p_ncL[k] != ncLt) continue;
      if (gkey != TG_NONE) {
        if (e.gkey == gkey) { pi = k; break; }
      } else if (e.gkey == TG_NONE && p_gpos[k] == gpos &&
                 tg_same_tx_pieces(ix.te_start, ix.te_end, se, so, e1, p_se[k], p_so[k], p_e1[k], tl + ncRt, +1) &&
                 tg_same_tx_pieces(ix.te_start, ix.te_end, se, so, e0, p_se[k], p_so[k], p_e0[k], ncLt, -1)) {
        pi = k;
        break;
      }
    }
    if (pi == n_prob) {
      if (n_prob >= TG_PMAX) return false;
      TgProbE& e = prob[n_prob++];
      e.lo_abs = t0; e.hi_abs = t1; e.r_abs = t0 + tr; e.q = tq; e.len = tl; e.seqsel = 1; e.gkey = gkey;
      e.task_r = -1; e.task_l = -1;
      const uint32_t k = n_prob - 1;
      p_ncR[k] = ncRt; p_ncL[k] = ncLt; p_gpos[k] = gpos; p_se[k] = se; p_so[k] = so; p_e0[k] = e0; p_e1[k] = e1;
    }
    TgCandE& c = hit.cand[n_cand++];
    c.t0 = t0; c.tx_idx = tx_idx; c.prob = pi; c.tr = tr; c.tlen = (uint32_t)(t1 - t0);
  }
  // tasks (the windows are those of extend_left_right, src/aligner.rs:360-375)
  uint32_t need = 0;
  for (uint32_t k = 0; k < n_prob; k++) {
    TgProbE& e = prob[k];
    const uint32_t ncR = p_ncR[k], ncL = p_ncL[k];
    e.task_r = ncR ? 0 : -1;
    e.task_l = ncL ? 0 : -1;
    need += (ncR ? 1u : 0u) + (ncL ? 1u : 0u);
  }
  unsigned long long base = need ? w.atomic_add(task_ctr, (unsigned long long)need) : 0;
  if (base + need > task_cap) {
    w.atomic_or(flags, TG_FLAG_TASK_POOL);
    return false;
  }
  for (uint32_t k = 0; k < n_prob; k++) {
    TgProbE& e = prob[k];
    const uint32_t xr_len = L - (e.q + e.len);
    if (e.task_r == 0) {
      TgTask& t = tasks[base];
      e.task_r = (int32_t)base++;
      t.read = read; t.side = 0; t.seqsel = e.seqsel; t.xoff = e.q + e.len; t.xlen = xr_len;
      t.y0 = e.r_abs + e.len;
      uint64_t yl = e.hi_abs - t.y0;
      t.ylen = (uint32_t)(yl > (uint64_t)xr_len + bw ? (uint64_t)xr_len + bw + 1 : yl);
      t.bw = bw; t.x_drop = (int32_t)s_xd;
      t.pad0 = 0; t.pad1 = 0;  // (pad0 = 1 marks a task the thread kernels must not take: the task buffer is shared with tg_swg_extend_batch)
    }
    if (e.task_l == 0) {
      TgTask& t = tasks[base];
      e.task_l = (int32_t)base++;
      t.read = read; t.side = 1; t.seqsel = e.seqsel; t.xoff = 0; t.xlen = e.q;
      t.y0 = e.r_abs;
      uint64_t ys0 = (e.r_abs - e.lo_abs > span) ? e.r_abs - span : e.lo_abs;
      uint64_t yl = e.r_abs - ys0;
      t.ylen = (uint32_t)(yl > (uint64_t)e.q + bw ? (uint64_t)e.q + bw + 1 : yl);
      t.bw = bw; t.x_drop = (int32_t)s_xd;
      t.pad0 = 0; t.pad1 = 0;
    }
  }
  hit.n_prob = n_prob; hit.n_cand = n_cand;
  for (uint32_t k = 0; k < n_prob; k++) hit.prob[k] = prob[k];
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

// ---- post: the rest of align_seed_hit (src/aligner.rs:240-314) -------------------------------------------------------
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
// rev(left.ops) ++ Match*len ++ right.ops (src/aligner.rs:388-394) from the stored (reversed) task operations.
// `out` must have room for L.ops_n + R.ops_n + 3 words.
TG_HD void tg_stitch_ops(const uint32_t* pool, const TgSideRes& L_, const TgSideRes& R_, uint32_t len, TgOps& out) {
  // tg_ops_push word by word, with the word under construction in a register: the stored task operations are already
  // run-length encoded, so runs only merge at the two junctions, and re-reading the previous word from memory for every
  // push made this the hottest spot of post (28 % of its instructions, profiles/r1_prep_post_hot_lines.txt)
  uint32_t* w = out.w;
  uint32_t n = out.n, cur = 0;
  bool have = false;
  if (n > 0) { cur = w[--n]; have = true; }
  auto push = [&](uint32_t kind, uint32_t run) {
    if (run == 0 && kind <= TG_OP_INS) return;
    if (have && kind <= TG_OP_INS && (cur & 7u) == kind) { cur += run << 3; return; }
    if (have) w[n++] = cur;
    cur = kind | (run << 3);
    have = true;
  };
  if (L_.xclip) push(TG_OP_XCLIP, L_.xclip);
  for (uint32_t i = 0; i < L_.ops_n; i++) { const uint32_t v = pool[L_.ops_off + i]; push(v & 7u, v >> 3); }
  push(TG_OP_MATCH, len);
  for (uint32_t i = R_.ops_n; i-- > 0;) { const uint32_t v = pool[R_.ops_off + i]; push(v & 7u, v >> 3); }
  if (R_.xclip) push(TG_OP_XCLIP, R_.xclip);
  if (have) w[n++] = cur;
  out.n = n;
}

struct TgHopsPool {  // operations of the evaluated hits (append-only within a batch of reads)
  uint32_t* w;
  unsigned long long* used;
  unsigned long long cap;
};

// Evaluates one item: score / type of the GenomeAlignment align_seed_hit would return, and -- when the alignment
// passes every filter that does not depend on hits in front of it -- the full candidate record with its operations.
template <class W>
TG_HDN void tg_item_post(W& w, const TgAlignParams& P, const TgReadState& st, const TgHit& hit, const TgTask* tasks,
                         const uint32_t* dp_ops, TgItemRes& ir, TgCand& cand, const TgHopsPool& hp, int* flags) {
  const TgIndexDev& ix = P.ix;
  const uint32_t L = st.L;
  unsigned long long cells = 0;
  uint32_t n_ext = 0;
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
  cells += pcells[0]; n_ext += pext[0];
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
    cells += pcells[ce.prob]; n_ext += pext[ce.prob];
    if (!have_tx || ta.score > best.score) { have_tx = true; best = ta; best_c = c; }
    if (ta.score >= (int32_t)L) break;
  }
  const bool exonic = have_tx && best.score >= gx.score;  // ties -> Exonic (:263)
  const int32_t s = exonic ? best.score : gx.score;
  const int32_t range = (int32_t)P.opts.multimap_score_range;
  ir.score = s;
  ir.cells = (uint32_t)cells;
  ir.n_ext = n_ext;
  ir.prev_acc = TG_NONE;
  bool keep = true;
  if (!P.opts.intron_mode && !exonic) keep = false;                                              // :146-151
  // :154-159 against the maximum at the START of the round (a lower bound of the running maximum `scan` applies)
  if (s < P.opts.min_aln_score || s < st.min_aln || s < st.max_aln - range) keep = false;
  ir.flags = 0;
  if (!keep) return;
  const TgRef aref = ix.refs[hit.ref_id];
  tg_aln& a = cand.a;
  a.ref_id = hit.ref_id; a.strand = (uint8_t)(aref.strand_rank & 1u); a.primary = 0; a.pad = 0; a.xlen = L;
  cand.name_rank = aref.strand_rank >> 1;
  const TgProbE& e = exonic ? hit.prob[hit.cand[best_c].prob] : hit.prob[0];
  const TgSideRes R_ = tg_side_result(tasks, e.task_r, L - (e.q + e.len)), L_ = tg_side_result(tasks, e.task_l, e.q);
  const uint32_t stitched_max = L_.ops_n + R_.ops_n + 3;
  uint32_t e0 = 0, e1 = 0;
  if (exonic) {
    e0 = TG_LDG(ix.tx_exon_off + hit.cand[best_c].tx_idx);
    e1 = TG_LDG(ix.tx_exon_off + hit.cand[best_c].tx_idx + 1);
  }
  // lifted ops: every exon boundary adds a Yclip and may split a run
  const uint32_t lifted_max = exonic ? stitched_max + 2 * (e1 - e0) + 2 : 0;
  const unsigned long long base = w.atomic_add(hp.used, (unsigned long long)(stitched_max + lifted_max));
  if (base + stitched_max + lifted_max > hp.cap || base + stitched_max + lifted_max > 0xFFFFFFFFull) {
    w.atomic_or(flags, TG_FLAG_HOPS_POOL);
    return;
  }
  uint32_t ys, ye;
  if (exonic) {
    const TgCandE& ce = hit.cand[best_c];
    TgOps gops{hp.w + base, 0}, tops{hp.w + base + lifted_max, 0};
    tg_stitch_ops(dp_ops, L_, R_, e.len, tops);
    tg_lift_tx_to_gx(ix.te_start, ix.te_end, e0, e1, tops, best.ystart, ys, ye, gops);
    a.aln_type = TG_ALN_EXONIC;
    a.score = best.score; a.xstart = best.xstart; a.xend = best.xend;
    a.tx_or_gene_idx = ce.tx_idx;
    a.tx_score = best.score; a.tx_ystart = best.ystart; a.tx_yend = best.yend; a.tx_ylen = ce.tlen;
    a.tx_xstart = best.xstart; a.tx_xend = best.xend;
    a.ops_off = (uint32_t)base; a.ops_len = gops.n;
    a.tx_ops_off = (uint32_t)(base + lifted_max); a.tx_ops_len = tops.n;
  } else {
    TgOps gops{hp.w + base, 0};
    tg_stitch_ops(dp_ops, L_, R_, e.len, gops);
    uint32_t gene = 0, grank = 0;
    const TgStabRange gr = tg_stab_begin<W>(w, ix.gene_stab, ix.n_gene_stab, ix.gene_maxlen, gx.ystart, gx.yend);
    const bool found = tg_stab_next<W>(w, ix.gene_stab, gr, 0u, grank, gene);
    a.aln_type = found ? TG_ALN_INTRONIC : TG_ALN_INTERGENIC;
    a.tx_or_gene_idx = found ? gene : 0xFFFFFFFFu;
    a.score = gx.score; a.xstart = gx.xstart; a.xend = gx.xend;
    a.tx_score = 0; a.tx_ystart = 0; a.tx_yend = 0; a.tx_ylen = 0; a.tx_xstart = 0; a.tx_xend = 0;
    ys = gx.ystart; ye = gx.yend;
    a.ops_off = (uint32_t)base; a.ops_len = gops.n;
    a.tx_ops_off = 0; a.tx_ops_len = 0;
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
    TgOps gops{hp.w + base, a.ops_len};
    tg_ops_reverse(gops);
  }
  a.ylen = r2.len;
  if (r2.len != aref.len) w.atomic_or(flags, TG_FLAG_YLEN);  // never: compact records derive ylen from ref_id
  ir.flags = TG_IF_KEEP;
}

// ---- scan: the serial part of the reference's loop (src/aligner.rs:146-174) over the batch of one read -----------------
// Items are consumed while they were evaluated under the state the read is really in.  At the first item that was not
// (an accepted hit in front of it narrowed the band), the rest of the batch -- and whatever hits lie behind it -- is
// submitted again right here, each item under a PREDICTED state: the stale evaluations almost always have the scores the
// correct ones will have, so replaying accept / narrow over them predicts the state every later hit will see.  The next
// scan validates every item against the true state again, so a wrong prediction only costs another round.
// Returns false when the read has to leave the round path (an item of the batch overflowed a table).
TG_HD uint32_t tg_narrow_limit(const tg_opts& o, uint32_t L, int32_t s) {
  // :162-171 (`score as usize` wraps for negative scores => saturating_sub gives 0)
  return (s < 0) ? 0u : ((L + o.multimap_score_range > (uint32_t)s) ? L + o.multimap_score_range - (uint32_t)s : 0u);
}
template <class W>
TG_HDN bool tg_scan_read(W& w, const tg_opts& o, TgReadState& st, TgItemRes* ires, uint32_t read,
                         unsigned long long* items_used, unsigned long long item_cap, int* flags) {
  const int32_t range = (int32_t)o.multimap_score_range;
  uint32_t consumed = 0;
  bool cut = false;
  for (uint32_t i = 0; i < st.batch_n; i++) {
    const uint32_t item = st.batch_first + i;
    TgItemRes& ir = ires[item];
    if (ir.flags & TG_IF_FAIL) return false;
    if (ir.state != tg_pack_state(st.bw, st.x_drop)) { cut = true; break; }  // evaluated under a state the read is not in
    consumed = i + 1;
    st.hits++; st.cells += ir.cells; st.n_ext += ir.n_ext;
    if (!(ir.flags & TG_IF_KEEP)) continue;
    const int32_t s = ir.score;
    if (s < st.max_aln - range) continue;  // :154-159 against the running maximum
    ir.prev_acc = st.acc_head;
    st.acc_head = item;
    st.n_acc++;
    const uint32_t lim = tg_narrow_limit(o, st.L, s);
    if (lim < st.bw) st.bw = lim;
    if (lim < st.x_drop) st.x_drop = lim;
    if (s > st.max_aln) st.max_aln = s;
  }
  const uint32_t rest = st.batch_n - consumed;  // stale items of this batch
  const uint32_t old_first = st.batch_first + consumed;
  st.next_hit += consumed;
  st.batch_n = 0; st.planned = 0;
  if (st.next_hit >= st.n_hits) { st.status = TG_RS_DONE; return true; }
  if (!cut) return true;
  // re-submit: the stale items under predicted states, then the hits behind the batch under the last prediction
  const uint32_t remaining = st.n_hits - st.next_hit;
  const uint32_t n_new = remaining < TG_BATCH_MAX ? remaining : TG_BATCH_MAX;
  const unsigned long long base = w.atomic_add(items_used, (unsigned long long)n_new);
  if (base + n_new > item_cap) {
    w.atomic_or(flags, TG_FLAG_ITEM_POOL);
    return true;
  }
  uint32_t pbw = st.bw, pxd = st.x_drop;
  int32_t pmax = st.max_aln;
  for (uint32_t t = 0; t < n_new; t++) {
    TgItemRes& nw = ires[base + t];
    nw.read = read; nw.hit = st.next_hit + t; nw.flags = 0; nw.prev_acc = TG_NONE;
    nw.state = tg_pack_state(pbw, pxd);
    if (t < rest) {
      const TgItemRes old = ires[old_first + t];
      if ((old.flags & TG_IF_KEEP) && old.score >= pmax - range) {
        const uint32_t lim = tg_narrow_limit(o, st.L, old.score);
        if (lim < pbw) pbw = lim;
        if (lim < pxd) pxd = lim;
        if (old.score > pmax) pmax = old.score;
      }
    }
  }
  st.batch_first = (uint32_t)base; st.batch_n = n_new; st.planned = 1;
  return true;
}

// ---- end of read (src/aligner.rs:177-187) for reads that finished on the round path -----------------------------------
// `idx` / `tmp` / `items`: scratch of n_acc entries each.  cands[] / ires[] are the item arrays.
template <class T>
TG_HD uint32_t tg_finalize_items(const TgCand* cands, const uint32_t* items, uint32_t n, int32_t max_aln_score, int32_t range,
                                 T* order, T* tmp) {
  // same algorithm as tg_finalize_read, candidates addressed through `items`
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++)
    if (cands[items[i]].a.score >= max_aln_score - range) order[m++] = (T)i;
  if (m == 0) return 0;
  tg_merge_sort_idx(order, tmp, m, [&](T x, T y) {
    const TgCand& a = cands[items[x]];
    const TgCand& b = cands[items[y]];
    if (a.name_rank != b.name_rank) return a.name_rank < b.name_rank;
    if (a.a.strand != b.a.strand) return a.a.strand < b.a.strand;
    return a.a.ystart < b.a.ystart;
  });
  uint32_t k = 0;
  uint64_t max_end = 0;
  for (uint32_t i = 0; i < m; i++) {
    const TgCand& cc = cands[items[order[i]]];
    bool fresh = k == 0 || cc.a.ystart >= max_end || cc.name_rank != cands[items[tmp[k - 1]]].name_rank ||
                 cc.a.strand != cands[items[tmp[k - 1]]].a.strand;
    if (fresh) {
      max_end = cc.a.yend;
      tmp[k++] = order[i];
    } else {
      if (cc.a.score > cands[items[tmp[k - 1]]].a.score) tmp[k - 1] = order[i];
      uint64_t ce = cands[items[tmp[k - 1]]].a.yend;
      if (ce > max_end) max_end = ce;
    }
  }
  for (uint32_t i = 0; i < k; i++) order[i] = tmp[i];
  tg_merge_sort_idx(order, tmp, k, [&](T x, T y) { return cands[items[x]].a.score > cands[items[y]].a.score; });
  return k;
}

// Writes the output records of one read.  Serial; `items`, `order`, `tmp` hold st.n_acc entries.
template <class W, class T>
TG_HDN void tg_round_final(W& w, const TgAlignParams& P, const TgReadState& st, const TgCand* cands, const TgItemRes* ires,
                           const uint32_t* hops, uint32_t* items, T* order, T* tmp, const TgAlignOut& out,
                           uint32_t r) {
  // accepted items in acceptance order (the chain runs newest -> oldest)
  uint32_t it = st.acc_head;
  for (uint32_t i = st.n_acc; i-- > 0;) { items[i] = it; it = ires[it].prev_acc; }
  const uint32_t k = tg_finalize_items(cands, items, st.n_acc, st.max_aln, (int32_t)P.opts.multimap_score_range, order, tmp);
  unsigned long long words = 0;
  for (uint32_t i = 0; i < k; i++) words += cands[items[order[i]]].a.ops_len + cands[items[order[i]]].a.tx_ops_len;
  unsigned long long abase = 0, obase = 0;
  uint32_t kk = k;
  if (k > 0) {
    abase = w.atomic_add(out.alns_used, (unsigned long long)k);
    obase = w.atomic_add(out.ops_used, words);
    if (abase + k > out.alns_cap || obase + words > out.ops_cap || obase + words + out.ops_base > 0xFFFFFFFFull) {
      w.atomic_or(out.flags, abase + k > out.alns_cap ? TG_FLAG_ALN_POOL : TG_FLAG_OPS_POOL);
      kk = 0;
    }
  }
  unsigned long long o = obase;
  for (uint32_t i = 0; i < kk; i++) {
    const TgCand& c = cands[items[order[i]]];
    const uint32_t n1 = c.a.ops_len, n2 = c.a.tx_ops_len;
    for (uint32_t t = 0; t < n1; t++) out.ops[o + t] = hops[c.a.ops_off + t];
    for (uint32_t t = 0; t < n2; t++) out.ops[o + n1 + t] = hops[c.a.tx_ops_off + t];
    tg_aln a = c.a;
    a.ops_off = (uint32_t)o;
    a.tx_ops_off = (uint32_t)(o + n1);
    if (a.aln_type != TG_ALN_EXONIC) { a.tx_ops_off = 0; a.tx_ops_len = 0; }
    a.primary = i == 0 ? 1 : 0;
    tg_out_write_aln(out, abase + i, a);
    o += n1 + n2;
  }
  tg_out_write_read(out, r, abase, kk);
}
