// tg_dpt.h -- SwgExtend::extend / trace (reference src/swg.rs:31-207) with ONE THREAD PER EXTENSION.
//
// The warp-cooperative fills of tg_core.h spend most of their issue slots on shuffles, pipeline fill/drain and idle
// lanes (a 91-bp read gives x <= 71 rows).  Here a thread owns a whole extension: the band of the current column lives
// in REGISTERS (template parameter WB = band slots, fully unrolled), there is no communication, and 32 independent
// extensions of similar shape (the caller sorts tasks by class and column count) keep all lanes busy.
//
//   * slots: in the first bw columns (src/swg.rs:75-113, "phase 1", quirk Q2) slot b is row b; afterwards
//     (:116-154, "phase 2") slot b is row (j - bw) + b, i.e. the window slides down one row per column.  The update is
//     done IN PLACE in ascending slot order: new[b] needs old[b+1] (same row, previous column), old[b] (diagonal) and
//     new[b-1] (row above), so no register is ever moved.
//   * match scores come from a bit profile: for each symbol c a 128-bit mask of the x positions holding c; one funnel
//     shift per column aligns the mask of y[j-1] with the slots, a cell tests one (static) bit.
//   * column maximum and its FIRST row come from one max over keys  D * 128 + (127 - slot)  (src/swg.rs:101-104 strict
//     '>' in row order); the x-drop test (:110-112) and the optional bound stop run once per column.
//   * trace: 2 bits per slot {0 diag, 1 Del, 2 Ins} (tie priority diag > Del > Ins, src/swg.rs:226-240), TW words per
//     column, written to a per-thread strided buffer; traceback re-derives Match/Subst from the profile.
//
// Limits (the caller routes everything else to the warp kernels): xlen <= TG_DPT_MAX_X, rows in band
// min(2*bw, xlen) + 1 <= WB <= TG_DPT_MAX_WB, x_drop >= bw.
//
// Written as portable C++ (like tg_core.h) so that csrc/hosttest.cpp can run the very same code on the CPU.
#pragma once
#include "tg_core.h"

#define TG_DPT_MAX_X 128
#define TG_DPT_MAX_WB 80
#define TG_DPT_MIN (-(1 << 20))  // "minus infinity": far below any real score, far from overflowing the packed keys
#define TG_DPT_NCLS 12           // class 0: not eligible; class 1: WB = 4; class c >= 2: WB = 8 * (c - 1)

TG_HD int tg_dpt_max(int a, int b) { return a > b ? a : b; }
TG_HD int tg_dpt_min(int a, int b) { return a < b ? a : b; }
TG_HD int tg_dpt_max3(int a, int b, int c) { return tg_dpt_max(tg_dpt_max(a, b), c); }

// a * b + c, forced onto the FMA pipe on the device (ptxas otherwise picks IADD3 / LEA on the busier ALU pipe)
TG_HD int tg_dpt_mad(int a, int b, int c) {
#ifdef __CUDA_ARCH__
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#else
  return a * b + c;
#endif
}

// band-slot class of an extension (0 = not eligible for the thread kernel)
TG_HD int tg_dpt_class(int xlen, int bw, int x_drop) {
  if (xlen > TG_DPT_MAX_X || x_drop < bw) return 0;
  const int rows = (2 * bw < xlen ? 2 * bw : xlen) + 1;
  if (rows > TG_DPT_MAX_WB) return 0;
  return rows <= 4 ? 1 : 1 + ((rows + 7) >> 3);
}
// slots of class `cls` and the smallest number of band rows an extension of that class has
TG_HD constexpr int tg_dpt_wb(int cls) { return cls == 1 ? 4 : 8 * (cls - 1); }
TG_HD constexpr int tg_dpt_min_rows(int wb) { return wb == 4 ? 1 : (wb == 8 ? 5 : wb - 7); }

struct TgDptMem {
  uint32_t* msk;     // match profile: word (sym * 4 + k) at msk[(sym * 4 + k) * mstride], sym in 0..7
  uint32_t mstride;
  uint32_t* tr;      // trace: word (col * TW + k) at tr[(col * TW + k) * tstride]
  uint32_t tstride;
  int one, k128;     // the constants 1 and 128 as RUNTIME values (kernel parameters): multiply-adds by them cannot be
                     // strength-reduced to IADD / LEA, so they issue on the FMA pipe instead of the saturated ALU pipe
};

// y symbols in extension order, 16 at a time (side 0: seq[y0 + t]; side 1: seq[y0 - 1 - t]).  The word after the one
// in use is fetched ahead (the fill walks forward): a refill is a random HBM access the thread would otherwise wait for.
struct TgDptY {
  const uint64_t* seq;
  uint64_t y0;
  int ncols, side;
  uint64_t word;
  int need;
  uint64_t word_next;
  int need_next, t_next;  // t_next < 0: nothing fetched ahead
  TG_HD void load(int t, uint64_t& w, int& nd) const {
    if (side == 0) {
      w = tg_ld16(seq, y0 + (uint64_t)t);
      nd = 16;
    } else {
      nd = ncols - t < 16 ? ncols - t : 16;
      w = tg_ld16(seq, y0 - (uint64_t)t - (uint64_t)nd);
    }
  }
  TG_HD void refill(int t) {  // t % 16 == 0
    if (t == t_next) { word = word_next; need = need_next; }
    else load(t, word, need);
    t_next = -1;
  }
  TG_HD void prefetch(int t) {  // t % 16 == 0; only columns below ncols are ever fetched
    if (t < ncols) { load(t, word_next, need_next); t_next = t; }
  }
  TG_HD uint32_t at(int t) const {  // after refill(t & ~15)
    const int u = t & 15;
    const int nib = side == 0 ? u : need - 1 - u;
    return (uint32_t)(word >> (60 - 4 * nib)) & 15u;
  }
};

// Builds the profile of x from the packed read: side 0: x[p] = read[xoff + p]; side 1: x[p] = read[xlen - 1 - p].
// The read is fetched 16 symbols at a time (it was one global load per symbol).
TG_HD void tg_dpt_profile(const TgDptMem& m, const uint64_t* rp, uint32_t xoff, int xlen, int side) {
  for (int k = 0; k < 24; k++) m.msk[k * m.mstride] = 0;  // rows of the codes a y symbol can have ($ A C G N T); a read byte outside
                                                          // ACGNT (code 7) lands in a row that is never looked up
  for (int p0 = 0; p0 < xlen; p0 += 16) {
    const int cnt = xlen - p0 < 16 ? xlen - p0 : 16;
    // side 0: symbols xoff + p0 ..; side 1: the cnt symbols that END at read position xlen - p0, taken backwards
    const uint64_t w = tg_ld16(rp, side == 0 ? (uint64_t)xoff + (uint64_t)p0 : (uint64_t)(xlen - p0 - cnt));
    uint32_t* col = m.msk + (size_t)(p0 >> 5) * m.mstride;
    const uint32_t bit0 = 1u << (p0 & 31);
    for (int u = 0; u < cnt; u++) {
      const int nib = side == 0 ? u : cnt - 1 - u;
      const uint32_t code = (uint32_t)(w >> (60 - 4 * nib)) & 7u;  // PAD never occurs below xlen
      col[(size_t)(code * 4) * m.mstride] |= bit0 << u;
    }
  }
}
// Same from raw bytes (SwgExtend::extend on caller-supplied sequences; symbols are the ACGNT codes).
TG_HD void tg_dpt_profile_codes(const TgDptMem& m, const uint8_t* xcodes, int xlen) {
  for (int k = 0; k < 24; k++) m.msk[k * m.mstride] = 0;
  for (int p = 0; p < xlen; p++) m.msk[((xcodes[p] & 7u) * 4 + (p >> 5)) * m.mstride] |= 1u << (p & 31);
}

TG_HD uint32_t tg_dpt_funnel(uint32_t lo, uint32_t hi, uint32_t s) {  // (hi:lo) >> s, s in [0, 31]
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

// bits of the profile of symbol `sym` for rows row0 .. row0 + 32*NW - 1 (row i <-> x[i-1]; row 0 has no symbol)
template <int NW>
TG_HD void tg_dpt_window(const TgDptMem& m, uint32_t sym, int row0, uint32_t* w) {
  // bit position of row i is i - 1: shift the 128-bit mask right by row0 - 1 (left by one when row0 == 0)
  const int s = row0 - 1;
  const uint32_t* base = m.msk + (size_t)(sym * 4) * m.mstride;
  if (s < 0) {
    uint32_t prev = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
      const uint32_t cur = k < 4 ? base[k * m.mstride] : 0u;
      w[k] = (cur << 1) | (prev >> 31);
      prev = cur;
    }
    return;
  }
  const int a = s >> 5;
  const uint32_t sh = (uint32_t)s & 31u;
  uint32_t cur = a < 4 ? base[a * m.mstride] : 0u;
#pragma unroll
  for (int k = 0; k < NW; k++) {
    const uint32_t nxt = a + k + 1 < 4 ? base[(a + k + 1) * m.mstride] : 0u;
    w[k] = tg_dpt_funnel(cur, nxt, sh);
    cur = nxt;
  }
}

struct TgDptResult {
  int score, xend, yend;
  uint32_t cells;
};

// One DP cell (src/swg.rs:82-99 / :121-140 + triple_max :226-240).  hC/hDm2: same row, previous column; diag: D - 2 of the
// row above in the previous column; rr/dvm2: R and D - 2 of the row above in this column.  Updates the running column
// state and returns the new D - 2.
TG_HD int tg_dpt_cell(const TgDptMem& m, int hC, int hDm2, int diag, uint32_t match, int b, int& rr, int& dvm2, int& c_out,
                      uint32_t& tbits, int& key, int& ubm) {
  const int c = tg_dpt_max(hC - 1, hDm2);
  const int r_ = tg_dpt_max(rr - 1, dvm2);
  const int d = diag + (match ? 3 : 1);  // (as multiply-adds on the FMA pipe: tried, same speed -- the loop is not ALU-pipe bound)
  const int nd = tg_dpt_max3(d, c, r_);
  // direction: 0 when nd == d, else 1 when nd == c, else 2 (nd >= d and nd >= c, so the differences are >= 0)
  const int f1 = tg_dpt_min(nd - d, 1), f2 = tg_dpt_min(nd - c, 1);
  tbits += (uint32_t)(f1 + f1 * f2) << (2 * (b & 15));
  c_out = c;
  rr = r_;
  // the integer ALU pipe is the bottleneck of this loop: plain additions go to the FMA pipe as multiply-adds
  const int nm2 = tg_dpt_mad(nd, m.one, -2);
  dvm2 = nm2;
  key = tg_dpt_max(key, tg_dpt_mad(nd, m.k128, 127 - b));
  ubm = tg_dpt_max(ubm, nd - b);
  return nm2;
}

// Fill.  Returns through `res`; trace in m.tr.  ncols = min(ylen, xlen + bw) >= 1, xlen >= 1.
// Band rows of the extension: min(2bw, xlen) + 1 in [tg_dpt_min_rows(WB), WB].
template <int WB>
TG_HDN void tg_dpt_fill(const TgDptMem& m, TgDptY& ys, int xlen, int ncols, int bw, int x_drop, bool bound_stop,
                        TgDptResult& res) {
  constexpr int TW = (2 * WB + 31) / 32;  // trace words per column
  constexpr int NW = (WB + 31) / 32;      // profile words per column
  constexpr int LB = tg_dpt_min_rows(WB); // slots 0 .. LB-1 exist in every column whose band is not clipped by xlen
  int Dm2[WB + 1], C[WB + 1];             // previous column: D - 2 and C per slot (slot WB: permanent "out of band")
  const int two_bw = 2 * bw;
#pragma unroll
  for (int b = 0; b <= WB; b++) {         // column 0 (src/swg.rs:62-71)
    const bool in0 = b <= two_bw;
    Dm2[b] = in0 ? (b == 0 ? -2 : -(b + 1) - 2) : TG_DPT_MIN;
    C[b] = b == 0 ? 0 : TG_DPT_MIN;
  }
  int max_score = 0, max_i = 0, max_j = 0;
  uint32_t cells = 0;
  int j = 1;
  bool stopped = false;
  // ---- phase 1: columns 1 .. min(bw, ncols), rows 0 .. min(2bw, xlen), slot = row ----------------------------------
  const int p1_cols = bw < ncols ? bw : ncols;
  const int span1 = two_bw < xlen ? two_bw : xlen;  // >= LB - 1
  for (; j <= p1_cols; j++) {
    if (((j - 1) & 15) == 0) ys.refill(j - 1);
    const uint32_t yc = ys.at(j - 1);
    uint32_t w[NW];
    tg_dpt_window<NW>(m, yc, 0, w);
    uint32_t tb[TW];
#pragma unroll
    for (int k = 0; k < TW; k++) tb[k] = 0;
    // row 0: only the horizontal (deletion) branch exists (d = R = MIN), quirk Q1: C[0] starts at 0
    int diag = Dm2[0];
    const int c0 = tg_dpt_max(C[0] - 1, Dm2[0]);
    C[0] = c0; Dm2[0] = c0 - 2;
    tb[0] = 1u;
    int key = c0 * 128 + 127, ubm = c0;
    int rr = TG_DPT_MIN, dvm2 = c0 - 2;
#pragma unroll
    for (int b = 1; b < WB; b++) {
      if (b >= LB && b > span1) break;
      const int old = Dm2[b];
      Dm2[b] = tg_dpt_cell(m, C[b], old, diag, (w[b >> 5] >> (b & 31)) & 1u, b, rr, dvm2, C[b], tb[b >> 4], key, ubm);
      diag = old;
    }
#pragma unroll
    for (int k = 0; k < TW; k++) m.tr[((size_t)(j - 1) * TW + k) * m.tstride] = tb[k];
    cells += (uint32_t)span1 + 1u;
    const int cm = key >> 7;
    if (cm > max_score) { max_score = cm; max_i = 127 - (key & 127); max_j = j; }
    if (cm < max_score - x_drop || (bound_stop && ubm + xlen <= max_score)) { stopped = true; break; }
  }
  if (!stopped) {
    // ---- phase 2, band not clipped: columns bw+1 .. min(ncols, xlen - bw), rows j-bw .. j+bw, slot = row - (j - bw)
    const int full_cols = ncols < xlen - bw ? ncols : xlen - bw;
    for (; j <= full_cols; j++) {
      if (((j - 1) & 15) == 0) ys.refill(j - 1);
      const uint32_t yc = ys.at(j - 1);
      const int lo = j - bw;
      uint32_t w[NW];
      tg_dpt_window<NW>(m, yc, lo, w);
      uint32_t tb[TW];
#pragma unroll
      for (int k = 0; k < TW; k++) tb[k] = 0;
      int key = TG_DPT_MIN * 128, ubm = TG_DPT_MIN;
      int rr = TG_DPT_MIN, dvm2 = TG_DPT_MIN;
#pragma unroll
      for (int b = 0; b < WB; b++) {
        if (b >= LB && b > two_bw) break;
        Dm2[b] = tg_dpt_cell(m, C[b + 1], Dm2[b + 1], Dm2[b], (w[b >> 5] >> (b & 31)) & 1u, b, rr, dvm2, C[b], tb[b >> 4], key, ubm);
      }
#pragma unroll
      for (int k = 0; k < TW; k++) m.tr[((size_t)(j - 1) * TW + k) * m.tstride] = tb[k];
      cells += (uint32_t)two_bw + 1u;
      const int cm = key >> 7;
      if (cm > max_score) { max_score = cm; max_i = lo + 127 - (key & 127); max_j = j; }
      if (cm < max_score - x_drop || (bound_stop && ubm - lo + xlen <= max_score)) { stopped = true; break; }
    }
  }
  if (!stopped) {
    // ---- phase 2, band clipped by the last row: rows j-bw .. xlen (fewer every column) ----------------------------------
    for (; j <= ncols; j++) {
      if (((j - 1) & 15) == 0) ys.refill(j - 1);
      const uint32_t yc = ys.at(j - 1);
      const int lo = j - bw;
      const int span = xlen - lo;  // >= 0 because j <= xlen + bw
      uint32_t w[NW];
      tg_dpt_window<NW>(m, yc, lo, w);
      uint32_t tb[TW];
#pragma unroll
      for (int k = 0; k < TW; k++) tb[k] = 0;
      int key = TG_DPT_MIN * 128, ubm = TG_DPT_MIN;
      int rr = TG_DPT_MIN, dvm2 = TG_DPT_MIN;
#pragma unroll
      for (int b = 0; b < WB; b++) {
        if (b > span) break;
        Dm2[b] = tg_dpt_cell(m, C[b + 1], Dm2[b + 1], Dm2[b], (w[b >> 5] >> (b & 31)) & 1u, b, rr, dvm2, C[b], tb[b >> 4], key, ubm);
      }
#pragma unroll
      for (int k = 0; k < TW; k++) m.tr[((size_t)(j - 1) * TW + k) * m.tstride] = tb[k];
      cells += (uint32_t)span + 1u;
      const int cm = key >> 7;
      if (cm > max_score) { max_score = cm; max_i = lo + 127 - (key & 127); max_j = j; }
      if (cm < max_score - x_drop || (bound_stop && ubm - lo + xlen <= max_score)) break;
    }
  }
  res.score = max_score; res.xend = max_i; res.yend = max_j; res.cells = cells;
}

// Traceback (src/swg.rs:170-207).  emit(kind, run) is called in generation order (end cell -> origin), i.e. for
// rev(operations), with equal consecutive unit operations already merged.  Returns the number of emitted words.
template <int WB, class Emit>
TG_HDN uint32_t tg_dpt_traceback(const TgDptMem& m, TgDptY& ys, int xlen, int bw, const TgDptResult& res, Emit&& emit) {
  constexpr int TW = (2 * WB + 31) / 32;
  uint32_t n = 0;
  int i = res.xend, j = res.yend;
  if (i < xlen) { emit(n, (uint32_t)TG_OP_XCLIP, (uint32_t)(xlen - i)); n++; }
  uint32_t cur_kind = 0xFFu, cur_run = 0;
  int ybase = -1;
  // one-word-per-column classes: the walk moves at most one column per step, so four columns are fetched at once
  // (independent loads) instead of one dependent L2 round trip per step
  uint32_t tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
  int tw_base = 0x7fffffff;
  while (i > 0 || j > 0) {
    uint32_t dir;
    if (j == 0) dir = 2;  // column 0 is all Ins (src/swg.rs:65,70)
    else {
      const int slot = j <= bw ? i : i - (j - bw);
      uint32_t word;
      if (TW == 1) {
        const int cj = j - 1;
        if (cj < tw_base) {
          tw_base = cj >= 3 ? cj - 3 : 0;
          tw0 = m.tr[(size_t)tw_base * m.tstride];
          tw1 = tw_base + 1 <= cj ? m.tr[(size_t)(tw_base + 1) * m.tstride] : 0u;
          tw2 = tw_base + 2 <= cj ? m.tr[(size_t)(tw_base + 2) * m.tstride] : 0u;
          tw3 = tw_base + 3 <= cj ? m.tr[(size_t)(tw_base + 3) * m.tstride] : 0u;
        }
        const int k = cj - tw_base;
        word = k == 0 ? tw0 : k == 1 ? tw1 : k == 2 ? tw2 : tw3;
      } else {
        word = m.tr[((size_t)(j - 1) * TW + (slot >> 4)) * m.tstride];
      }
      dir = (word >> (2 * (slot & 15))) & 3u;
    }
    uint32_t kind;
    if (dir == 0) {
      if (((j - 1) & ~15) != ybase) { ybase = (j - 1) & ~15; ys.refill(ybase); }
      const uint32_t yc = ys.at(j - 1);
      const int p = i - 1;
      const bool eq = (m.msk[(yc * 4 + (p >> 5)) * m.mstride] >> (p & 31)) & 1u;
      kind = eq ? TG_OP_MATCH : TG_OP_SUBST;
      i--; j--;
    } else if (dir == 1) {
      kind = TG_OP_DEL;
      j--;
    } else {
      kind = TG_OP_INS;
      i--;
    }
    if (kind == cur_kind) cur_run++;
    else {
      if (cur_run) { emit(n, cur_kind, cur_run); n++; }
      cur_kind = kind; cur_run = 1;
    }
  }
  if (cur_run) { emit(n, cur_kind, cur_run); n++; }
  return n;
}
