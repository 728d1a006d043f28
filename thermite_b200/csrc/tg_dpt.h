// tg_dpt.h -- SwgExtend::extend / trace (reference src/swg.rs:31-207): ONE THREAD PER PAIR OF EXTENSIONS, 16-bit scores.
//
// The warp-cooperative fills of tg_core.h spend most of their issue slots on shuffles, pipeline fill/drain and idle
// lanes (a 91-bp read gives x <= 71 rows).  Here a thread owns two whole extensions of the same shape (same xlen and
// band width; the caller sorts tasks so that neighbours agree): extension A lives in the low 16 bits and extension B in
// the high 16 bits of every 32-bit register, and the max-plus recurrence runs on the packed-halfword integer
// instructions of sm_90+ (VIADDMNMX.S16x2, VIMNMX3.S16x2, VIMNMX.U16x2): one instruction, two cells.  The band of the
// current column lives in REGISTERS (template parameter WB = band slots, fully unrolled), there is no communication.
//
//   * scores are stored BIASED: value + TG_P2_BIAS (16384), "minus infinity" = TG_P2_MIN (4096).  Real scores of an
//     extension with x <= 128 symbols stay within +-1000 of the bias and the minus-infinity chains lose at most one per
//     cell, so every halfword stays in [0, 32767]: signed and unsigned comparisons agree, and plain 32-bit additions /
//     subtractions of packed words (issued as multiply-adds on the FMA pipe) never carry from one half into the other.
//   * slots: in the first bw columns (src/swg.rs:75-113, "phase 1", quirk Q2) slot b is row b; afterwards
//     (:116-154, "phase 2") slot b is row (j - bw) + b, i.e. the window slides down one row per column.  The update is
//     done IN PLACE in ascending slot order: new[b] needs old[b+1] (same row, previous column), old[b] (diagonal) and
//     new[b-1] (row above), so no register is ever moved.
//   * match scores come from a bit profile: for each symbol c a 128-bit mask of the x positions holding c; one funnel
//     shift per column and extension aligns the mask of y[j-1] with the slots, a byte permute interleaves the two
//     extensions' masks 16 slots at a time, a cell extracts its two bits with one shift and one mask.
//   * column maximum: a running packed max; the FIRST row that reaches it (src/swg.rs:101-104, strict '>' in row order)
//     is the LAST strict improvement of the running max, recorded as one bit per slot and read back only in columns that
//     raise the extension's maximum.  The x-drop test (:110-112) and the optional bound stop run once per
//     column and extension; an extension that has stopped keeps being computed (its half of the registers is ignored)
//     until its partner stops as well.
//   * trace: two bit planes per column, "not diagonal" and "Ins" (tie priority diag > Del > Ins, src/swg.rs:226-240),
//     16 slots of both extensions per word; written to a per-thread strided buffer; traceback walks both extensions in
//     one loop (two independent chains of dependent loads) and re-derives Match/Subst from the profile.
//
// Limits (the caller routes everything else to the warp kernels): xlen <= TG_DPT_MAX_X, rows in band
// min(2*bw, xlen) + 1 <= WB <= TG_DPT_MAX_WB, x_drop >= bw.
//
// Written as portable C++ (like tg_core.h) so that csrc/hosttest.cpp can run the very same code on the CPU.
#pragma once
#include "tg_core.h"

#define TG_DPT_MAX_X 128
#define TG_DPT_MAX_WB 80
#define TG_DPT_NCLS 12           // class 0: not eligible; class 1: WB = 4; class c >= 2: WB = 8 * (c - 1)
#define TG_P2_BIAS 16384
#define TG_P2_MIN 4096

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
// trace words per column of a PAIR: classes of at most 8 slots keep both bit planes in one word
TG_HD constexpr int tg_dpt_twp(int wb) { return wb <= 8 ? 1 : 2 * ((wb + 15) / 16); }

// Sort key of a task inside its class (< TG_DPT_CBINS): tasks with equal keys have the same (band width, xlen) unless
// one of the two was clamped, so neighbours in sorted order can share a thread.  Inside a class either the band width
// (band-limited: 2 bw <= xlen) or xlen (rows = xlen + 1) takes one of at most 8 values.
#define TG_DPT_CBINS 1024
TG_HD uint32_t tg_dpt_subkey(int xlen, int bw) {
  // (long extensions first: the kernels hand out work in sorted order, the short ones fill the tail of a launch)
  if (2 * bw <= xlen) return (uint32_t)((3 - (bw & 3)) * 128 + (xlen < 127 ? 127 - xlen : 0));
  return 512u + (uint32_t)((7 - (xlen & 7)) * 64 + (bw < 63 ? 63 - bw : 0));
}

// ---- packed pairs of 16-bit lanes ---------------------------------------------------------------------------------------
TG_HD uint32_t tg_p2_both(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
TG_HD int tg_p2_half(uint32_t p, int h) { return (int)((p >> (16 * h)) & 0xFFFFu); }
#ifndef __CUDA_ARCH__
TG_HD int tg_p2_s(uint32_t p, int h) { return (int)(int16_t)(uint16_t)(p >> (16 * h)); }
TG_HD uint32_t tg_p2_pack(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | (((uint32_t)hi & 0xFFFFu) << 16); }
#endif
TG_HD uint32_t tg_p2_addmax(uint32_t a, uint32_t b, uint32_t c) {  // per half: max(a + b, c), signed
#ifdef __CUDA_ARCH__
  return __viaddmax_s16x2(a, b, c);
#else
  int r[2];
  for (int h = 0; h < 2; h++) {
    const int s = (int)(int16_t)(uint16_t)(tg_p2_s(a, h) + tg_p2_s(b, h));
    r[h] = s > tg_p2_s(c, h) ? s : tg_p2_s(c, h);
  }
  return tg_p2_pack(r[0], r[1]);
#endif
}
TG_HD uint32_t tg_p2_max(uint32_t a, uint32_t b) {  // signed
#ifdef __CUDA_ARCH__
  return __vmaxs2(a, b);
#else
  int r[2];
  for (int h = 0; h < 2; h++) r[h] = tg_p2_s(a, h) > tg_p2_s(b, h) ? tg_p2_s(a, h) : tg_p2_s(b, h);
  return tg_p2_pack(r[0], r[1]);
#endif
}
TG_HD uint32_t tg_p2_max3(uint32_t a, uint32_t b, uint32_t c) {  // signed
#ifdef __CUDA_ARCH__
  return __vimax3_s16x2(a, b, c);
#else
  return tg_p2_max(tg_p2_max(a, b), c);
#endif
}
TG_HD uint32_t tg_p2_minu(uint32_t a, uint32_t b) {  // unsigned
#ifdef __CUDA_ARCH__
  return __vminu2(a, b);
#else
  int r[2];
  for (int h = 0; h < 2; h++) r[h] = tg_p2_half(a, h) < tg_p2_half(b, h) ? tg_p2_half(a, h) : tg_p2_half(b, h);
  return tg_p2_pack(r[0], r[1]);
#endif
}
TG_HD uint32_t tg_p2_min3u(uint32_t a, uint32_t b, uint32_t c) {  // unsigned
#ifdef __CUDA_ARCH__
  return __vimin3_u16x2(a, b, c);
#else
  return tg_p2_minu(tg_p2_minu(a, b), c);
#endif
}
// low / high halfwords of two words side by side: {a.lo, b.lo} and {a.hi, b.hi}
TG_HD uint32_t tg_p2_zip_lo(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, 0x5410);
#else
  return (a & 0xFFFFu) | (b << 16);
#endif
}
TG_HD uint32_t tg_p2_zip_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, 0x7632);
#else
  return (a >> 16) | (b & 0xFFFF0000u);
#endif
}

// a * b + c, forced onto the FMA pipe on the device (ptxas otherwise picks IADD3 / LEA / SHF on the busier ALU pipe).
// The multipliers come from constant memory (tg_dpt_k): a value ptxas cannot see cannot be strength-reduced.
TG_HD uint32_t tg_dpt_mad(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#else
  return a * b + c;
#endif
}
// [s] = 1 << s for s < 16, [16] = 2, [17] = -1
#define TG_DPT_K_INIT {1u, 2u, 4u, 8u, 16u, 32u, 64u, 128u, 256u, 512u, 1024u, 2048u, 4096u, 8192u, 16384u, 32768u, 2u, 0xFFFFFFFFu}
#ifdef __CUDA_ARCH__
static __constant__ uint32_t tg_dpt_k[18] = TG_DPT_K_INIT;
#else
static const uint32_t tg_dpt_k[18] = TG_DPT_K_INIT;
#endif
#define TG_DPT_K1 tg_dpt_k[0]
#define TG_DPT_K2 tg_dpt_k[16]
#define TG_DPT_KM1 tg_dpt_k[17]

struct TgDptMem {
  uint32_t* msk;     // match profiles: half h, word (sym * 4 + k) at msk[(h * 32 + sym * 4 + k) * mstride], sym in 0..7
  uint32_t mstride;
  uint32_t* tr;      // trace: word (col * TWP + k) at tr[(col * TWP + k) * tstride]
  uint32_t tstride;
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
  TG_HD void init(const uint64_t* s, uint64_t y, int nc, int sd) {
    seq = s; y0 = y; ncols = nc; side = sd; word = 0; need = 0; word_next = 0; need_next = 0; t_next = -1;
  }
  TG_HD void load(int t, uint64_t& w, int& nd) const {
    if (side == 0) {
      w = tg_ld16(seq, y0 + (uint64_t)t);
      nd = 16;
    } else {
      nd = ncols - t < 16 ? ncols - t : 16;
      w = tg_ld16(seq, y0 - (uint64_t)t - (uint64_t)nd);
    }
  }
  TG_HD void refill(int t) {  // t % 16 == 0, t < ncols
    if (t == t_next) { word = word_next; need = need_next; }
    else load(t, word, need);
    t_next = -1;
  }
  TG_HD void refill_ahead(int t) {  // refill(t) and fetch the following word (only columns below ncols are ever fetched)
    refill(t);
    if (t + 16 < ncols) { load(t + 16, word_next, need_next); t_next = t + 16; }
  }
  TG_HD uint32_t at(int t) const {  // after refill(t & ~15)
    const int u = t & 15;
    const int nib = side == 0 ? u : need - 1 - u;
    return (uint32_t)(word >> (60 - 4 * nib)) & 15u;
  }
};

// Builds the profile of x (half h) from the packed read: side 0: x[p] = read[xoff + p]; side 1: x[p] = read[xlen - 1 - p].
// 16 symbols per load; the eight groups of one 32-position word are collected in registers ([code][position] bytes of two
// 64-bit accumulators per 16 symbols would need dynamic register indexing, so the words are updated in shared memory, but
// the read itself is fetched once per 16 symbols instead of once per symbol).
TG_HD void tg_dpt_profile(const TgDptMem& m, int h, const uint64_t* rp, uint32_t xoff, int xlen, int side) {
  uint32_t* base = m.msk + (size_t)(h * 32) * m.mstride;
  for (int k = 0; k < 24; k++) base[k * m.mstride] = 0;  // rows of the codes a y symbol can have ($ A C G N T); a read byte outside ACGNT
                                                         // (code 7) lands in a row that is never looked up
  for (int p0 = 0; p0 < xlen; p0 += 16) {
    const int cnt = xlen - p0 < 16 ? xlen - p0 : 16;
    // side 0: symbols xoff + p0 ..; side 1: the cnt symbols that END at read position xlen - p0, taken backwards
    const uint64_t w = tg_ld16(rp, side == 0 ? (uint64_t)xoff + (uint64_t)p0 : (uint64_t)(xlen - p0 - cnt));
    uint32_t* col = base + (size_t)(p0 >> 5) * m.mstride;
    const uint32_t bit0 = 1u << (p0 & 31);
    for (int u = 0; u < cnt; u++) {
      const int nib = side == 0 ? u : cnt - 1 - u;
      const uint32_t code = (uint32_t)(w >> (60 - 4 * nib)) & 7u;  // PAD never occurs below xlen
      col[(size_t)(code * 4) * m.mstride] |= bit0 << u;
    }
  }
}
// Same from raw bytes (SwgExtend::extend on caller-supplied sequences; symbols are the ACGNT codes).
TG_HD void tg_dpt_profile_codes(const TgDptMem& m, int h, const uint8_t* xcodes, int xlen) {
  uint32_t* base = m.msk + (size_t)(h * 32) * m.mstride;
  for (int k = 0; k < 24; k++) base[k * m.mstride] = 0;
  for (int p = 0; p < xlen; p++) base[((xcodes[p] & 7u) * 4 + (p >> 5)) * m.mstride] |= 1u << (p & 31);
}

TG_HD uint32_t tg_dpt_funnel(uint32_t lo, uint32_t hi, uint32_t s) {  // (hi:lo) >> s, s in [0, 31]
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

// bits of half h's profile of symbol `sym` for rows row0 .. row0 + 32*NW - 1 (row i <-> x[i-1]; row 0 has no symbol)
template <int NW>
TG_HD void tg_dpt_window(const TgDptMem& m, int h, uint32_t sym, int row0, uint32_t* w) {
  // bit position of row i is i - 1: shift the 128-bit mask right by row0 - 1 (left by one when row0 == 0)
  const int s = row0 - 1;
  const uint32_t* base = m.msk + (size_t)(h * 32 + sym * 4) * m.mstride;
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

struct TgDpt2Result {  // [h]: extension A (low halves) and B (high halves)
  int score[2], xend[2], yend[2];
  uint32_t cells[2];
};

// One DP cell of both extensions (src/swg.rs:82-99 / :121-140 + triple_max :226-240).  hC / hS: C and D - 2 of the same
// row in the previous column; diag: D - 2 of the row above in the previous column; rr / dv: R and D - 2 of the row above
// in this column.  X: match bits of 16 slots of both extensions.  Returns the new D - 2.
// 10 ALU-pipe instructions (2 VIADDMNMX for C and R, SHF + LOP3 for the match bits, VIMNMX3 for D, VIMNMX + VIMNMX3 for
// the two trace bits, VIMNMX + VIMNMX for the running maximum and its improvement bit, VIADDMNMX for the bound) and
// 9 multiply-adds on the FMA pipe, for two cells.
TG_HD uint32_t tg_dpt2_cell(uint32_t hC, uint32_t hS, uint32_t diag, uint32_t X, int b, bool merge, uint32_t& rr, uint32_t& dv,
                            uint32_t& c_out, uint32_t& T1, uint32_t& T2, uint32_t& M, uint32_t& best, uint32_t& ubm) {
  const int s = b & 15;
  const uint32_t c = tg_p2_addmax(hC, 0xFFFFFFFFu, hS);
  const uint32_t r_ = tg_p2_addmax(rr, 0xFFFFFFFFu, dv);
  const uint32_t t = (X >> s) & 0x10001u;
  const uint32_t d = tg_dpt_mad(tg_dpt_mad(t, TG_DPT_K2, diag), TG_DPT_K1, 0x10001u);  // diag + 1 (mismatch) or + 3 (match)
  const uint32_t nd = tg_p2_max3(d, c, r_);
  // direction: nd >= d and nd >= c, so the differences are >= 0 in both halves
  const uint32_t u1 = tg_dpt_mad(d, TG_DPT_KM1, nd), u2 = tg_dpt_mad(c, TG_DPT_KM1, nd);
  const uint32_t f1 = tg_p2_minu(u1, 0x10001u);           // 1: not the diagonal
  const uint32_t f12 = tg_p2_min3u(u1, u2, 0x10001u);     // 1: neither the diagonal nor Del, i.e. Ins
  T1 = tg_dpt_mad(f1, tg_dpt_k[s], T1);
  if (merge) T1 = tg_dpt_mad(f12, tg_dpt_k[(s & 7) + 8], T1);
  else T2 = tg_dpt_mad(f12, tg_dpt_k[s], T2);
  const uint32_t nb = tg_p2_max(best, nd);
  const uint32_t g = tg_p2_minu(tg_dpt_mad(best, TG_DPT_KM1, nb), 0x10001u);  // 1: this cell raised the column maximum
  M = tg_dpt_mad(g, tg_dpt_k[s], M);
  best = nb;
  ubm = tg_p2_addmax(nd, tg_p2_both(-b), ubm);
  c_out = c;
  rr = r_;
  const uint32_t nS = tg_dpt_mad(nd, TG_DPT_K1, 0xFFFDFFFEu);  // nd - 2 in both halves
  dv = nS;
  return nS;
}

// One column of both extensions.  MODE 0: phase 1 (slot = row, row 0 has only the horizontal branch, quirk Q1);
// MODE 1: sliding band, not clipped by the last row (slots 0 .. limit, the first LB of them unconditionally);
// MODE 2: sliding band clipped by the last row (slots 0 .. limit).  S / C: D - 2 and C of the previous column per slot.
template <int WB, int MODE>
TG_HD void tg_dpt2_column(const TgDptMem& m, uint32_t* S, uint32_t* C, uint32_t ycA, uint32_t ycB, int j, int row0, int limit,
                          uint32_t& best, uint32_t& ubm, uint32_t* M) {
  constexpr int NX = (WB + 15) / 16;      // packed words of 16 slots
  constexpr int NW = (WB + 31) / 32;      // profile words per column
  constexpr int LB = tg_dpt_min_rows(WB); // slots 0 .. LB-1 exist in every column whose band is not clipped by xlen
  constexpr bool MERGE = WB <= 8;
  constexpr int TWP = tg_dpt_twp(WB);
  uint32_t wA[NW], wB[NW], X[NX], T1[NX], T2[NX];
  tg_dpt_window<NW>(m, 0, ycA, row0, wA);
  tg_dpt_window<NW>(m, 1, ycB, row0, wB);
#pragma unroll
  for (int k = 0; k < NX; k++) {
    X[k] = (k & 1) ? tg_p2_zip_hi(wA[k >> 1], wB[k >> 1]) : tg_p2_zip_lo(wA[k >> 1], wB[k >> 1]);
    T1[k] = 0; T2[k] = 0; M[k] = 0;
  }
  uint32_t rr = tg_p2_both(TG_P2_MIN), dv = tg_p2_both(TG_P2_MIN);
  if (MODE == 0) {
    uint32_t diag = S[0];
    const uint32_t c0 = tg_p2_addmax(C[0], 0xFFFFFFFFu, S[0]);
    C[0] = c0; S[0] = c0 - 0x20002u;
    T1[0] = 0x10001u;  // Del
    M[0] = 0x10001u;
    best = c0; ubm = c0;
    dv = S[0];
#pragma unroll
    for (int b = 1; b < WB; b++) {
      if (b >= LB && b > limit) break;
      const uint32_t old = S[b];
      S[b] = tg_dpt2_cell(C[b], old, diag, X[b >> 4], b, MERGE, rr, dv, C[b], T1[b >> 4], T2[b >> 4], M[b >> 4], best, ubm);
      diag = old;
    }
  } else {
    best = tg_p2_both(TG_P2_MIN); ubm = tg_p2_both(TG_P2_MIN);
#pragma unroll
    for (int b = 0; b < WB; b++) {
      if ((MODE == 2 || b >= LB) && b > limit) break;
      S[b] = tg_dpt2_cell(C[b + 1], S[b + 1], S[b], X[b >> 4], b, MERGE, rr, dv, C[b], T1[b >> 4], T2[b >> 4], M[b >> 4], best, ubm);
    }
  }
  uint32_t* tcol = m.tr + (size_t)(j - 1) * TWP * m.tstride;
  if (MERGE) tcol[0] = T1[0];
  else {
#pragma unroll
    for (int k = 0; k < NX; k++) { tcol[(2 * k) * m.tstride] = T1[k]; tcol[(2 * k + 1) * m.tstride] = T2[k]; }
  }
}

TG_HD int tg_dpt_clz(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
// highest slot whose improvement bit is set in half h
template <int NX>
TG_HD int tg_dpt2_last_slot(const uint32_t* M, int h) {
  int slot = 0;
#pragma unroll
  for (int k = 0; k < NX; k++) {
    const uint32_t v = (M[k] >> (16 * h)) & 0xFFFFu;
    if (v) slot = 16 * k + 31 - tg_dpt_clz(v);
  }
  return slot;
}

// Fill of a pair.  Both extensions have xlen >= 1 symbols and band width bw; half h has ncols[h] = min(ylen, xlen + bw) >= 1
// columns and x-drop threshold x_drop[h] >= bw.  Returns through `res`; trace in m.tr.
// Band rows of the extensions: min(2bw, xlen) + 1 in [tg_dpt_min_rows(WB), WB].
template <int WB>
TG_HDN void tg_dpt2_fill(const TgDptMem& m, TgDptY& ysA, TgDptY& ysB, int xlen, int bw, int ncolsA, int ncolsB, int x_dropA,
                         int x_dropB, bool bound_stop, TgDpt2Result& res) {
  constexpr int NX = (WB + 15) / 16;
  uint32_t S[WB + 1], C[WB + 1];          // previous column: D - 2 and C per slot (slot WB: permanent "out of band")
  const int two_bw = 2 * bw;
#pragma unroll
  for (int b = 0; b <= WB; b++) {         // column 0 (src/swg.rs:62-71)
    const bool in0 = b <= two_bw;
    S[b] = tg_p2_both(in0 ? TG_P2_BIAS + (b == 0 ? -2 : -(b + 1) - 2) : TG_P2_MIN);
    C[b] = tg_p2_both(b == 0 ? TG_P2_BIAS : TG_P2_MIN);
  }
  int max_score[2] = {0, 0}, max_i[2] = {0, 0}, max_j[2] = {0, 0};
  const int ncols[2] = {ncolsA, ncolsB}, x_drop[2] = {x_dropA, x_dropB};
  uint32_t cells[2] = {0, 0};
  bool stop[2] = {false, false};
  const int ncols_max = ncolsA > ncolsB ? ncolsA : ncolsB;
  uint32_t best, ubm, M[NX];
  int j = 1;
  // after a column: running maxima, x-drop (src/swg.rs:110-112 / :151-153), bound stop, last column of a half
#define TG_DPT2_AFTER(lo_, n_cells_)                                                                                    \
  _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                                        \
    if (!stop[h]) {                                                                                                      \
      cells[h] += (uint32_t)(n_cells_);                                                                                  \
      const int cm = tg_p2_half(best, h) - TG_P2_BIAS;                                                                   \
      if (cm > max_score[h]) { max_score[h] = cm; max_i[h] = (lo_) + tg_dpt2_last_slot<NX>(M, h); max_j[h] = j; }        \
      if (cm < max_score[h] - x_drop[h] || j >= ncols[h] ||                                                              \
          (bound_stop && tg_p2_half(ubm, h) - TG_P2_BIAS - (lo_) + xlen <= max_score[h]))                                \
        stop[h] = true;                                                                                                  \
    }                                                                                                                    \
  }
#define TG_DPT2_Y()                                                                                                      \
  if (((j - 1) & 15) == 0) {                                                                                             \
    if (j - 1 < ncolsA) ysA.refill_ahead(j - 1);                                                                         \
    if (j - 1 < ncolsB) ysB.refill_ahead(j - 1);                                                                         \
  }                                                                                                                      \
  const uint32_t ycA = ysA.at(j - 1) & 7u, ycB = ysB.at(j - 1) & 7u;
  // ---- phase 1: columns 1 .. min(bw, ncols), rows 0 .. min(2bw, xlen), slot = row ----------------------------------
  const int p1_cols = bw < ncols_max ? bw : ncols_max;
  const int span1 = two_bw < xlen ? two_bw : xlen;  // >= LB - 1
  for (; j <= p1_cols; j++) {
    TG_DPT2_Y()
    tg_dpt2_column<WB, 0>(m, S, C, ycA, ycB, j, 0, span1, best, ubm, M);
    TG_DPT2_AFTER(0, span1 + 1)
    if (stop[0] && stop[1]) break;
  }
  if (!(stop[0] && stop[1])) {
    // ---- phase 2, band not clipped: columns bw+1 .. min(ncols, xlen - bw), rows j-bw .. j+bw, slot = row - (j - bw)
    const int full_cols = ncols_max < xlen - bw ? ncols_max : xlen - bw;
    for (; j <= full_cols; j++) {
      TG_DPT2_Y()
      const int lo = j - bw;
      tg_dpt2_column<WB, 1>(m, S, C, ycA, ycB, j, lo, two_bw, best, ubm, M);
      TG_DPT2_AFTER(lo, two_bw + 1)
      if (stop[0] && stop[1]) break;
    }
  }
  if (!(stop[0] && stop[1])) {
    // ---- phase 2, band clipped by the last row: rows j-bw .. xlen (fewer every column) ----------------------------------
    for (; j <= ncols_max; j++) {
      TG_DPT2_Y()
      const int lo = j - bw;
      const int span = xlen - lo;  // >= 0 because j <= xlen + bw
      tg_dpt2_column<WB, 2>(m, S, C, ycA, ycB, j, lo, span, best, ubm, M);
      TG_DPT2_AFTER(lo, span + 1)
      if (stop[0] && stop[1]) break;
    }
  }
#undef TG_DPT2_AFTER
#undef TG_DPT2_Y
#pragma unroll
  for (int h = 0; h < 2; h++) { res.score[h] = max_score[h]; res.xend[h] = max_i[h]; res.yend[h] = max_j[h]; res.cells[h] = cells[h]; }
}

// ---- gapless shortcut for the traceback ------------------------------------------------------------------------------
// If the maximum lies on the main diagonal (xend == yend = n) and the diagonal itself scores it (matches - mismatches of
// x[0..n) against y[0..n) == score), the reference's traceback IS the diagonal: with P(i) the score of the diagonal
// prefix, optimality of the whole diagonal gives D(i,i) = P(i) for every i (a better path to (i,i) plus the diagonal
// suffix would beat the maximum), hence d(i,i) = D(i-1,i-1) + s = D(i,i), and triple_max prefers the diagonal on ties
// (src/swg.rs:226-240).  The operations are then the runs of the symbol-equality mask, computed 16 symbols at a time
// from the packed sequences: no trace word is read.  Most extensions of real reads are gapless.
TG_HD int tg_dpt_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
TG_HD uint32_t tg_dpt_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popc(v);
#else
  return (uint32_t)__builtin_popcount(v);
#endif
}
TG_HD uint32_t tg_dpt_brev(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
// flags at bits 0, 4, ..., 28 -> bits 0 .. 7
TG_HD uint32_t tg_dpt_compress8(uint32_t t) {
  t = (t | (t >> 3)) & 0x03030303u;
  t = (t | (t >> 6)) & 0x000F000Fu;
  return (t | (t >> 12)) & 0xFFu;
}
// 16 symbols of an extension-order sequence, block p0 .. p0 + 15.  side 0: symbol u in nibble 15 - u (top nibble first);
// side 1 (the sequence runs downwards from position `end`, exclusive): symbol u in nibble u.
TG_HD uint64_t tg_dpt_block16(const uint64_t* seq, uint64_t start_or_end, int p0, int side) {
  if (side == 0) return tg_ld16(seq, start_or_end + (uint64_t)p0);
  const long long s = (long long)start_or_end - 16 - p0;
  if (s >= 0) return tg_ld16(seq, (uint64_t)s);
  return tg_ld16(seq, 0) >> (4 * (int)(-s));  // the sequence starts less than 16 symbols below: -s <= 15
}
// Equality mask of x[0..n) and y[0..n) (bit p of hi:lo) when the diagonal shortcut applies; false otherwise.  n <= 128.
TG_HDN bool tg_dpt_diag_mask(const uint64_t* xseq, uint32_t xoff, int xlen, int side, const uint64_t* yseq, uint64_t y0, int xend,
                             int yend, int score, uint64_t& lo, uint64_t& hi) {
  if (xend != yend) return false;
  const int n = xend;
  lo = 0; hi = 0;
  int matches = 0;
#pragma unroll
  for (int b = 0; b < TG_DPT_MAX_X / 16; b++) {
    const int p0 = 16 * b;
    if (p0 < n) {
      const int cnt = n - p0 < 16 ? n - p0 : 16;
      const uint64_t xw = tg_dpt_block16(xseq, side == 0 ? (uint64_t)xoff : (uint64_t)xlen, p0, side);
      const uint64_t yw = tg_dpt_block16(yseq, y0, p0, side);
      const uint64_t z = xw ^ yw;
      const uint64_t eqf = ~(z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x1111111111111111ULL;
      uint32_t e = tg_dpt_compress8((uint32_t)eqf) | (tg_dpt_compress8((uint32_t)(eqf >> 32)) << 8);  // bit k: nibble k from the bottom
      if (side == 0) e = tg_dpt_brev(e) >> 16;
      e &= 0xFFFFu >> (16 - cnt);
      matches += (int)tg_dpt_popc(e);
      if (b < 4) lo |= (uint64_t)e << (16 * (b & 3));
      else hi |= (uint64_t)e << (16 * (b & 3));
    }
  }
  return 2 * matches - n == score;
}
// rev(operations) of the diagonal alignment with equality mask hi:lo: emit(h, idx, kind, run) as the traceback does.
template <class Emit>
TG_HD uint32_t tg_dpt_diag_emit(uint64_t lo, uint64_t hi, int n, int xlen, int h, Emit&& emit) {
  uint32_t idx = 0;
  if (n < xlen) { emit(h, idx, (uint32_t)TG_OP_XCLIP, (uint32_t)(xlen - n)); idx++; }
  if (n == 0) return idx;
  // left-align: symbol n - 1 in bit 127
  {
    const int sh = 128 - n;  // 0 .. 127
    if (sh >= 64) { hi = lo << (sh - 64); lo = 0; }
    else if (sh > 0) { hi = (hi << sh) | (lo >> (64 - sh)); lo <<= sh; }
  }
  int rem = n;
  while (rem > 0) {
    const bool v = (hi >> 63) & 1u;
    const uint64_t xh = v ? ~hi : hi, xl = v ? ~lo : lo;
    int len = xh ? tg_dpt_clz64(xh) : 64 + tg_dpt_clz64(xl);
    if (len > rem) len = rem;
    emit(h, idx, (uint32_t)(v ? TG_OP_MATCH : TG_OP_SUBST), (uint32_t)len);
    idx++;
    if (len >= 64) { hi = len >= 128 ? 0 : lo << (len - 64); lo = 0; }
    else { hi = (hi << len) | (lo >> (64 - len)); lo <<= len; }
    rem -= len;
  }
  return idx;
}

// Traceback of both extensions (src/swg.rs:170-207) in one loop.  emit(h, idx, kind, run) is called in generation order
// (end cell -> origin), i.e. for rev(operations), with equal consecutive unit operations already merged.
// live[h]: walk extension h; n[h] receives the number of emitted words.
template <int WB, class Emit>
TG_HDN void tg_dpt2_traceback(const TgDptMem& m, TgDptY& ysA, TgDptY& ysB, int xlen, int bw, const TgDpt2Result& res, bool liveA,
                              bool liveB, uint32_t& nA, uint32_t& nB, Emit&& emit) {
  constexpr bool MERGE = WB <= 8;
  constexpr int TWP = tg_dpt_twp(WB);
  int wi[2] = {res.xend[0], res.xend[1]}, wj[2] = {res.yend[0], res.yend[1]}, ybase[2] = {-1, -1};
  uint32_t n[2] = {0, 0}, cur_kind[2] = {0xFFu, 0xFFu}, cur_run[2] = {0, 0};
  bool live[2] = {liveA, liveB};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (live[h] && wi[h] < xlen) { emit(h, 0u, (uint32_t)TG_OP_XCLIP, (uint32_t)(xlen - wi[h])); n[h] = 1; }
    live[h] = live[h] && (wi[h] > 0 || wj[h] > 0);
  }
  while (live[0] || live[1]) {
    uint32_t t1[2] = {0, 0};
    const uint32_t* tw[2] = {m.tr, m.tr};
    int slot[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {  // both loads are issued before either is used
      if (live[h] && wj[h] > 0) {
        slot[h] = wj[h] <= bw ? wi[h] : wi[h] - (wj[h] - bw);
        tw[h] = m.tr + ((size_t)(wj[h] - 1) * TWP + (MERGE ? 0 : 2 * (slot[h] >> 4))) * m.tstride;
        t1[h] = *tw[h];
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (!live[h]) continue;
      TgDptY& ys = h == 0 ? ysA : ysB;
      uint32_t dir;
      if (wj[h] == 0) dir = 2;  // column 0 is all Ins (src/swg.rs:65,70)
      else {
        const int sh = (MERGE ? slot[h] : (slot[h] & 15)) + 16 * h;
        if (((t1[h] >> sh) & 1u) == 0) dir = 0;
        else if (MERGE) dir = 1 + ((t1[h] >> (sh + 8)) & 1u);
        else dir = 1 + ((tw[h][m.tstride] >> sh) & 1u);
      }
      uint32_t kind;
      if (dir == 0) {
        if (((wj[h] - 1) & ~15) != ybase[h]) { ybase[h] = (wj[h] - 1) & ~15; ys.refill(ybase[h]); }
        const uint32_t yc = ys.at(wj[h] - 1) & 7u;
        const int p = wi[h] - 1;
        const bool eq = (m.msk[(size_t)(h * 32 + yc * 4 + (p >> 5)) * m.mstride] >> (p & 31)) & 1u;
        kind = eq ? TG_OP_MATCH : TG_OP_SUBST;
        wi[h]--; wj[h]--;
      } else if (dir == 1) {
        kind = TG_OP_DEL;
        wj[h]--;
      } else {
        kind = TG_OP_INS;
        wi[h]--;
      }
      if (kind == cur_kind[h]) cur_run[h]++;
      else {
        if (cur_run[h]) { emit(h, n[h], cur_kind[h], cur_run[h]); n[h]++; }
        cur_kind[h] = kind; cur_run[h] = 1;
      }
      live[h] = wi[h] > 0 || wj[h] > 0;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++)
    if (cur_run[h]) { emit(h, n[h], cur_kind[h], cur_run[h]); n[h]++; }
  nA = n[0]; nB = n[1];
}
