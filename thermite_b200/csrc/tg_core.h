// tg_core.h -- the alignment hot path as warp-cooperative code, written once and compiled twice:
//   * by nvcc for sm_100a inside the kernels of thermite_gpu.cu (W = DevWarp: 32 lanes, shuffles, votes);
//   * by g++ inside csrc/hosttest.cpp (W = HostWarp: 1 lane, or 32 emulated lanes on threads) so the
//     logic can be unit-tested in a container without a GPU.  The host build is TEST-ONLY: the shipped
//     library never executes it (no CPU fallback; see thermite_gpu.cu).
//
// Reference being replaced (paths under /root/reference): src/index.rs:228-255 (all_smems),
// src/aligner.rs:123-449, src/swg.rs:31-240, src/txome.rs:77-160.
#pragma once
#include "tg_internal.h"

#ifdef __CUDACC__
#define TG_HD __host__ __device__ __forceinline__
#define TG_HDN __host__ __device__
#else
#define TG_HD inline
#define TG_HDN
#endif

#ifdef __CUDACC__
#define TG_NOINLINE __noinline__
#else
#define TG_NOINLINE
#endif

#ifdef __CUDA_ARCH__
#define TG_LDG(p) __ldg(p)
#define TG_CLZ64(x) __clzll((long long)(x))
#define TG_CTZ64(x) (__ffsll((long long)(x)) - 1)
#else
#define TG_LDG(p) (*(p))
#define TG_CLZ64(x) __builtin_clzll(x)
#define TG_CTZ64(x) __builtin_ctzll(x)
#endif

#define TG_MINV (-(1 << 29))
#define TG_DIRECT 0x80000000u

// ------------------------------------------------------------------------------------------------
// packed 4-bit sequences: 16 symbols per u64, first symbol in the top nibble.  Every packed array has
// at least 2 readable words past its last symbol.
// ------------------------------------------------------------------------------------------------
TG_HD uint64_t tg_ld16(const uint64_t* s, uint64_t pos) {
  uint64_t w = pos >> 4;
  uint32_t sh = (uint32_t)(pos & 15) * 4;
  uint64_t a = TG_LDG(s + w);
  if (sh == 0) return a;
  uint64_t b = TG_LDG(s + w + 1);
  return (a << sh) | (b >> (64 - sh));
}
TG_HD uint64_t tg_ld16_local(const uint64_t* s, uint32_t pos) {  // shared / host memory
  uint32_t w = pos >> 4, sh = (pos & 15) * 4;
  uint64_t a = s[w];
  if (sh == 0) return a;
  return (a << sh) | (s[w + 1] >> (64 - sh));
}
TG_HD uint32_t tg_code_at(const uint64_t* s, uint64_t pos) {
  return (uint32_t)(TG_LDG(s + (pos >> 4)) >> ((15 - (uint32_t)(pos & 15)) * 4)) & 15u;
}

// read byte -> symbol code after to_ascii_uppercase (src/aligner.rs:125)
TG_HD uint32_t tg_ascii_code(uint8_t c) {
  if (c >= 'a' && c <= 'z') c -= 32;
  switch (c) {
    case 'A': return TG_C_A;
    case 'C': return TG_C_C;
    case 'G': return TG_C_G;
    case 'N': return TG_C_N;
    case 'T': return TG_C_T;
    default: return TG_C_OTHER;
  }
}

// ------------------------------------------------------------------------------------------------
// k-mer hashing (table build and probe must agree)
// ------------------------------------------------------------------------------------------------
TG_HD uint64_t tg_mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
TG_HD uint64_t tg_top_nibbles(uint32_t n) { return n >= 16 ? ~0ull : ~(~0ull >> (4 * n)); }
TG_HD uint64_t tg_hash_kmer(uint64_t w0, uint64_t w1) { return tg_mix64(w0 ^ tg_mix64(w1 + 0x9e3779b97f4a7c15ULL)); }
// true when some symbol of the top-n nibbles has its low three bits all set (codes 7 and 15)
TG_HD bool tg_has_bad_symbol(uint64_t w) { return (w & (w >> 1) & (w >> 2) & 0x1111111111111111ULL) != 0; }
// true when some nibble of w is zero (the '$' separator)
TG_HD bool tg_has_zero_nibble(uint64_t w) {
  uint64_t t = w | (w >> 1);
  t |= t >> 2;
  return (~t & 0x1111111111111111ULL) != 0;
}
TG_HD uint32_t tg_tag_of(uint64_t h) {
  uint32_t t = (uint32_t)(h >> 32);
  return t ? t : 1u;
}

// Longest common prefix of read[q..L) (packed, padded with 0xF) and seq[pos..).  *read_le is set when the
// read suffix sorts before the text suffix or is a prefix of it.
TG_HD uint32_t tg_lcp(const uint64_t* rp, uint32_t q, uint32_t L, const uint64_t* seq, uint64_t pos, bool* read_le) {
  uint32_t maxl = L - q, off = 0;
  while (off < maxl) {
    uint64_t a = tg_ld16_local(rp, q + off), b = tg_ld16(seq, pos + off);
    uint64_t x = a ^ b;
    if (x) {
      uint32_t nb = (uint32_t)TG_CLZ64(x) >> 2;
      uint32_t l = off + nb;
      if (l >= maxl) { *read_le = true; return maxl; }
      uint32_t sh = (15 - nb) * 4;
      *read_le = ((a >> sh) & 15) < ((b >> sh) & 15);
      return l;
    }
    off += 16;
  }
  *read_le = true;
  return maxl;
}

// ------------------------------------------------------------------------------------------------
// Seeding (replaces Index::all_smems, src/index.rs:228-255).
//
// E(q) = q + (longest prefix of read[q..] that occurs in the text).  The SMEMs with length >= k are the
// intervals [q, E(q)) with E(q)-q >= k and (q == 0 or E(q-1) < E(q)); the occurrences of an SMEM are a
// contiguous range of suffix-array rows inside the range of its leading k-mer.
// ------------------------------------------------------------------------------------------------
struct TgSeedHit {
  uint32_t e;    // E(q), 0 when read[q..q+k) does not occur
  uint32_t lo;   // first SA row of the occurrences (or the text position when direct)
  uint32_t cnt;  // number of occurrences | TG_DIRECT
};

// lcp of read[q..L) and seq[pos..) given that the first `known` symbols are equal
TG_HD uint32_t tg_lcp_from(const uint64_t* rp, uint32_t q, uint32_t L, const uint64_t* seq, uint64_t pos, uint32_t known,
                           bool* read_le) {
  uint32_t maxl = L - q, off = known & ~15u;  // restart at the packed word that holds symbol `known`
  while (off < maxl) {
    uint64_t a = tg_ld16_local(rp, q + off), b = tg_ld16(seq, pos + off);
    uint64_t x = a ^ b;
    if (x) {
      uint32_t nb = (uint32_t)TG_CLZ64(x) >> 2;
      uint32_t l = off + nb;
      if (l >= maxl) { *read_le = true; return maxl; }
      uint32_t sh = (15 - nb) * 4;
      *read_le = ((a >> sh) & 15) < ((b >> sh) & 15);
      return l;
    }
    off += 16;
  }
  *read_le = true;
  return maxl;
}

// DEFER: return false (out untouched) instead of searching when the k-mer occurs more than once -- the device runs
// those probes in a second kernel so that a warp holds either cheap single-occurrence probes or searches, not a mix.
template <bool DEFER = false>
TG_HD bool tg_seed_offset(const uint64_t* rp, uint32_t L, uint32_t q, uint32_t k, const TgSlot* slots,
                          uint64_t slot_mask, const uint64_t* text4, const uint32_t* sa, TgSeedHit& out) {
  if (!DEFER) { out.e = 0; out.lo = 0; out.cnt = 0; }
  TgSeedHit none{0, 0, 0};
  uint64_t w0 = tg_ld16_local(rp, q) & tg_top_nibbles(k);
  uint64_t w1 = k > 16 ? (tg_ld16_local(rp, q + 16) & tg_top_nibbles(k - 16)) : 0;
  if (tg_has_bad_symbol(w0) || tg_has_bad_symbol(w1)) { out = none; return true; }
  uint64_t h = tg_hash_kmer(w0, w1);
  uint32_t tag = tg_tag_of(h);
  uint64_t idx = h & slot_mask;
  for (;;) {
#ifdef __CUDA_ARCH__
    uint4 raw = __ldg((const uint4*)(slots + idx));
    TgSlot s{raw.x, raw.y, raw.z, raw.w};
#else
    TgSlot s = slots[idx];
#endif
    if (s.tag == 0) { out = none; return true; }
    if (s.tag == tag) {
      bool le;
      if (s.count == 1) {
        uint32_t l = tg_lcp(rp, q, L, text4, s.lo, &le);
        if (l >= k) { out.e = q + l; out.lo = s.lo; out.cnt = 1u | TG_DIRECT; return true; }
      } else {
        if (DEFER) return false;
        // All rows of the group share their first k symbols; one row tells whether they are the read's k-mer (a tag
        // collision is a different k-mer).  Then: first row whose suffix is >= the read suffix, by binary search that
        // only compares beyond the prefix both current bounds are known to share with the read.
        const uint32_t lo = s.lo, hi = s.lo + s.count;
        uint32_t a = lo, b = hi;
        uint32_t la = tg_lcp(rp, q, L, text4, TG_LDG(sa + lo), &le);  // lcp with row lo
        if (la >= k) {
          uint32_t lb = k;  // lcp with the (virtual) bound b: at least the k-mer
          uint32_t l_ins = 0, l_prev = 0;  // lcp of the row at / before the insertion point
          if (le) { b = lo; l_ins = la; }  // the read sorts before (or equals a prefix of) the first row
          else {
            a = lo + 1; l_prev = la;
            while (a < b) {
              const uint32_t mid = a + ((b - a) >> 1);
              const uint32_t known = la < lb ? la : lb;
              const uint32_t l = tg_lcp_from(rp, q, L, text4, TG_LDG(sa + mid), known, &le);
              if (le) { b = mid; lb = l; l_ins = l; } else { a = mid + 1; la = l; l_prev = l; }
            }
          }
          const uint32_t ins = a;
          if (ins >= hi) l_ins = 0;
          const uint32_t best = l_ins > l_prev ? l_ins : l_prev;
          // rows with lcp >= best form one contiguous range around ins; it is usually a single row, so gallop outwards
          uint32_t left = ins, right = ins;  // [left, right)
          if (l_prev >= best && ins > lo) {  // lcp is non-decreasing on [lo, ins)
            left = ins - 1;
            uint32_t step = 1;
            uint32_t known_ok = left;  // row known to have lcp >= best
            for (;;) {  // find a row with lcp < best (or run out), doubling the step
              if (known_ok == lo) { left = lo; break; }
              const uint32_t probe = known_ok - lo > step ? known_ok - step : lo;
              if (tg_lcp_from(rp, q, L, text4, TG_LDG(sa + probe), k, &le) >= best) { known_ok = probe; step <<= 1; continue; }
              // first row in (probe, known_ok] with lcp >= best
              uint32_t x = probe + 1, y = known_ok;
              while (x < y) {
                const uint32_t mid = x + ((y - x) >> 1);
                if (tg_lcp_from(rp, q, L, text4, TG_LDG(sa + mid), k, &le) >= best) y = mid; else x = mid + 1;
              }
              left = x;
              break;
            }
          }
          if (l_ins >= best && ins < hi) {   // lcp is non-increasing on [ins, hi)
            uint32_t step = 1;
            uint32_t known_ok = ins;
            for (;;) {
              if (known_ok == hi - 1) { right = hi; break; }
              const uint32_t probe = hi - 1 - known_ok > step ? known_ok + step : hi - 1;
              if (tg_lcp_from(rp, q, L, text4, TG_LDG(sa + probe), k, &le) >= best) { known_ok = probe; step <<= 1; continue; }
              // last row in [known_ok, probe) with lcp >= best
              uint32_t x = known_ok, y = probe - 1;
              while (x < y) {
                const uint32_t mid = x + ((y - x + 1) >> 1);
                if (tg_lcp_from(rp, q, L, text4, TG_LDG(sa + mid), k, &le) >= best) x = mid; else y = mid - 1;
              }
              right = x + 1;
              break;
            }
          }
          out.e = q + best; out.lo = left; out.cnt = right - left;
          return true;
        }
      }
    }
    idx = (idx + 1) & slot_mask;
  }
}

// Serial part of seeding: pick the SMEM starts from E[], order them as Index::all_smems does and write the
// records.  hits[q] valid for q + k <= L.
// Order: bio emits, for i0 = 0, max-end, ...: the SMEMs covering i0 by DESCENDING start; then
// src/index.rs:251-253 stable-sorts by len ascending and reverses => len DESC, ties in REVERSE emission order.
// `sampled`: the table was filled by the probe waves (TG_PROBE_STRIDE): offsets inside a closed bracket hold nothing and
// are stepped over -- their E equals the bracket's, so none of them starts an SMEM.
#define TG_PROBE_STRIDE 8u
TG_HD uint32_t tg_probe_step(const TgSeedHit* row, uint32_t q, uint32_t q_last);
TG_HD uint32_t tg_smem_count(const TgSeedHit* hits, uint32_t L, uint32_t k, bool sampled = false) {  // number of SMEMs
  if (L < k || k == 0) return 0;
  uint32_t n = 0, prev_e = 0;
  for (uint32_t q = 0; q + k <= L;) {
    const uint32_t e = hits[q].e;
    if (e != 0 && (q == 0 || prev_e < e)) n++;
    prev_e = e;
    if (e == L) break;  // E is non-decreasing: every later offset ends at L too and starts no SMEM
    q = sampled ? tg_probe_step(hits, q, L - k) : q + 1;
  }
  return n;
}
// out: room for tg_smem_count() records; the `pad` field is used as scratch (emission group) and left 0
TG_HD uint32_t tg_smem_select(const TgSeedHit* hits, uint32_t L, uint32_t k, tg_seed* out, bool sampled = false) {
  if (L < k || k == 0) return 0;
  uint32_t n = 0, prev_e = 0;
  for (uint32_t q = 0; q + k <= L;) {
    const TgSeedHit h = hits[q];
    const uint32_t e = h.e;
    if (e != 0 && (q == 0 || prev_e < e)) {
      out[n].query_idx = q; out[n].len = e - q; out[n].sa_lo = h.lo;
      out[n].count = h.cnt & ~TG_DIRECT; out[n].direct = (h.cnt & TG_DIRECT) ? 1u : 0u; out[n].pad = 0;
      n++;
    }
    prev_e = e;
    if (e == L) break;
    q = sampled ? tg_probe_step(hits, q, L - k) : q + 1;
  }
  // emission groups
  uint32_t i0 = 0, idx = 0, g = 0;
  while (idx < n) {
    if (out[idx].query_idx > i0) i0 = out[idx].query_idx;
    uint32_t maxend = i0 + 1;
    while (idx < n && out[idx].query_idx <= i0) {
      out[idx].pad = g;
      uint32_t e = out[idx].query_idx + out[idx].len;
      if (e > maxend) maxend = e;
      idx++;
    }
    i0 = maxend;
    g++;
  }
  // insertion sort by (len DESC, group DESC, start ASC)
  for (uint32_t i = 1; i < n; i++) {
    const tg_seed cur = out[i];
    uint32_t j = i;
    while (j > 0) {
      const tg_seed p = out[j - 1];
      bool p_before = p.len > cur.len || (p.len == cur.len && (p.pad > cur.pad || (p.pad == cur.pad && p.query_idx < cur.query_idx)));
      if (p_before) break;
      out[j] = p;
      j--;
    }
    out[j] = cur;
  }
  for (uint32_t i = 0; i < n; i++) out[i].pad = 0;
  return n;
}

// ------------------------------------------------------------------------------------------------
// RLE op buffers
// ------------------------------------------------------------------------------------------------
struct TgOps {
  uint32_t* w;
  uint32_t n;
};
TG_HD void tg_ops_push(TgOps& o, uint32_t kind, uint32_t run) {
  if (run == 0 && kind <= TG_OP_INS) return;
  if (kind <= TG_OP_INS && o.n > 0 && (o.w[o.n - 1] & 7u) == kind) o.w[o.n - 1] += run << 3;
  else o.w[o.n++] = kind | (run << 3);
}
TG_HD void tg_ops_append_reversed(TgOps& dst, const TgOps& src) {
  for (uint32_t i = src.n; i-- > 0;) tg_ops_push(dst, src.w[i] & 7u, src.w[i] >> 3);
}
TG_HD void tg_ops_reverse(TgOps& o) {
  for (uint32_t i = 0, j = o.n; i + 1 < j; i++) {
    j--;
    uint32_t t = o.w[i]; o.w[i] = o.w[j]; o.w[j] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Banded Smith-Waterman-Gotoh extension (replaces SwgExtend::extend / trace, src/swg.rs:31-207) as an
// anti-diagonal wavefront: lane l owns rows [l*R, l*R+R) of the DP matrix and is one column behind lane l-1.
//
// Cell (i, j) (row i over x, column j over y) is in the reference's band iff
//     lo(j) <= i <= hi(j),  lo(j) = max(0, j - bw),  hi(j) = min(xlen, max(2*bw, j + bw))
// (rows 0..2bw for the first bw columns: quirk Q2), out-of-band neighbours count as MIN_SCORE, row-0 cells only
// take the "deletion" branch whose column-0 seed is C[0] = 0 (quirk Q1).  Requires x_drop >= bw (then the
// phase-1 x-drop break of src/swg.rs:110-112 is unreachable and no stale state exists: quirk Q4).
// Scores are exact in 32-bit; MIN_SCORE is represented by TG_MINV and never wins a max against a real cell.
// ------------------------------------------------------------------------------------------------
#define TG_CM_MIN (-(1 << 20))  // "no in-band cell yet" for the per-column running maximum (fits the 22-bit packing)

struct TgSwgResult {
  int score, xend, yend;
};

// trace layout: bytes; row (j-1) holds LANES * TB bytes, lane l's TB = (2R+7)/8 bytes at offset l*TB.
template <int R>
struct TgTraceBytes { static constexpr int value = (2 * R + 7) / 8; };

TG_HD int tg_max(int a, int b) { return a > b ? a : b; }

// W: warp policy (lane, LANES, shfl_up, shfl, any, sync).  xs: x symbols (xlen), ys: y symbols (ncols).
// ncols = min(ylen, xlen + bw) (later columns are empty and only trigger the x-drop break).
//
// Lane state per owned row r (row i = lane*R + r), all for the previous column:
//   Dm2[r] = D - 2 (what both the vertical-gap open of the row below and the horizontal-gap open of the next
//   column consume), C[r].  Out-of-band cells hold TG_MINV-ish values: "below the band" rows are never written
//   (they keep their MIN initialisation), "above the band" rows only need to hand MIN to the row below, so the
//   per-cell band test feeds just two selects.  C of dead rows drifts by -1 per column, far from wrapping.
template <int R, class W>
TG_HDN TG_NOINLINE void tg_swg_fill(W& w, const uint8_t* xs, const uint8_t* ys, int xlen, int ncols, int bw, int x_drop,
                                    uint8_t* trace, TgSwgResult& res, unsigned long long& cells, bool bound_stop) {
  constexpr int TB = TgTraceBytes<R>::value;
  constexpr int NW = (R + 15) / 16;
  const int lane = w.lane();
  const int nl = (xlen + R) / R;  // lanes that own at least one existing row
  const int i0 = lane * R;
  int Dm2[R], C[R];
  uint8_t xc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = i0 + r;
    const bool in0 = i <= 2 * bw;  // column 0 initialises rows 0..w-1 (src/swg.rs:62-71)
    Dm2[r] = in0 ? (i == 0 ? -2 : -(i + 1) - 2) : TG_MINV;
    C[r] = (in0 && i == 0) ? 0 : TG_MINV;
    xc[r] = (i >= 1 && i <= xlen) ? xs[i - 1] : (uint8_t)0xFE;
  }
  // D(i0-1, 0): the diagonal input of this lane's first row at column 1
  int diag_in;
  {
    const int ip = i0 - 1;
    diag_in = (ip < 0 || ip > 2 * bw) ? TG_MINV : (ip == 0 ? 0 : -(ip + 1));
  }
  int sendDm2 = TG_MINV, sendR = TG_MINV, sendCM = TG_CM_MIN * 1024, sendUB = TG_CM_MIN;
  const int rem0 = xlen - i0;  // x symbols left below row i0: a cell (i, j) can gain at most xlen - i more
  int max_score = 0, max_i = 0, max_j = 0;
  unsigned long long ccount = 0;
  bool stop = false;
  const bool is_last = lane == nl - 1;
  const int nsteps = ncols + nl - 1;
  const int two_bw = 2 * bw;
  for (int t = 0; t < nsteps; t++) {
    int upDm2 = w.shfl_up(sendDm2, 1);
    int upR = w.shfl_up(sendR, 1);
    int pcm = w.shfl_up(sendCM, 1);
    int ub = w.shfl_up(sendUB, 1);
    if (lane == 0) { upDm2 = TG_MINV; upR = TG_MINV; pcm = TG_CM_MIN * 1024; ub = TG_CM_MIN; }
    const int j = t - lane + 1;
    if (lane < nl && j >= 1 && j <= ncols) {
      const int next_diag = upDm2 + 2;  // D(i0-1, j) is the diagonal input of column j+1
      const uint8_t y = ys[j - 1];
      const int lo = tg_max(j - bw, 0);
      int hi = tg_max(j + bw, two_bw);
      hi = hi < xlen ? hi : xlen;
      const unsigned span = (unsigned)(hi - lo);
      const int rel0 = i0 - lo;
      ccount += span + 1u;
      int cm = pcm >> 10, cr = pcm & 1023;
      int dg = diag_in;
      uint32_t bits[NW];
#pragma unroll
      for (int b = 0; b < NW; b++) bits[b] = 0;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int c = tg_max(C[r] - 1, Dm2[r]);
        const int rr = tg_max(upR - 1, upDm2);
        const int d = dg + (xc[r] == y ? 1 : -1);
        const int nd = tg_max(tg_max(d, c), rr);
        const uint32_t gapdir = (nd != c) ? (2u << (2 * (r & 15))) : (1u << (2 * (r & 15)));
        if (nd != d) bits[r >> 4] |= gapdir;
        const bool inb = (unsigned)(rel0 + r) <= span;
        const int dsel = inb ? nd : TG_MINV;
        dg = Dm2[r] + 2;
        Dm2[r] = dsel - 2;
        C[r] = c;
        upDm2 = dsel - 2;
        upR = inb ? rr : TG_MINV;
        if (dsel > cm) cr = i0 + r;
        cm = tg_max(cm, dsel);
        ub = tg_max(ub, dsel + (rem0 - r));
      }
      diag_in = next_diag;
      sendDm2 = upDm2; sendR = upR; sendCM = cm * 1024 + cr; sendUB = ub;
      uint8_t* tp = trace + ((size_t)(j - 1) * W::LANES + lane) * TB;
#pragma unroll
      for (int b = 0; b < TB; b++) tp[b] = (uint8_t)(bits[b >> 2] >> (8 * (b & 3)));
      // column bookkeeping (src/swg.rs:101-112 / :142-153): only the last lane sees the complete column, the other
      // lanes run the same instructions on partial maxima and their result is ignored
      if (cm > max_score) { max_i = cr; max_j = j; }
      max_score = tg_max(max_score, cm);
      // x-drop of the reference, or (optionally) a proof that no later cell can STRICTLY exceed the running maximum:
      // every path into a later column crosses this one at some (i, j) and gains at most +1 per remaining x symbol,
      // so max_i(D(i,j) + xlen - i) <= max_score leaves score, end cell and traceback unchanged.
      stop = is_last && ((cm < max_score - x_drop) || (bound_stop && ub <= max_score));
    }
    if (w.any(stop)) break;
  }
  res.score = w.shfl(max_score, nl - 1);
  res.xend = w.shfl(max_i, nl - 1);
  res.yend = w.shfl(max_j, nl - 1);
  if (is_last) cells += ccount;  // cells the reference visits: rows lo(j)..hi(j) of every processed column
  w.sync();
}

// Narrow bands (2*bw + 1 <= 31 cells per column; every hit after a good first one): column-synchronous variant.
// Lane l tracks the single in-band row i with i % 32 == l, all lanes work on the SAME column, and the vertical gap
// state R(i) = max_{i' < i}(H(i') - 2 - (i - 1 - i')) with H = max(diag, horizontal) -- an exact rewrite of
// R(i) = max(R(i-1) - 1, D(i-1) - 2) because D = max(H, R) and extending beats re-opening -- is an exclusive prefix
// max over the band rows, done with log2(band) shuffles.  One column costs ~60 instructions instead of ~110 per
// wavefront step and there is no pipeline fill/drain.  Trace: byte [(j-1)*32 + (i & 31)], 2 bits used.
// W adds: reduce_max_i32(int), reduce_min_u32(u32).
template <class W>
TG_HDN TG_NOINLINE void tg_swg_fill_narrow(W& w, const uint8_t* xs, const uint8_t* ys, int xlen, int ncols, int bw, int x_drop,
                                           uint8_t* trace, TgSwgResult& res, unsigned long long& cells, bool bound_stop) {
  const int lane = w.lane();
  const int two_bw = 2 * bw;
  int myrow = lane;  // column 0: rows 0..2bw live in lanes 0..2bw (src/swg.rs:62-71)
  int Dp = lane <= two_bw ? (lane == 0 ? 0 : -(lane + 1)) : TG_MINV;
  int Cp = lane == 0 ? 0 : TG_MINV;
  uint8_t xc = (lane >= 1 && lane <= xlen) ? xs[lane - 1] : (uint8_t)0xFE;
  int max_score = 0, max_i = 0, max_j = 0;
  unsigned long long ccount = 0;
  for (int j = 1; j <= ncols; j++) {
    const int lo = tg_max(j - bw, 0);
    int hi = tg_max(j + bw, two_bw);
    hi = hi < xlen ? hi : xlen;
    const int rl = (lane - lo) & 31;  // position of this lane's row inside the band
    const int i = lo + rl;
    // D(i-1, j-1) lives in the previous lane; read it BEFORE that lane may recycle itself for a new row (the row
    // above the band's top row was in the band one column ago and is the top row's diagonal input)
    const int diag = w.shfl(Dp, (lane + 31) & 31);
    if (i != myrow) {  // the old row left the band at the top; the new one enters from below with MIN state
      myrow = i;
      Dp = TG_MINV; Cp = TG_MINV;
      xc = (i >= 1 && i <= xlen) ? xs[i - 1] : (uint8_t)0xFE;
    }
    const bool inb = i <= hi;
    const uint8_t y = ys[j - 1];
    const int c = tg_max(Cp - 1, Dp - 2);
    const int d = (i == 0 ? TG_MINV : diag) + (xc == y ? 1 : -1);
    const int h = inb ? tg_max(d, c) : TG_MINV;
    // exclusive prefix max of A(i') = H(i') + i' over the band rows above this one
    int v = h + i;
    const int span = hi - lo;
    for (int dd = 1; dd <= span; dd <<= 1) {
      const int t = w.shfl(v, (lane - dd) & 31);
      if (rl >= dd) v = tg_max(v, t);
    }
    const int ex = w.shfl(v, (lane + 31) & 31);
    const int rr = rl >= 1 ? ex - i - 1 : TG_MINV;
    const int nd = inb ? tg_max(h, rr) : TG_MINV;
    const uint32_t dir = (nd == d) ? 0u : ((nd == c) ? 1u : 2u);
    Dp = nd; Cp = c;
    trace[(size_t)(j - 1) * 32 + (i & 31)] = (uint8_t)dir;
    // column bookkeeping (uniform): band max, first row attaining it, upper bound on anything later
    const int cm = w.reduce_max_i32(nd);
    ccount += (unsigned)(span + 1);
    if (cm > max_score) {
      max_i = (int)w.reduce_min_u32((inb && nd == cm) ? (uint32_t)i : 0x7fffffffu);
      max_j = j;
      max_score = cm;
    }
    bool stop = cm < max_score - x_drop;
    if (bound_stop) {
      const int ub = w.reduce_max_i32(inb ? nd + (xlen - i) : TG_MINV);
      stop = stop || ub <= max_score;
    }
    if (stop) break;
  }
  res.score = max_score; res.xend = max_i; res.yend = max_j;
  if (lane == 0) cells += ccount;
  w.sync();
}

// Traceback (src/swg.rs:170-207) in generation order (end cell -> origin), i.e. rev(operations).
// Uniform across lanes; only lane 0 of W writes.
template <int R, class W>
TG_HDN TG_NOINLINE void tg_swg_traceback(W& w, const uint8_t* xs, const uint8_t* ys, int xlen, const uint8_t* trace,
                             const TgSwgResult& res, TgOps& out) {
  constexpr int TB = TgTraceBytes<R == 0 ? 1 : R>::value;
  if (w.lane() == 0) {
    int i = res.xend, j = res.yend;
    if (i < xlen) tg_ops_push(out, TG_OP_XCLIP, (uint32_t)(xlen - i));
    while (i > 0 || j > 0) {
      uint32_t dir;
      if (j == 0) dir = 2;  // column 0 is all Ins (src/swg.rs:65,70)
      else if constexpr (R == 0) {
        dir = trace[(size_t)(j - 1) * 32 + (i & 31)] & 3u;
      } else {
        constexpr int RR = R == 0 ? 1 : R;
        const int rr = i % RR;
        const uint8_t* tp = trace + ((size_t)(j - 1) * W::LANES + (i / RR)) * TB;
        dir = ((uint32_t)tp[rr >> 2] >> (2 * (rr & 3))) & 3u;
      }
      if (dir == 0) {
        tg_ops_push(out, xs[i - 1] == ys[j - 1] ? TG_OP_MATCH : TG_OP_SUBST, 1);
        i--; j--;
      } else if (dir == 1) {
        tg_ops_push(out, TG_OP_DEL, 1);
        j--;
      } else {
        tg_ops_push(out, TG_OP_INS, 1);
        i--;
      }
    }
  }
  out.n = (uint32_t)w.shfl((int)out.n, 0);
  w.sync();
}

// Full extension: early return for empty x or y (src/swg.rs:39-55), fill, traceback.
// `out` receives rev(operations).  R is chosen by the caller so that 32*R > xlen (or R*LANES > xlen on host).
template <int R, class W>
TG_HDN void tg_swg_extend_r(W& w, const uint8_t* xs, const uint8_t* ys, int xlen, int ylen, int bw, int x_drop,
                            uint8_t* trace, TgSwgResult& res, TgOps& out, unsigned long long& cells,
                            unsigned long long& n_ext, bool bound_stop) {
  if (xlen == 0 || ylen == 0) {
    res.score = 0; res.xend = 0; res.yend = 0;
    if (xlen > 0 && w.lane() == 0) tg_ops_push(out, TG_OP_XCLIP, (uint32_t)xlen);
    out.n = (uint32_t)w.shfl((int)out.n, 0);
    w.sync();
    return;
  }
  int ncols = ylen < xlen + bw ? ylen : xlen + bw;
  if constexpr (R == 0) tg_swg_fill_narrow<W>(w, xs, ys, xlen, ncols, bw, x_drop, trace, res, cells, bound_stop);
  else tg_swg_fill<R, W>(w, xs, ys, xlen, ncols, bw, x_drop, trace, res, cells, bound_stop);
  if (w.lane() == 0) n_ext++;
  tg_swg_traceback<R, W>(w, xs, ys, xlen, trace, res, out);
}

// rows-per-lane classes compiled for the device
TG_HD int tg_swg_rows_class(int xlen, int lanes) {
  int need = (xlen + lanes) / lanes;  // ceil((xlen+1)/lanes)
  if (need <= 1) return 1;
  if (need <= 2) return 2;
  if (need <= 3) return 3;
  if (need <= 4) return 4;
  if (need <= 6) return 6;
  if (need <= 8) return 8;
  if (need <= 12) return 12;
  return 16;
}

// RMAX bounds the rows-per-lane classes that get instantiated (register pressure of the kernel is set by the
// largest one); callers guarantee tg_swg_rows_class(xlen) <= RMAX.
template <class W, int RMAX = 16>
TG_HDN void tg_swg_extend(W& w, const uint8_t* xs, const uint8_t* ys, int xlen, int ylen, int bw, int x_drop,
                          uint8_t* trace, TgSwgResult& res, TgOps& out, unsigned long long& cells,
                          unsigned long long& n_ext, bool bound_stop = false) {
  if constexpr (W::LANES == 1) {
    tg_swg_extend_r<TG_MAX_READ_LEN + 1, W>(w, xs, ys, xlen, ylen, bw, x_drop, trace, res, out, cells, n_ext, bound_stop);
    return;
  } else {
    if (W::LANES == 32 && 2 * bw + 1 <= 31) {
      tg_swg_extend_r<0, W>(w, xs, ys, xlen, ylen, bw, x_drop, trace, res, out, cells, n_ext, bound_stop);
      return;
    }
    const int cls = tg_swg_rows_class(xlen, W::LANES);
#define TG_SWG_CASE(RR)                                                                                   \
  if constexpr (RMAX >= RR) {                                                                             \
    if (cls == RR) { tg_swg_extend_r<RR, W>(w, xs, ys, xlen, ylen, bw, x_drop, trace, res, out, cells, n_ext, bound_stop); return; } \
  }
    TG_SWG_CASE(1) TG_SWG_CASE(2) TG_SWG_CASE(3) TG_SWG_CASE(4) TG_SWG_CASE(6) TG_SWG_CASE(8) TG_SWG_CASE(12) TG_SWG_CASE(16)
#undef TG_SWG_CASE
  }
}
// bytes of trace needed per column for a given longest x
TG_HD int tg_trace_bytes_per_col(int max_xlen, int lanes) {
  if (lanes == 1) return (2 * (TG_MAX_READ_LEN + 1) + 7) / 8;
  int R = tg_swg_rows_class(max_xlen, lanes);
  return lanes * ((2 * R + 7) / 8);
}

#define TG_TREE_STACK 64  // (scratch kept in the warp layout; the tree walk itself was replaced by the stab lists below)

// ------------------------------------------------------------------------------------------------
// Warp-cooperative interval stabbing on the start-sorted lists (replaces the pointer-chasing tree walk on the
// device): a 32-ary search bounds the candidates, the lanes test them in parallel, and results are handed out
// in ascending find() rank, which is exactly the order rust-bio's IntervalTree::find yields them.
// W adds: ballot(bool) -> u32 lane mask, reduce_min_u32(u32).
// ------------------------------------------------------------------------------------------------
TG_HD TgStab tg_stab_load(const TgStab* a, uint32_t i) {
#ifdef __CUDA_ARCH__
  const uint4 v = __ldg((const uint4*)(a + i));
  return TgStab{v.x, v.y, v.z, v.w};
#else
  return a[i];
#endif
}
TG_HD uint32_t tg_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popc(v);
#else
  return (uint32_t)__builtin_popcount(v);
#endif
}
// first index whose start >= key
template <class W>
TG_HDN uint32_t tg_stab_lower_bound(W& w, const TgStab* a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  if constexpr (W::LANES == 1) {
    while (lo < hi) {
      uint32_t mid = lo + ((hi - lo) >> 1);
      if (a[mid].start < key) lo = mid + 1; else hi = mid;
    }
    return lo;
  } else {
    const uint32_t lane = (uint32_t)w.lane();
    while (hi - lo > (uint32_t)W::LANES) {
      const uint32_t step = (hi - lo + W::LANES - 1) / W::LANES;
      const uint32_t idx = lo + lane * step;
      const bool less = idx < hi && tg_stab_load(a, idx).start < key;
      const uint32_t c = tg_popc(w.ballot(less));  // starts are sorted: the probes that are < key form a prefix
      const uint32_t nhi = lo + c * step;
      if (c) lo = lo + (c - 1) * step + 1;
      if (nhi < hi) hi = nhi;
    }
    const uint32_t idx = lo + lane;
    const bool less = idx < hi && tg_stab_load(a, idx).start < key;
    return lo + tg_popc(w.ballot(less));
  }
}
struct TgStabRange {
  uint32_t lo, hi, qs, qe;
};
template <class W>
TG_HDN TgStabRange tg_stab_begin(W& w, const TgStab* a, uint32_t n, uint32_t maxlen, uint32_t qs, uint32_t qe) {
  TgStabRange r;
  r.qs = qs; r.qe = qe;
  // an interval reaching past qs starts after qs - maxlen; one starting at or after qe cannot intersect
  r.lo = tg_stab_lower_bound(w, a, n, qs >= maxlen ? qs - maxlen + 1 : 0u);
  r.hi = tg_stab_lower_bound(w, a, n, qe);
  return r;
}
// the intersecting interval with the smallest rank >= min_rank; false when there is none
template <class W>
TG_HDN bool tg_stab_next(W& w, const TgStab* a, const TgStabRange& r, uint32_t min_rank, uint32_t& rank, uint32_t& data) {
  uint32_t best = 0xFFFFFFFFu, bdata = 0;
  for (uint32_t i = r.lo + (uint32_t)w.lane(); i < r.hi; i += W::LANES) {
    TgStab e = tg_stab_load(a, i);
    if (e.start < r.qe && r.qs < e.end && e.rank >= min_rank && e.rank < best) { best = e.rank; bdata = e.data; }
  }
  const uint32_t m = w.reduce_min_u32(best);
  if (m == 0xFFFFFFFFu) return false;
  const uint32_t owners = w.ballot(best == m);
  int src = 0;
  while (!((owners >> src) & 1u)) src++;
  data = (uint32_t)w.shfl((int)bdata, src);
  rank = m;
  return true;
}

// Index::idx_to_ref (src/index.rs:287-290): first ref whose end_idx > idx
TG_HD uint32_t tg_idx_to_ref(const TgRef* refs, uint32_t n_refs, uint32_t idx) {
  uint32_t lo = 0, hi = n_refs;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (TG_LDG(&refs[mid].end_idx) <= idx) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// per-read state shared by the lanes of one warp ("shared memory" on the device)
// ------------------------------------------------------------------------------------------------
struct TgWarpMem {
  uint8_t* rd;       // read symbols [L]
  uint8_t* xs;       // staged x (left extension: reversed read prefix)
  uint8_t* ys;       // staged y [max_cols]
  uint8_t* trace;    // [max_cols * bytes_per_col]; also scratch for the final filters
  uint32_t* opsA;    // RLE buffers, ops_cap words each
  uint32_t* opsB;
  uint32_t* opsC;
  uint32_t* opsT;
  int32_t* stack;    // TG_TREE_STACK
  uint32_t ops_cap;
  uint64_t* rp;      // packed read (L/16 + 3 words, padded with 0xF)
  bool bound_stop;   // stop an extension once no later cell can beat the running maximum (same records, fewer cells)
};

struct TgAlignParams {
  TgIndexDev ix;
  tg_opts opts;
};

// One side-by-side extension result (bio Alignment without the ops)
struct TgAln {
  int32_t score;
  uint32_t ystart, yend, xstart, xend;
};

#ifdef TG_PROFILE_PHASES
#define TG_NPHASE 12
TG_HD long long tg_clock() {
#ifdef __CUDA_ARCH__
  return clock64();
#else
  return 0;
#endif
}
#define TG_TDECL() long long tg_t0_ = tg_clock()
#define TG_T0() tg_t0_ = tg_clock()
#define TG_T(ctr, k) do { long long tg_t1_ = tg_clock(); (ctr).ph[k] += (unsigned long long)(tg_t1_ - tg_t0_); tg_t0_ = tg_t1_; } while (0)
#else
#define TG_TDECL() do {} while (0)
#define TG_T0() do {} while (0)
#define TG_T(ctr, k) do {} while (0)
#endif
struct TgCounters {
  unsigned long long cells, n_ext, hits;
#ifdef TG_PROFILE_PHASES
  unsigned long long ph[TG_NPHASE];
#endif
};

// exact-match run lengths between the packed read and a packed sequence (extend_seed_match, src/aligner.rs:410-426)
TG_HD uint32_t tg_match_fwd(const uint64_t* rp, uint32_t q, uint32_t L, const uint64_t* seq, uint64_t pos, uint64_t seq_end) {
  uint64_t room = seq_end - pos;
  uint32_t maxn = L - q;
  if ((uint64_t)maxn > room) maxn = (uint32_t)room;
  uint32_t n = 0;
  while (n < maxn) {
    uint64_t x = tg_ld16_local(rp, q + n) ^ tg_ld16(seq, pos + n);
    if (x) { n += (uint32_t)TG_CLZ64(x) >> 2; break; }
    n += 16;
  }
  return n < maxn ? n : maxn;
}
TG_HD uint32_t tg_match_bwd(const uint64_t* rp, uint32_t q, const uint64_t* seq, uint64_t pos, uint64_t seq_start) {
  uint64_t room = pos - seq_start;
  uint32_t maxn = q;
  if ((uint64_t)maxn > room) maxn = (uint32_t)room;
  uint32_t n = 0;
  while (n < maxn) {
    uint32_t m = maxn - n < 16 ? maxn - n : 16;
    uint32_t sh = 4 * (16 - m);
    uint64_t a = tg_ld16_local(rp, q - n - m) >> sh, b = tg_ld16(seq, pos - n - m) >> sh;
    uint64_t x = a ^ b;
    if (x) { n += (uint32_t)TG_CTZ64(x) >> 2; break; }
    n += m;
  }
  return n < maxn ? n : maxn;
}

// An extend_left_right problem: reference sequence seq[lo_abs, hi_abs) with the seed at r_abs.
struct TgProblem {
  const uint64_t* seq;
  uint64_t lo_abs, hi_abs, r_abs;
  uint32_t q, len;
};
// The y windows extend_left_right hands to SwgExtend (only the first xlen+bw columns can be touched)
TG_HD void tg_problem_windows(const TgProblem& p, uint32_t L, uint32_t bw, uint32_t& ncR, uint32_t& ncL) {
  uint32_t xr = L - (p.q + p.len);
  uint64_t yr = p.hi_abs - (p.r_abs + p.len);
  ncR = xr == 0 ? 0u : (uint32_t)(yr < (uint64_t)xr + bw ? yr : (uint64_t)xr + bw);
  uint64_t span = (uint64_t)L + bw;
  uint64_t ys0 = (p.r_abs - p.lo_abs > span) ? p.r_abs - span : p.lo_abs;
  uint64_t yl = p.r_abs - ys0;
  ncL = p.q == 0 ? 0u : (uint32_t)(yl < (uint64_t)p.q + bw ? yl : (uint64_t)p.q + bw);
}
template <class W>
TG_HDN bool tg_same_symbols(W& w, const uint64_t* sa, uint64_t pa, const uint64_t* sb, uint64_t pb, uint32_t n) {
  bool diff = false;
  for (uint32_t c = (uint32_t)w.lane() * 16; c < n; c += W::LANES * 16) {
    uint64_t a = tg_ld16(sa, pa + c), b = tg_ld16(sb, pb + c);
    uint32_t m = n - c;
    if (m < 16) { uint64_t k = tg_top_nibbles(m); a &= k; b &= k; }
    diff |= a != b;
  }
  return !w.any(diff);
}
// Two problems are the same DP (same x parts, same y symbols in every column that can be visited): then
// extend_left_right returns the same score, read span, operations and the same offsets relative to the seed.
template <class W>
TG_HDN bool tg_same_problem(W& w, const TgProblem& a, const TgProblem& b, uint32_t L, uint32_t bw) {
  if (a.q != b.q || a.len != b.len) return false;
  uint32_t ar, al, br, bl;
  tg_problem_windows(a, L, bw, ar, al);
  tg_problem_windows(b, L, bw, br, bl);
  if (ar != br || al != bl) return false;
  if (ar && !tg_same_symbols(w, a.seq, a.r_abs + a.len, b.seq, b.r_abs + b.len, ar)) return false;
  if (al && !tg_same_symbols(w, a.seq, a.r_abs - al, b.seq, b.r_abs - bl, al)) return false;
  return true;
}

// extend_left_right (src/aligner.rs:352-407) on packed sequence `seq`: the reference sequence is
// seq[lo_abs, hi_abs), the seed sits at absolute position r_abs.  Result coordinates are absolute in `seq`.
// `out` receives the stitched operations: rev(left.ops) ++ Match*len ++ right.ops.
template <class W, int RMAX = 16>
TG_HDN void tg_extend_left_right(W& w, TgWarpMem& m, const uint64_t* seq, uint64_t lo_abs, uint64_t hi_abs,
                                 uint64_t r_abs, uint32_t q, uint32_t len, uint32_t L, uint32_t bw, int32_t x_drop,
                                 TgAln& aln, TgOps& out, TgCounters& ctr) {
  const int lane = w.lane();
  TgSwgResult rr, rl;
  TG_TDECL();
  // ---- right: x = read[q+len..], y = seq[r+len .. hi) (src/aligner.rs:360-362)
  TgOps tmp{m.opsT, 0};
  {
    int xlen = (int)(L - (q + len));
    uint64_t y0 = r_abs + len;
    uint64_t ylen64 = hi_abs - y0;
    int ylen = ylen64 > (uint64_t)(xlen + (int)bw) ? xlen + (int)bw + 1 : (int)ylen64;  // only the first xlen+bw columns matter
    int ncols = ylen < xlen + (int)bw ? ylen : xlen + (int)bw;
    if (xlen > 0)
      for (int t = lane; t < ncols; t += W::LANES) m.ys[t] = (uint8_t)tg_code_at(seq, y0 + (uint64_t)t);
    w.sync();
    TG_T(ctr, 0);
    tg_swg_extend<W, RMAX>(w, m.rd + q + len, m.ys, xlen, ylen, (int)bw, x_drop, m.trace, rr, tmp, ctr.cells, ctr.n_ext, m.bound_stop);
    TG_T(ctr, 1);
  }
  // ---- left: x = rev(read[..q]), y = rev(seq[max(r-(L+bw), lo) .. r)) (src/aligner.rs:364-375)
  out.n = 0;
  {
    int xlen = (int)q;
    uint64_t span = (uint64_t)L + bw;
    uint64_t ys0 = (r_abs - lo_abs > span) ? r_abs - span : lo_abs;
    uint64_t ylen64 = r_abs - ys0;
    int ylen = ylen64 > (uint64_t)(xlen + (int)bw) ? xlen + (int)bw + 1 : (int)ylen64;
    int ncols = ylen < xlen + (int)bw ? ylen : xlen + (int)bw;
    for (int t = lane; t < xlen; t += W::LANES) m.xs[t] = m.rd[q - 1 - t];
    if (xlen > 0)
      for (int t = lane; t < ncols; t += W::LANES) m.ys[t] = (uint8_t)tg_code_at(seq, r_abs - 1 - (uint64_t)t);
    w.sync();
    TG_T(ctr, 0);
    tg_swg_extend<W, RMAX>(w, m.xs, m.ys, xlen, ylen, (int)bw, x_drop, m.trace, rl, out, ctr.cells, ctr.n_ext, m.bound_stop);
    TG_T(ctr, 1);
  }
  // ---- stitch (src/aligner.rs:377-406)
  if (lane == 0) {
    tg_ops_push(out, TG_OP_MATCH, len);
    tg_ops_append_reversed(out, tmp);
  }
  out.n = (uint32_t)w.shfl((int)out.n, 0);
  w.sync();
  aln.score = rl.score + (int32_t)len + rr.score;
  aln.ystart = (uint32_t)(r_abs - (uint64_t)rl.yend);
  aln.yend = (uint32_t)(r_abs + len + (uint64_t)rr.yend);
  aln.xstart = q - (uint32_t)rl.xend;
  aln.xend = q + len + (uint32_t)rr.xend;
  TG_T(ctr, 2);
}

// lift_mem_to_tx (src/txome.rs:82-103): first exon in tx order intersecting the seed
// (exon_at / sum_at: the exon that was found and the transcript offset of its first symbol, for callers that go on walking)
TG_HD bool tg_lift_mem_to_tx(const uint32_t* te_start, const uint32_t* te_end, uint32_t e0, uint32_t e1,
                             uint32_t ref_idx, uint32_t q, uint32_t len, uint32_t& t_ref, uint32_t& t_q, uint32_t& t_len,
                             uint32_t* exon_at = nullptr, uint32_t* sum_at = nullptr) {
  uint32_t exon_sum = 0;
  for (uint32_t e = e0; e < e1; e++) {
    uint32_t es = TG_LDG(te_start + e), ee = TG_LDG(te_end + e);
    uint32_t a0 = ref_idx, a1 = ref_idx + len;
    if ((a0 >= es && a0 < ee) || (es >= a0 && es < a1)) {
      uint32_t start = (a0 > es ? a0 - es : 0) + exon_sum;
      uint32_t start_off = es > a0 ? es - a0 : 0;
      uint32_t end = (a1 < ee ? a1 : ee) - es + exon_sum;
      t_ref = start; t_q = q + start_off; t_len = end - start;
      if (exon_at) { *exon_at = e; *sum_at = exon_sum; }
      return true;
    }
    exon_sum += ee - es;
  }
  return false;  // reference: unreachable!()
}

// lift_tx_to_gx (src/txome.rs:110-160) on RLE words.  Before EVERY unit op the reference advances at most
// one exon when the running transcript coordinate sits on an exon end, pushing Yclip(intron length) -- also
// in front of ops that consume no reference (the documented trailing-Ins quirk).
TG_HD void tg_lift_tx_to_gx(const uint32_t* te_start, const uint32_t* te_end, uint32_t e0, uint32_t e1,
                            const TgOps& tx_ops, uint32_t tx_ystart, uint32_t& g_ystart, uint32_t& g_yend, TgOps& out) {
  out.n = 0;
  uint32_t i = tx_ystart, exon_sum = 0, ex = e0;
  while (exon_sum + (TG_LDG(te_end + ex) - TG_LDG(te_start + ex)) <= i) {
    exon_sum += TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
    ex++;
  }
  g_ystart = TG_LDG(te_start + ex) + (i - exon_sum);
  for (uint32_t k = 0; k < tx_ops.n; k++) {
    uint32_t kind = tx_ops.w[k] & 7u, run = tx_ops.w[k] >> 3;
    bool consumes = kind == TG_OP_MATCH || kind == TG_OP_SUBST || kind == TG_OP_DEL;
    uint32_t units = (kind <= TG_OP_INS) ? run : 1u;  // a clip is one op
    while (units > 0) {
      uint32_t elen = TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
      if (ex + 1 < e1 && exon_sum + elen <= i) {
        exon_sum += elen;
        ex++;
        tg_ops_push(out, TG_OP_YCLIP, TG_LDG(te_start + ex) - TG_LDG(te_end + ex - 1));
        elen = TG_LDG(te_end + ex) - TG_LDG(te_start + ex);
      }
      if (kind > TG_OP_INS) {  // Xclip
        tg_ops_push(out, kind, run);
        units = 0;
      } else if (!consumes) {
        tg_ops_push(out, kind, units);
        units = 0;
      } else {
        uint32_t room = (ex + 1 < e1) ? (exon_sum + elen - i) : units;  // units until the next boundary check fires
        uint32_t take = units < room ? units : room;
        if (take == 0) take = 1;  // zero-length exon guard (cannot happen for valid GTF)
        tg_ops_push(out, kind, take);
        i += take;
        units -= take;
      }
    }
  }
  g_yend = TG_LDG(te_start + ex) + (i - exon_sum);
}

// Candidate record kept per accepted hit (before the end-of-read filters)
struct TgCand {
  tg_aln a;           // ops_off / tx_ops_off index the warp's ops arena
  uint32_t name_rank; // sort key of filter_overlapping
};

// align_seed_hit (src/aligner.rs:198-314).  Fills `c` and leaves gx ops in gx_ops / tx ops in tx_ops
// (views into the warp buffers).
template <class W, int RMAX = 16>
TG_HDN void tg_align_seed_hit(W& w, TgWarpMem& m, const TgAlignParams& P, uint32_t L, uint32_t ref_idx, uint32_t q,
                              uint32_t len, uint32_t bw, int32_t x_drop, TgCand& c, TgOps& gx_ops, TgOps& tx_ops,
                              TgCounters& ctr) {
  const TgIndexDev& ix = P.ix;
  const int lane = w.lane();
  TG_TDECL();
  uint32_t ref_id = tg_idx_to_ref(ix.refs, ix.n_refs, ref_idx);
  TgRef aref = ix.refs[ref_id];
  // genome window (src/aligner.rs:212-215)
  uint64_t span = (uint64_t)L + bw;
  uint64_t seq_start = ref_idx > span ? ref_idx - span : 0;
  if (seq_start < aref.start_idx) seq_start = aref.start_idx;
  uint64_t seq_end = (uint64_t)ref_idx + len + L + bw;
  if (seq_end > (uint64_t)aref.end_idx - 1) seq_end = (uint64_t)aref.end_idx - 1;
  TgAln gx;
  TgOps A{m.opsA, 0};
  TgProblem pg{ix.text4, seq_start, seq_end, ref_idx, q, len}, pb = pg;
  unsigned long long g_cells, g_ext, b_cells = 0, b_ext = 0;
  TG_T(ctr, 3);
  {
    unsigned long long c0 = ctr.cells, e0n = ctr.n_ext;
    tg_extend_left_right<W, RMAX>(w, m, ix.text4, seq_start, seq_end, ref_idx, q, len, L, bw, x_drop, gx, A, ctr);
    // counters live on different lanes (cells: the DP's last lane, extensions: lane 0): take the warp total
    g_cells = w.sum64(ctr.cells - c0);
    g_ext = w.sum64(ctr.n_ext - e0n);
  }

  // transcripts whose exon intersects the SEED (src/aligner.rs:231-258)
  bool have_tx = false;
  uint32_t best_tx = 0;
  TgAln best{0, 0, 0, 0, 0};
  uint32_t best_tlen = 0;
  TgOps Bcur{m.opsB, 0}, Bbest{m.opsC, 0};
  const TgStabRange xr = tg_stab_begin<W>(w, ix.exon_stab, ix.n_exon_stab, ix.exon_maxlen, ref_idx, ref_idx + len);
  uint32_t next_rank = 0;
  TG_T0();
  for (;;) {
    uint32_t tx_idx = 0, xrank = 0;
    const bool more = tg_stab_next<W>(w, ix.exon_stab, xr, next_rank, xrank, tx_idx);
    TG_T(ctr, 4);
    if (!more) break;
    next_rank = xrank + 1;
    uint32_t e0 = TG_LDG(ix.tx_exon_off + tx_idx), e1 = TG_LDG(ix.tx_exon_off + tx_idx + 1);
    uint64_t t0 = TG_LDG(ix.tx_seq_off + tx_idx), t1 = TG_LDG(ix.tx_seq_off + tx_idx + 1);
    uint32_t tlen = (uint32_t)(t1 - t0);
    uint32_t tr = 0, tq = 0, tl = 0;
    if (!tg_lift_mem_to_tx(ix.te_start, ix.te_end, e0, e1, ref_idx, q, len, tr, tq, tl)) continue;
    // extend_seed_match (src/aligner.rs:410-426): right, then left
    tl += tg_match_fwd(m.rp, tq + tl, L, ix.txseq4, t0 + tr + tl, t1);
    {
      uint32_t back = tg_match_bwd(m.rp, tq, ix.txseq4, t0 + tr, t0);
      tr -= back; tq -= back; tl += back;
    }
    TgAln ta;
    TgProblem pt{ix.txseq4, t0, t1, t0 + tr, tq, tl};
    unsigned long long c0 = ctr.cells, e0n = ctr.n_ext;
    bool reuse_best = false;
    TG_T(ctr, 5);
    if (tg_same_problem<W>(w, pt, pg, L, bw)) {
      // same DP as the genome extension: reuse its result (the reference recomputes it; count its cells)
      ta = gx;
      ta.ystart = (uint32_t)(pt.r_abs - ((uint64_t)ref_idx - gx.ystart));
      ta.yend = (uint32_t)(pt.r_abs + ((uint64_t)gx.yend - ref_idx));
      for (uint32_t i = lane; i < A.n; i += W::LANES) Bcur.w[i] = A.w[i];
      Bcur.n = A.n;
      if (lane == 0) { ctr.cells += g_cells; ctr.n_ext += g_ext; }
      w.sync();
    } else if (have_tx && tg_same_problem<W>(w, pt, pb, L, bw)) {
      // same DP as the best transcript so far: equal score, so it cannot replace it (:249)
      ta = best;
      reuse_best = true;
      if (lane == 0) { ctr.cells += b_cells; ctr.n_ext += b_ext; }
    } else {
      tg_extend_left_right<W, RMAX>(w, m, ix.txseq4, t0, t1, t0 + tr, tq, tl, L, bw, x_drop, ta, Bcur, ctr);
    }
    TG_T0();
    unsigned long long t_cells = w.sum64(ctr.cells - c0), t_ext = w.sum64(ctr.n_ext - e0n);
    if (!reuse_best) {
      ta.ystart -= (uint32_t)t0;  // back to transcript coordinates (t0 < 2^32 is checked at index build)
      ta.yend -= (uint32_t)t0;
      if (!have_tx || ta.score > best.score) {  // strictly better only: first wins ties (:249)
        have_tx = true; best = ta; best_tx = tx_idx; best_tlen = tlen;
        pb = pt; b_cells = t_cells; b_ext = t_ext;
        uint32_t* t = Bcur.w; Bcur.w = Bbest.w; Bbest.w = t;
        Bbest.n = Bcur.n;
      }
    }
    TG_T(ctr, 6);
    if (ta.score >= (int32_t)L) break;  // :253-257
  }
  TG_T0();

  tg_aln& a = c.a;
  a.ref_id = ref_id;
  a.strand = (uint8_t)(aref.strand_rank & 1u);
  a.primary = 0; a.pad = 0;
  a.xlen = L;
  c.name_rank = aref.strand_rank >> 1;
  uint32_t ys, ye;
  if (have_tx && best.score >= gx.score) {  // ties -> Exonic (:263)
    uint32_t e0 = TG_LDG(ix.tx_exon_off + best_tx), e1 = TG_LDG(ix.tx_exon_off + best_tx + 1);
    TgOps G{m.opsA, 0};
    if (lane == 0) tg_lift_tx_to_gx(ix.te_start, ix.te_end, e0, e1, Bbest, best.ystart, ys, ye, G);
    G.n = (uint32_t)w.shfl((int)G.n, 0);
    ys = (uint32_t)w.shfl((int)ys, 0);
    ye = (uint32_t)w.shfl((int)ye, 0);
    w.sync();
    a.aln_type = TG_ALN_EXONIC;
    a.score = best.score; a.xstart = best.xstart; a.xend = best.xend;
    a.tx_or_gene_idx = best_tx;
    a.tx_score = best.score; a.tx_ystart = best.ystart; a.tx_yend = best.yend; a.tx_ylen = best_tlen;
    a.tx_xstart = best.xstart; a.tx_xend = best.xend;
    gx_ops = G;
    tx_ops = Bbest;
  } else {
    // first gene (in find order) whose span intersects the ALIGNMENT (src/aligner.rs:283-306)
    uint32_t gene = 0, grank = 0;
    const TgStabRange gr = tg_stab_begin<W>(w, ix.gene_stab, ix.n_gene_stab, ix.gene_maxlen, gx.ystart, gx.yend);
    const bool found = tg_stab_next<W>(w, ix.gene_stab, gr, 0u, grank, gene);
    a.aln_type = found ? TG_ALN_INTRONIC : TG_ALN_INTERGENIC;
    a.tx_or_gene_idx = found ? gene : 0xFFFFFFFFu;
    a.score = gx.score; a.xstart = gx.xstart; a.xend = gx.xend;
    a.tx_score = 0; a.tx_ystart = 0; a.tx_yend = 0; a.tx_ylen = 0; a.tx_xstart = 0; a.tx_xend = 0;
    ys = gx.ystart; ye = gx.yend;
    gx_ops = A;
    tx_ops = TgOps{m.opsC, 0};
  }
  // concat_to_chr_aln (src/aligner.rs:429-449) -- keyed on the alignment's own ystart
  uint32_t rid2 = tg_idx_to_ref(ix.refs, ix.n_refs, ys);
  TgRef r2 = ix.refs[rid2];
  if (r2.strand_rank & 1u) {
    a.ystart = ys - r2.start_idx;
    a.yend = ye - r2.start_idx;
  } else {
    a.ystart = (uint64_t)r2.len - (ye - r2.start_idx);
    a.yend = (uint64_t)r2.len - (ys - r2.start_idx);
    if (lane == 0) tg_ops_reverse(gx_ops);
    w.sync();
  }
  a.ylen = r2.len;
  TG_T(ctr, 7);
}

// stable bottom-up merge sort of an index array (a -> sorted in a; b is scratch of n entries)
template <class T, class Less>
TG_HD void tg_merge_sort_idx(T* a, T* b, uint32_t n, Less less) {
  T* src = a;
  T* dst = b;
  for (uint32_t width = 1; width < n; width <<= 1) {
    for (uint32_t lo = 0; lo < n; lo += 2 * width) {
      uint32_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
      uint32_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) dst[k++] = less(src[j], src[i]) ? src[j++] : src[i++];  // ties keep the left run first
      while (i < mid) dst[k++] = src[i++];
      while (j < hi) dst[k++] = src[j++];
    }
    T* t = src; src = dst; dst = t;
  }
  if (src != a)
    for (uint32_t i = 0; i < n; i++) a[i] = src[i];
}

// End-of-read filters on the accepted candidates (src/aligner.rs:177-187 and filter_overlapping :317-349).
// Serial; `order`/`tmp` are scratch arrays of n entries.  Returns the number of output records; order[0..ret)
// lists candidate indices in output order.
TG_HD uint32_t tg_finalize_read(const TgCand* cands, uint32_t n, int32_t max_aln_score, int32_t range,
                                uint16_t* order, uint16_t* tmp) {
  // retain (:177-179)
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++)
    if (cands[i].a.score >= max_aln_score - range) order[m++] = (uint16_t)i;
  if (m == 0) return 0;
  // stable sort by (ref_name, strand false<true, ystart) (:322-327)
  tg_merge_sort_idx(order, tmp, m, [&](uint16_t x, uint16_t y) {
    const TgCand& a = cands[x];
    const TgCand& b = cands[y];
    if (a.name_rank != b.name_rank) return a.name_rank < b.name_rank;
    if (a.a.strand != b.a.strand) return a.a.strand < b.a.strand;
    return a.a.ystart < b.a.ystart;
  });
  // sweep (:329-346)
  uint32_t k = 0;
  uint64_t max_end = 0;
  for (uint32_t i = 0; i < m; i++) {
    const TgCand& cc = cands[order[i]];
    bool fresh = k == 0 || cc.a.ystart >= max_end || cc.name_rank != cands[tmp[k - 1]].name_rank ||
                 cc.a.strand != cands[tmp[k - 1]].a.strand;
    if (fresh) {
      max_end = cc.a.yend;
      tmp[k++] = order[i];
    } else {
      if (cc.a.score > cands[tmp[k - 1]].a.score) tmp[k - 1] = order[i];
      uint64_t ce = cands[tmp[k - 1]].a.yend;
      if (ce > max_end) max_end = ce;
    }
  }
  // stable sort by -score (:183)
  for (uint32_t i = 0; i < k; i++) order[i] = tmp[i];
  tg_merge_sort_idx(order, tmp, k, [&](uint16_t x, uint16_t y) { return cands[x].a.score > cands[y].a.score; });
  return k;
}

// ------------------------------------------------------------------------------------------------
// k-mer table build over the suffix array (one thread per SA row on the device).
// A row starts a k-mer group when its first k symbols contain no '$' and differ from the previous row's.
// ------------------------------------------------------------------------------------------------
TG_HD bool tg_kmer_valid(const uint64_t* text4, uint64_t text_len, uint32_t pos, uint32_t k, uint64_t& w0, uint64_t& w1) {
  if ((uint64_t)pos + k > text_len) return false;
  w0 = tg_ld16(text4, pos) & tg_top_nibbles(k);
  w1 = k > 16 ? (tg_ld16(text4, (uint64_t)pos + 16) & tg_top_nibbles(k - 16)) : 0;
  // a zero nibble inside the first k symbols is a '$'
  uint64_t m0 = w0 | ~tg_top_nibbles(k);  // force the unused nibbles non-zero
  if (tg_has_zero_nibble(m0)) return false;
  if (k > 16) {
    uint64_t m1 = w1 | ~tg_top_nibbles(k - 16);
    if (tg_has_zero_nibble(m1)) return false;
  }
  return true;
}
// Is SA row r the first row of a valid k-mer group?  If so returns its key words.
TG_HD bool tg_kmer_group_start(const uint64_t* text4, uint64_t text_len, const uint32_t* sa, uint64_t r, uint32_t k,
                               uint64_t& w0, uint64_t& w1) {
  if (!tg_kmer_valid(text4, text_len, TG_LDG(sa + r), k, w0, w1)) return false;
  if (r == 0) return true;
  uint64_t p0, p1;
  if (!tg_kmer_valid(text4, text_len, TG_LDG(sa + r - 1), k, p0, p1)) return true;
  return p0 != w0 || p1 != w1;
}
// number of consecutive rows from r that share the k-mer (galloping + binary search)
TG_HD uint32_t tg_kmer_group_count(const uint64_t* text4, uint64_t text_len, const uint32_t* sa, uint64_t r, uint32_t k,
                                   uint64_t w0, uint64_t w1) {
  auto same = [&](uint64_t row) {
    uint64_t a0, a1;
    return tg_kmer_valid(text4, text_len, TG_LDG(sa + row), k, a0, a1) && a0 == w0 && a1 == w1;
  };
  uint64_t step = 1, lo = r;  // lo: last row known to match
  while (r + step < text_len && same(r + step)) { lo = r + step; step <<= 1; }
  uint64_t hi = r + step < text_len ? r + step : text_len;  // first row known NOT to match (or end)
  while (lo + 1 < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    if (same(mid)) lo = mid; else hi = mid;
  }
  return (uint32_t)(lo - r + 1);
}

// ------------------------------------------------------------------------------------------------
// Read-level drivers.  W adds: atomic_add(unsigned long long*, v) and atomic_or(int*, v).
// ------------------------------------------------------------------------------------------------
enum { TG_FLAG_SEED_POOL = 1, TG_FLAG_ALN_POOL = 2, TG_FLAG_OPS_POOL = 4, TG_FLAG_READ_CAP = 8, TG_FLAG_ARENA = 16,
       // an alignment whose own start lies on another Ref than its hit (concat_to_chr_aln keys ylen on the alignment,
       // ref_name / strand on the hit): cannot happen -- windows are clamped to the hit's Ref and a transcript's exons lie on
       // one Ref -- and the compact record (tg_aln_c) relies on it; checked so that it would fail loudly
       TG_FLAG_YLEN = 256 };

struct TgSeedMem {  // per-warp scratch of the seeding stage
  uint64_t* rp;       // packed read, (maxL/16 + 3) words
  TgSeedHit* hits;    // maxL
  tg_seed* sm;        // maxL
};
struct TgSeedOut {
  tg_seed* pool;
  unsigned long long* pool_used;
  unsigned long long pool_cap;
  uint64_t* read_first;
  uint32_t* read_count;
  int* flags;
  unsigned long long* n_smems;
};

template <class W>
TG_HDN void tg_seed_read(W& w, TgSeedMem& m, const uint8_t* bases, uint64_t off, uint32_t L, uint32_t k,
                         const TgSlot* slots, uint64_t slot_mask, const uint64_t* text4, const uint32_t* sa,
                         const TgSeedOut& out, uint32_t r) {
  const int lane = w.lane();
  const uint32_t nw = L / 16 + 3;
  for (uint32_t wi = lane; wi < nw; wi += W::LANES) {
    uint64_t word = 0;
    for (uint32_t t = 0; t < 16; t++) {
      uint32_t p = wi * 16 + t;
      uint64_t c = p < L ? tg_ascii_code(TG_LDG(bases + off + p)) : (uint64_t)TG_C_PAD;
      word |= c << ((15 - t) * 4);
    }
    m.rp[wi] = word;
  }
  w.sync();
  for (uint32_t q = lane; q + k <= L; q += W::LANES) {
    TgSeedHit h;
    tg_seed_offset(m.rp, L, q, k, slots, slot_mask, text4, sa, h);
    m.hits[q] = h;
  }
  w.sync();
  int n = 0;
  if (lane == 0) n = (int)tg_smem_select(m.hits, L, k, m.sm);
  n = w.shfl(n, 0);
  unsigned long long base = 0;
  if (lane == 0 && n > 0) base = w.atomic_add(out.pool_used, (unsigned long long)n);
  base = w.shfl64(base, 0);
  w.sync();
  if (base + (unsigned long long)n > out.pool_cap) {
    if (lane == 0) w.atomic_or(out.flags, TG_FLAG_SEED_POOL);
    n = 0;
  }
  for (int i = lane; i < n; i += W::LANES) out.pool[base + i] = m.sm[i];
  if (lane == 0) {
    out.read_first[r] = base;
    out.read_count[r] = (uint32_t)n;
    if (n > 0) w.atomic_add(out.n_smems, (unsigned long long)n);
  }
  w.sync();
}

// ---- thread-per-offset seeding (the device pipeline): pack -> probe -> select --------------------------------------------
// One packed word (16 symbols) of a read; positions >= L are padding.
TG_HD uint64_t tg_pack_word(const uint8_t* bases, uint64_t off, uint32_t L, uint32_t wi) {
  uint64_t word = 0;
  for (uint32_t t = 0; t < 16; t++) {
    const uint32_t p = wi * 16 + t;
    const uint64_t c = p < L ? tg_ascii_code(TG_LDG(bases + off + p)) : (uint64_t)TG_C_PAD;
    word |= c << ((15 - t) * 4);
  }
  return word;
}
// Probe schedule.  E(q) is non-decreasing in q, so two probed offsets qa < qb with the same end E(qa) == E(qb) != 0 pin
// E(q) = E(qa) for every q between them, and none of those q starts an SMEM: they need no table probe at all.
//   wave 0: q = 0.  If the whole read matches (E(0) == L) nothing else is probed.
//   wave 1: the sample offsets S, 2S, ... and the last offset L - k.
//   wave 2: every other offset, probed only when its two bracketing samples disagree (or found nothing).
TG_HD bool tg_probe_is_sample(uint32_t q, uint32_t q_last) { return q % TG_PROBE_STRIDE == 0 || q == q_last; }
// the j-th wave-1 offset of a read whose last offset is q_last (0xFFFFFFFF: none)
TG_HD uint32_t tg_probe_sample(uint32_t j, uint32_t q_last) {
  const uint32_t q = (j + 1) * TG_PROBE_STRIDE;
  if (q < q_last) return q;
  return q - TG_PROBE_STRIDE < q_last ? q_last : 0xFFFFFFFFu;
}
// wave 2: true when the bracketing samples make the probe unnecessary (the entry is then never written nor read:
// tg_smem_count / tg_smem_select step over closed brackets)
TG_HD bool tg_probe_bracketed(const TgSeedHit* row, uint32_t q, uint32_t q_last) {
  const uint32_t qa = q - q % TG_PROBE_STRIDE;
  const uint32_t qb = qa + TG_PROBE_STRIDE < q_last ? qa + TG_PROBE_STRIDE : q_last;
  const uint32_t ea = row[qa].e, eb = row[qb].e;
  return ea == eb && ea != 0;
}
// at sample offset q of a sampled probe table: the offset to continue with (the other end of a closed bracket, else q + 1)
TG_HD uint32_t tg_probe_step(const TgSeedHit* row, uint32_t q, uint32_t q_last) {
  if (q % TG_PROBE_STRIDE != 0 || q >= q_last) return q + 1;
  const uint32_t qb = q + TG_PROBE_STRIDE < q_last ? q + TG_PROBE_STRIDE : q_last;
  const uint32_t ea = row[q].e;
  return (qb > q + 1 && ea != 0 && ea == row[qb].e) ? qb : q + 1;
}

// Index::all_smems for one read from its probe results: SMEM records into the seed pool.
template <class W>
TG_HDN void tg_seed_select_read(W& w, const TgSeedHit* hits, uint32_t L, uint32_t k, const TgSeedOut& out, uint32_t r) {
  const uint32_t n = tg_smem_count(hits, L, k, true);
  unsigned long long base = 0;
  uint32_t nn = n;
  if (n) {
    base = w.atomic_add(out.pool_used, (unsigned long long)n);
    if (base + n > out.pool_cap) {
      w.atomic_or(out.flags, TG_FLAG_SEED_POOL);
      nn = 0;
    }
  }
  if (nn) tg_smem_select(hits, L, k, out.pool + base, true);
  out.read_first[r] = base;
  out.read_count[r] = nn;
  if (nn) w.atomic_add(out.n_smems, (unsigned long long)nn);
}

struct TgWarpScratch {  // per-warp global scratch of the extension stage
  TgCand* cands;       // TG_MAX_ALNS_PER_READ
  uint32_t* arena;     // ops of the accepted candidates
  uint32_t arena_cap;
  uint16_t* order;     // 2 * TG_MAX_ALNS_PER_READ (order + tmp of tg_finalize_read)
};
struct TgAlignOut {
  uint64_t* read_aln_first;
  uint32_t* read_aln_count;
  tg_aln* alns;
  uint32_t* ops;
  unsigned long long* alns_used;
  unsigned long long* ops_used;
  unsigned long long alns_cap, ops_cap;
  int* flags;
  // compact output (tg_aln_c, include/thermite_gpu.h): when alns_c is set, records go there instead of `alns` and the
  // per-read first index to first32; first_base / ops_base = where this pool segment starts in the caller's result
  // (tg_multi_align_batch: one segment per GPU), added to every first index and operation offset written
  tg_aln_c* alns_c = nullptr;
  uint32_t* first32 = nullptr;
  unsigned long long first_base = 0, ops_base = 0;
};
// the GenomeAlignment record `a` (ops_off = position in out.ops, transcript operations directly behind) as record
// `idx` of the output pool
TG_HD void tg_out_write_aln(const TgAlignOut& out, unsigned long long idx, const tg_aln& a) {
  if (out.alns_c) {
    tg_aln_c c;
    c.ystart = (uint32_t)a.ystart; c.yend = (uint32_t)a.yend;
    c.tx_ystart = (uint32_t)a.tx_ystart; c.tx_yend = (uint32_t)a.tx_yend;
    c.ref_id = a.ref_id; c.tx_or_gene_idx = a.tx_or_gene_idx;
    c.ops_off = a.ops_off + (uint32_t)out.ops_base;
    c.score = (int16_t)a.score; c.xstart = (uint16_t)a.xstart; c.xend = (uint16_t)a.xend;
    c.ops_len = (uint16_t)a.ops_len; c.tx_ops_len = (uint16_t)a.tx_ops_len;
    c.aln_type = a.aln_type; c.primary = a.primary;
    out.alns_c[idx] = c;
  } else {
    out.alns[idx] = a;
  }
}
TG_HD void tg_out_write_read(const TgAlignOut& out, uint32_t r, unsigned long long first, uint32_t count) {
  if (out.first32) out.first32[r] = (uint32_t)(first + out.first_base);
  else out.read_aln_first[r] = first;
  out.read_aln_count[r] = count;
}

// align_read (src/aligner.rs:123-190) for one read, seeds already ordered as Index::all_smems orders them.
template <class W, int RMAX = 16>
TG_HDN void tg_align_read(W& w, TgWarpMem& m, const TgAlignParams& P, const uint8_t* bases, uint64_t off, uint32_t L,
                          const tg_seed* seeds, uint32_t n_seeds, const TgWarpScratch& sc, const TgAlignOut& out,
                          uint32_t r, TgCounters& ctr) {
  const int lane = w.lane();
  TG_TDECL();
  for (uint32_t i = lane; i < L; i += W::LANES) m.rd[i] = (uint8_t)tg_ascii_code(TG_LDG(bases + off + i));
  w.sync();
  for (uint32_t wi = lane; wi < L / 16 + 3; wi += W::LANES) {
    uint64_t word = 0;
    for (uint32_t t = 0; t < 16; t++) {
      uint32_t p = wi * 16 + t;
      word |= (uint64_t)(p < L ? m.rd[p] : (uint8_t)TG_C_PAD) << ((15 - t) * 4);
    }
    m.rp[wi] = word;
  }
  w.sync();
  // :130-138
  float prod = P.opts.min_aln_score_percent * (float)L;
  int32_t pct_score = (int32_t)prod;
  const int32_t min_aln_score = pct_score > P.opts.min_aln_score ? pct_score : P.opts.min_aln_score;
  int32_t max_aln_score = min_aln_score;
  uint32_t bw = (min_aln_score < 0) ? 0u : (L > (uint32_t)min_aln_score ? L - (uint32_t)min_aln_score : 0u);
  uint32_t x_drop = bw;
  const int32_t range = (int32_t)P.opts.multimap_score_range;
  uint32_t n_acc = 0, arena_used = 0;
  bool capped = false;
  TG_T(ctr, 8);

  for (uint32_t si = 0; si < n_seeds && !capped; si++) {
    tg_seed sd = seeds[si];
    for (uint32_t rk = sd.count; rk-- > 0 && !capped;) {  // reversed SA-rank order (src/index.rs:251-253)
      uint32_t ref_idx = sd.direct ? sd.sa_lo : TG_LDG(P.ix.sa + sd.sa_lo + rk);
      if (lane == 0) ctr.hits++;
      TgCand c;
      TgOps gx_ops{nullptr, 0}, tx_ops{nullptr, 0};
      tg_align_seed_hit<W, RMAX>(w, m, P, L, ref_idx, sd.query_idx, sd.len, bw, (int32_t)x_drop, c, gx_ops, tx_ops, ctr);
      TG_T0();
      if (lane == 0 && c.a.ylen != P.ix.refs[c.a.ref_id].len) w.atomic_or(out.flags, TG_FLAG_YLEN);
      if (!P.opts.intron_mode && c.a.aln_type != TG_ALN_EXONIC) continue;  // :146-151
      int32_t s = c.a.score;
      if (s < P.opts.min_aln_score || s < min_aln_score || s < max_aln_score - range) continue;  // :154-159
      // :162-171 (`score as usize` wraps for negative scores => saturating_sub gives 0)
      uint32_t lim = (s < 0) ? 0u : ((L + P.opts.multimap_score_range > (uint32_t)s) ? L + P.opts.multimap_score_range - (uint32_t)s : 0u);
      if (lim < bw) bw = lim;
      if (lim < x_drop) x_drop = lim;
      if (s > max_aln_score) max_aln_score = s;
      // keep the candidate
      uint32_t need = gx_ops.n + tx_ops.n;
      if (n_acc >= TG_MAX_ALNS_PER_READ || arena_used + need > sc.arena_cap) {
        // drop what the final `retain` (:177-179) would drop anyway, compacting records and their ops in place
        w.sync_global();
        int kept = 0, au = 0;
        if (lane == 0) {
          uint32_t o = 0, kk = 0;
          for (uint32_t i = 0; i < n_acc; i++) {
            TgCand ci = sc.cands[i];
            if (ci.a.score < max_aln_score - range) continue;
            uint32_t nn = ci.a.ops_len + ci.a.tx_ops_len, src = ci.a.ops_off;
            for (uint32_t t = 0; t < nn; t++) sc.arena[o + t] = sc.arena[src + t];
            ci.a.ops_off = o; ci.a.tx_ops_off = o + ci.a.ops_len;
            sc.cands[kk++] = ci;
            o += nn;
          }
          kept = (int)kk; au = (int)o;
        }
        n_acc = (uint32_t)w.shfl(kept, 0);
        arena_used = (uint32_t)w.shfl(au, 0);
        w.sync_global();
      }
      if (n_acc >= TG_MAX_ALNS_PER_READ || arena_used + need > sc.arena_cap) {
        if (lane == 0) w.atomic_or(out.flags, n_acc >= TG_MAX_ALNS_PER_READ ? TG_FLAG_READ_CAP : TG_FLAG_ARENA);
        capped = true;
        break;
      }
      c.a.ops_off = arena_used; c.a.ops_len = gx_ops.n;
      c.a.tx_ops_off = arena_used + gx_ops.n; c.a.tx_ops_len = tx_ops.n;
      for (uint32_t i = lane; i < gx_ops.n; i += W::LANES) sc.arena[arena_used + i] = gx_ops.w[i];
      for (uint32_t i = lane; i < tx_ops.n; i += W::LANES) sc.arena[arena_used + gx_ops.n + i] = tx_ops.w[i];
      if (lane == 0) sc.cands[n_acc] = c;
      arena_used += need;
      n_acc++;
      w.sync();
      TG_T(ctr, 9);
    }
  }
  w.sync_global();
  TG_T0();
  // :177-187
  uint16_t* order = sc.order;
  uint16_t* tmp = order + TG_MAX_ALNS_PER_READ;
  int k = 0;
  unsigned long long words = 0;
  if (lane == 0) {
    k = (int)tg_finalize_read(sc.cands, n_acc, max_aln_score, range, order, tmp);
    for (int i = 0; i < k; i++) words += sc.cands[order[i]].a.ops_len + sc.cands[order[i]].a.tx_ops_len;
  }
  k = w.shfl(k, 0);
  words = w.shfl64(words, 0);
  unsigned long long abase = 0, obase = 0;
  if (lane == 0 && k > 0) {
    abase = w.atomic_add(out.alns_used, (unsigned long long)k);
    obase = w.atomic_add(out.ops_used, words);
  }
  abase = w.shfl64(abase, 0);
  obase = w.shfl64(obase, 0);
  w.sync();
  if (abase + (unsigned long long)k > out.alns_cap || obase + words > out.ops_cap || obase + words + out.ops_base > 0xFFFFFFFFull) {
    if (lane == 0) w.atomic_or(out.flags, abase + (unsigned long long)k > out.alns_cap ? TG_FLAG_ALN_POOL : TG_FLAG_OPS_POOL);
    k = 0;
  }
  unsigned long long o = obase;
  for (int i = 0; i < k; i++) {
    const TgCand& c = sc.cands[order[i]];
    uint32_t n1 = c.a.ops_len, n2 = c.a.tx_ops_len, src = c.a.ops_off;
    for (uint32_t t = lane; t < n1 + n2; t += W::LANES) out.ops[o + t] = sc.arena[src + t];
    if (lane == 0) {
      tg_aln a = c.a;
      a.ops_off = (uint32_t)o;
      a.tx_ops_off = (uint32_t)(o + n1);
      if (a.aln_type != TG_ALN_EXONIC) { a.tx_ops_off = 0; a.tx_ops_len = 0; }
      a.primary = i == 0 ? 1 : 0;
      tg_out_write_aln(out, abase + i, a);
    }
    o += n1 + n2;
  }
  if (lane == 0) tg_out_write_read(out, r, abase, (uint32_t)k);
  w.sync();
  TG_T(ctr, 10);
}
