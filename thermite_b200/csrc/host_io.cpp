// host_io.cpp -- host-side ingest and output of libthermite_gpu: FASTQ text -> read batches, and flat
// alignment records -> PAF / SAM text.  Replaces the per-read loop of align_reads_from_file
// (reference src/aligner.rs:51-115) and the record builders of src/aln_writer.rs:47-116, 118-253, 256-358.
// SAM text follows noodles 0.1.0's writer as far as it can be recalled (not vendored: unpinned).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "tg_internal.h"
#include "host_text.h"

namespace {

// multimapq (src/aln_writer.rs:332-340)
unsigned mapq_of(uint64_t n) {
  if (n <= 1) return 255;
  if (n >= 5) return 0;
  static const unsigned tab[5] = {255, 255, (unsigned)std::lround(-10.0f * std::log10(1.0f - 1.0f / 2.0f)),
                                  (unsigned)std::lround(-10.0f * std::log10(1.0f - 1.0f / 3.0f)),
                                  (unsigned)std::lround(-10.0f * std::log10(1.0f - 1.0f / 4.0f))};
  return tab[n];
}

// decimal digits of v at d, two at a time; returns the end
inline char* put_num(char* d, uint64_t v) {
  static const char lut[201] =
      "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263"
      "646566676869707172737475767778798081828384858687888990919293949596979899";
  char buf[24];
  int k = 24;
  while (v >= 100) {
    const unsigned q = (unsigned)(v % 100);
    v /= 100;
    buf[--k] = lut[2 * q + 1]; buf[--k] = lut[2 * q];
  }
  if (v >= 10) { buf[--k] = lut[2 * v + 1]; buf[--k] = lut[2 * v]; }
  else buf[--k] = (char)('0' + v);
  memcpy(d, buf + k, (size_t)(24 - k));
  return d + (24 - k);
}

// to_noodles_cigar (src/aln_writer.rs:279-323): Match and Subst both print as M and merge.  Writes at d (at most 21 bytes
// per operation word), returns the end.
char* put_cigar(char* d, const uint32_t* w, uint32_t n) {
  static const char sym[6] = {'M', 'M', 'D', 'I', 'S', 'N'};
  uint32_t i = 0;
  while (i < n) {
    uint32_t kind = w[i] & 7u, run = w[i] >> 3;
    if (kind <= TG_OP_SUBST) {
      uint64_t tot = run;
      while (i + 1 < n && (w[i + 1] & 7u) <= TG_OP_SUBST) { tot += w[i + 1] >> 3; i++; }
      d = put_num(d, tot);
      *d++ = 'M';
    } else if (kind >= TG_OP_XCLIP) {
      // equal adjacent clips collapse into one entry carrying a single clip's length (:309-316, :292-294)
      while (i + 1 < n && w[i + 1] == w[i]) i++;
      d = put_num(d, run);
      *d++ = sym[kind];
    } else {
      d = put_num(d, run);
      *d++ = sym[kind];
    }
    i++;
  }
  return d;
}
inline char* put_str(char* d, const char* lit, size_t n) { memcpy(d, lit, n); return d + n; }
#define PUT_LIT(d, lit) put_str(d, lit, sizeof(lit) - 1)

char comp(char c) {  // bio::alphabets::dna::revcomp keeps case and maps IUPAC codes; ACGTN is what reads carry
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;
  }
}

}  // namespace

extern "C" {

tg_status tg_format_sam_header(const tg_index_host* ix, char** out, size_t* out_len) {
  if (!ix || !out || !out_len) return tg_fail(TG_ERR_INVALID, "null argument");
  // build_sam_header (src/aln_writer.rs:256-276): a map keyed by name, so the two strands collapse
  TgOut o;
  std::vector<std::string> seen;
  for (uint32_t i = 0; i < ix->hdr()->n_refs; i++) {
    const std::string& nm = ix->ref_names[i];
    bool dup = false;
    for (auto& s : seen) if (s == nm) { dup = true; break; }
    if (dup) continue;
    seen.push_back(nm);
    uint64_t v[4];
    tg_index_host_ref(ix, i, v);
    o.s += "@SQ\tSN:"; o.s += nm; o.s += "\tLN:"; o.num(v[2]); o.s.push_back('\n');
  }
  o.s += "@PG\tID:thermite\n";
  *out = (char*)malloc(o.s.size() + 1);
  memcpy(*out, o.s.data(), o.s.size());
  (*out)[o.s.size()] = 0;
  *out_len = o.s.size();
  return TG_OK;
}

}  // extern "C"

// records of the reads [r0, r1) as PAF / SAM text (appended to o)
void tg_format_reads(const tg_index_host* ix, const TgRecView& v, const uint8_t* bases, const uint64_t* offs,
                     const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                     int sam, uint32_t r0, uint32_t r1, TgOut& o) {
  const TgBlobHeader* bh = ix->hdr();
  const TgRef* refs = (const TgRef*)(ix->blob.data() + bh->off_refs);
  const uint64_t* tso = (const uint64_t*)(ix->blob.data() + bh->off_tx_seq_off);
  for (uint32_t r = r0; r < r1; r++) {
    // the record pool is in the device's allocation order, not in read order: fetch ahead (records 8 reads ahead, their
    // operation words 4 reads ahead)
    if (r + 8 < r1 && v.count[r + 8]) {
      const uint64_t f8 = v.first64 ? v.first64[r + 8] : (uint64_t)v.first32[r + 8];
      if (v.wide) __builtin_prefetch(v.wide + f8); else __builtin_prefetch(v.comp + f8);
    }
    if (r + 4 < r1 && v.count[r + 4]) {
      const uint64_t f4 = v.first64 ? v.first64[r + 4] : (uint64_t)v.first32[r + 4];
      __builtin_prefetch(v.ops + (v.wide ? v.wide[f4].ops_off : v.comp[f4].ops_off));
    }
    const char* nm = (const char*)names + name_offs[r];
    size_t nm_len = name_offs[r + 1] - name_offs[r];
    size_t qn_len = nm_len;  // format_read_name: cut at the first space (src/aln_writer.rs:344-349)
    if (sam) for (size_t i = 0; i < nm_len; i++) if (nm[i] == ' ') { qn_len = i; break; }
    const uint64_t L = offs[r + 1] - offs[r];
    const uint32_t cnt = v.count[r];
    const uint64_t first = v.first64 ? v.first64[r] : (uint64_t)v.first32[r];
    if (cnt == 0) {
      if (sam) {  // unmapped_sam_record (src/aln_writer.rs:241-253)
        o.s.append(nm, qn_len);
        o.s += "\t4\t*\t0\t255\t*\t*\t0\t0\t";
        if (L) o.s.append((const char*)bases + offs[r], L); else o.s.push_back('*');
        o.s.push_back('\t');
        size_t ql = qual_offs[r + 1] - qual_offs[r];
        if (ql) o.s.append((const char*)quals + qual_offs[r], ql); else o.s.push_back('*');
        o.s.push_back('\n');
      }
      continue;  // PAF prints nothing for unmapped reads (src/aligner.rs:77)
    }
    for (uint32_t i = 0; i < cnt; i++) {
      tg_aln a;
      if (v.wide) a = v.wide[first + i];
      else if (!tg_expand_one(refs, (uint32_t)bh->n_refs, tso, (uint32_t)bh->n_txs, v.comp[first + i], (uint32_t)L, a)) continue;
      const uint32_t* w = v.ops + a.ops_off;
      const std::string& rname = ix->ref_names[a.ref_id];
      if (!sam) {  // PafEntry (src/aln_writer.rs:47-116)
        uint64_t n_match = 0, n_match_gap = 0;
        for (uint32_t k = 0; k < a.ops_len; k++) {
          uint32_t kind = w[k] & 7u, run = w[k] >> 3;
          if (kind == TG_OP_MATCH) n_match += run;
          if (kind <= TG_OP_INS) n_match_gap += run;
          else if (kind == TG_OP_XCLIP) n_match_gap += 1;
        }
        // one reservation per line, then plain stores (a line holds the two names and 11 numbers of at most 20 digits)
        char* d = o.s.room(nm_len + rname.size() + 11 * 21 + 16);
        memcpy(d, nm, nm_len); d += nm_len; *d++ = '\t';
        d = put_num(d, L); *d++ = '\t';
        d = put_num(d, a.xstart); *d++ = '\t';
        d = put_num(d, a.xend); *d++ = '\t';
        *d++ = a.strand ? '+' : '-'; *d++ = '\t';
        memcpy(d, rname.data(), rname.size()); d += rname.size(); *d++ = '\t';
        d = put_num(d, a.ylen); *d++ = '\t';
        d = put_num(d, a.ystart); *d++ = '\t';
        d = put_num(d, a.yend); *d++ = '\t';
        d = put_num(d, n_match); *d++ = '\t';
        d = put_num(d, n_match_gap); *d++ = '\t';
        d = put_num(d, mapq_of(cnt)); *d++ = '\t'; *d++ = '\n';
        o.s.n = (size_t)(d - o.s.p);
      } else {  // aln_to_sam_record (src/aln_writer.rs:118-238)
        unsigned flags = 0;
        if (!a.strand) flags |= 0x10;
        if (!a.primary) flags |= 0x100;
        uint64_t n_mis = 0;
        for (uint32_t k = 0; k < a.ops_len; k++) if ((w[k] & 7u) == TG_OP_SUBST) n_mis += w[k] >> 3;
        // one reservation per line, then plain stores
        const std::string* tx_id = nullptr; const std::string* gid = nullptr; const std::string* gname = nullptr;
        if (a.aln_type == TG_ALN_EXONIC) {
          const uint32_t g = ix->tx_gene[a.tx_or_gene_idx];
          tx_id = &ix->tx_ids[a.tx_or_gene_idx]; gid = &ix->gene_ids[g]; gname = &ix->gene_names[g];
        } else if (a.aln_type == TG_ALN_INTRONIC) {
          gid = &ix->gene_ids[a.tx_or_gene_idx]; gname = &ix->gene_names[a.tx_or_gene_idx];
        }
        const size_t qn = qual_offs[r + 1] - qual_offs[r];
        char* d = o.s.room(qn_len + rname.size() + (size_t)L + qn + 21 * ((size_t)a.ops_len + a.tx_ops_len) + (tx_id ? tx_id->size() : 0) +
                           (gid ? gid->size() + gname->size() : 0) + 256);
        memcpy(d, nm, qn_len); d += qn_len; *d++ = '\t';
        d = put_num(d, flags); *d++ = '\t';
        memcpy(d, rname.data(), rname.size()); d += rname.size(); *d++ = '\t';
        d = put_num(d, a.ystart + 1); *d++ = '\t';
        d = put_num(d, mapq_of(cnt)); *d++ = '\t';
        d = put_cigar(d, w, a.ops_len);
        d = PUT_LIT(d, "\t*\t0\t0\t");
        const char* sq = (const char*)bases + offs[r];
        if (L == 0) *d++ = '*';
        else if (a.strand) { memcpy(d, sq, L); d += L; }
        else for (uint64_t k = L; k-- > 0;) *d++ = comp(sq[k]);
        *d++ = '\t';
        const char* ql = (const char*)quals + qual_offs[r];
        if (qn == 0) *d++ = '*';
        else if (a.strand) { memcpy(d, ql, qn); d += qn; }
        else for (size_t k = qn; k-- > 0;) *d++ = ql[k];
        d = PUT_LIT(d, "\tAS:i:");
        if (a.score < 0) { *d++ = '-'; d = put_num(d, (uint64_t)(-(int64_t)a.score)); } else d = put_num(d, (uint64_t)a.score);
        d = PUT_LIT(d, "\tNH:i:"); d = put_num(d, cnt);
        d = PUT_LIT(d, "\tHI:i:"); d = put_num(d, i + 1);
        d = PUT_LIT(d, "\tnM:i:"); d = put_num(d, n_mis);
        if (a.aln_type == TG_ALN_EXONIC) {
          d = PUT_LIT(d, "\tTX:Z:"); d = put_str(d, tx_id->data(), tx_id->size()); d = PUT_LIT(d, ",+"); d = put_num(d, a.tx_ystart); *d++ = ',';
          d = put_cigar(d, v.ops + a.tx_ops_off, a.tx_ops_len);
          d = PUT_LIT(d, "\tGX:Z:"); d = put_str(d, gid->data(), gid->size());
          d = PUT_LIT(d, "\tGN:Z:"); d = put_str(d, gname->data(), gname->size());
          d = PUT_LIT(d, "\tRE:A:E");
        } else if (a.aln_type == TG_ALN_INTRONIC) {
          d = PUT_LIT(d, "\tGX:Z:"); d = put_str(d, gid->data(), gid->size());
          d = PUT_LIT(d, "\tGN:Z:"); d = put_str(d, gname->data(), gname->size());
          d = PUT_LIT(d, "\tRE:A:N");
        } else {
          d = PUT_LIT(d, "\tRE:A:I");
        }
        *d++ = '\n';
        o.s.n = (size_t)(d - o.s.p);
      }
    }
  }
}

extern "C" {

tg_status tg_format_batch(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                          const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                          int sam, char** out, size_t* out_len) {
  if (!ix || !res || !offs || !names || !name_offs || !out || !out_len) return tg_fail(TG_ERR_INVALID, "null argument");
  if (sam && (!quals || !qual_offs || !bases)) return tg_fail(TG_ERR_INVALID, "SAM output needs bases and qualities");
  // Reads are formatted independently: contiguous ranges on the host's cores, pieces concatenated in read order
  // (byte-identical to the single-threaded text).
  const uint32_t n = res->n_reads;
  uint32_t T = (uint32_t)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 64);
  if (n < 32768) T = 1;
  std::vector<TgOut> parts(T);
  auto work = [&](uint32_t t) {
    const uint32_t r0 = (uint32_t)((uint64_t)n * t / T), r1 = (uint32_t)((uint64_t)n * (t + 1) / T);
    parts[t].s.reserve((size_t)(r1 - r0) * (sam ? 400 : 90));
    TgRecView v;
    v.first64 = res->read_aln_first; v.count = res->read_aln_count; v.wide = res->alns; v.ops = res->ops;
    tg_format_reads(ix, v, bases, offs, names, name_offs, quals, qual_offs, sam, r0, r1, parts[t]);
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  size_t total = 0;
  std::vector<size_t> at(T);
  for (uint32_t t = 0; t < T; t++) { at[t] = total; total += parts[t].s.size(); }
  *out = (char*)malloc(total + 1);
  if (!*out) return tg_fail(TG_ERR_INTERNAL, "out of memory");
  auto copy = [&](uint32_t t) { memcpy(*out + at[t], parts[t].s.data(), parts[t].s.size()); };
  if (T == 1) copy(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; t++) th.emplace_back(copy, t);
    for (auto& x : th) x.join();
  }
  (*out)[total] = 0;
  *out_len = total;
  return TG_OK;
}

// One segment of FASTQ text (starts at a record start): 4-line records, blank lines between records skipped, a
// truncated last record dropped (needletail::parse_fastx_file as src/aligner.rs:51-55 uses it).
struct FastqSeg {
  std::string b, nm, q;
  std::vector<uint64_t> bo, no, qo;  // end offsets of every record inside b / nm / q
  bool bad = false;
};
static void parse_fastq_segment(const char* text, size_t begin, size_t end, FastqSeg& o) {
  size_t p = begin;
  auto next_line = [&](size_t& lb, size_t& le) {  // false at the end of the segment
    if (p >= end) return false;
    const void* nl = memchr(text + p, '\n', end - p);
    const size_t e = nl ? (size_t)((const char*)nl - text) : end;
    lb = p; le = e;
    if (le > lb && text[le - 1] == '\r') le--;
    p = e + 1;
    return true;
  };
  size_t lb[4], le[4];
  for (;;) {
    if (!next_line(lb[0], le[0])) return;
    if (lb[0] == le[0]) continue;  // blank line between records
    bool full = true;
    for (int k = 1; k < 4; k++) full = full && next_line(lb[k], le[k]);
    if (!full) return;             // truncated last record
    if (text[lb[0]] != '@') { o.bad = true; return; }
    o.nm.append(text + lb[0] + 1, le[0] - lb[0] - 1);
    o.b.append(text + lb[1], le[1] - lb[1]);
    o.q.append(text + lb[3], le[3] - lb[3]);
    o.no.push_back(o.nm.size()); o.bo.push_back(o.b.size()); o.qo.push_back(o.q.size());
  }
}
// first record start at or after p: a line that starts with '@' whose second next line starts with '+' (a quality line may
// start with '@', but then the second next line is a sequence line, which never starts with '+')
static size_t fastq_record_start(const char* text, size_t len, size_t p) {
  if (p == 0) return 0;
  const void* nl = memchr(text + p - 1, '\n', len - (p - 1));
  if (!nl) return len;
  size_t c = (size_t)((const char*)nl - text) + 1;
  while (c < len) {
    const void* n1 = memchr(text + c, '\n', len - c);
    if (!n1) return len;
    const size_t l1 = (size_t)((const char*)n1 - text) + 1;
    if (text[c] == '@' && l1 < len) {
      const void* n2 = memchr(text + l1, '\n', len - l1);
      if (!n2) return len;
      const size_t l2 = (size_t)((const char*)n2 - text) + 1;
      if (l2 < len && text[l2] == '+') return c;
    }
    c = l1;
  }
  return len;
}

tg_status tg_parse_fastq(const char* text, size_t len, uint32_t* n_reads, uint8_t** bases, uint64_t** offs, uint8_t** names,
                         uint64_t** name_offs, uint8_t** quals, uint64_t** qual_offs) {
  if (!text || !n_reads || !bases || !offs || !names || !name_offs || !quals || !qual_offs)
    return tg_fail(TG_ERR_INVALID, "null argument");
  // segments of the text on the host's cores, cut at record starts; pieces concatenated in order
  uint32_t T = (uint32_t)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 64);
  if (len < (8u << 20)) T = 1;
  std::vector<size_t> cut(T + 1, len);
  cut[0] = 0;
  for (uint32_t t = 1; t < T; t++) cut[t] = std::max(cut[t - 1], fastq_record_start(text, len, (size_t)((double)len * t / T)));
  std::vector<FastqSeg> seg(T);
  if (T == 1) parse_fastq_segment(text, 0, len, seg[0]);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; t++) th.emplace_back([&, t]() { parse_fastq_segment(text, cut[t], cut[t + 1], seg[t]); });
    for (auto& x : th) x.join();
  }
  size_t nb = 0, nn = 0, nq = 0, nr = 0;
  for (auto& g : seg) {
    if (g.bad) return tg_fail(TG_ERR_IO, "FASTQ record does not start with '@'");
    nb += g.b.size(); nn += g.nm.size(); nq += g.q.size(); nr += g.bo.size();
  }
  if (nr > 0xFFFFFFFFull) return tg_fail(TG_ERR_CAPACITY, "more than 2^32 reads in one FASTQ text");
  *n_reads = (uint32_t)nr;
  *bases = (uint8_t*)malloc(nb ? nb : 1);
  *names = (uint8_t*)malloc(nn ? nn : 1);
  *quals = (uint8_t*)malloc(nq ? nq : 1);
  *offs = (uint64_t*)malloc((nr + 1) * 8);
  *name_offs = (uint64_t*)malloc((nr + 1) * 8);
  *qual_offs = (uint64_t*)malloc((nr + 1) * 8);
  if (!*bases || !*names || !*quals || !*offs || !*name_offs || !*qual_offs) return tg_fail(TG_ERR_INTERNAL, "out of memory");
  (*offs)[0] = 0; (*name_offs)[0] = 0; (*qual_offs)[0] = 0;
  std::vector<size_t> ab(T), an(T), aq(T), ar(T);
  size_t cb = 0, cn = 0, cq = 0, cr = 0;
  for (uint32_t t = 0; t < T; t++) {
    ab[t] = cb; an[t] = cn; aq[t] = cq; ar[t] = cr;
    cb += seg[t].b.size(); cn += seg[t].nm.size(); cq += seg[t].q.size(); cr += seg[t].bo.size();
  }
  auto place = [&](uint32_t t) {
    const FastqSeg& g = seg[t];
    memcpy(*bases + ab[t], g.b.data(), g.b.size());
    memcpy(*names + an[t], g.nm.data(), g.nm.size());
    memcpy(*quals + aq[t], g.q.data(), g.q.size());
    for (size_t i = 0; i < g.bo.size(); i++) {
      (*offs)[ar[t] + i + 1] = ab[t] + g.bo[i];
      (*name_offs)[ar[t] + i + 1] = an[t] + g.no[i];
      (*qual_offs)[ar[t] + i + 1] = aq[t] + g.qo[i];
    }
  };
  if (T == 1) place(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; t++) th.emplace_back(place, t);
    for (auto& x : th) x.join();
  }
  return TG_OK;
}

}  // extern "C"

// ---- FASTQ text in two passes (host_stream.cpp): count, then fill one batch in place from many threads ------------------
// Same record rules as parse_fastq_segment: 4-line records, blank lines between records skipped, '\r' dropped, a truncated
// last record left unconsumed.
namespace {
struct LineWalker {
  const char* text;
  size_t p, end;
  bool final;  // `end` is the end of the file: a last line without '\n' is a line; otherwise it is an incomplete one
  bool next(size_t& lb, size_t& le) {
    if (p >= end) return false;
    const void* nl = memchr(text + p, '\n', end - p);
    if (!nl && !final) return false;
    const size_t e = nl ? (size_t)((const char*)nl - text) : end;
    lb = p; le = e;
    if (le > lb && text[le - 1] == '\r') le--;
    p = e + 1;
    return true;
  }
};
}  // namespace

void tg_fastq_count(const char* text, size_t begin, size_t end, bool final, TgFastqCount& c) {
  LineWalker w{text, begin, end, final};
  c = TgFastqCount();
  c.consumed = begin;
  size_t lb[4], le[4];
  for (;;) {
    const size_t rec_start = w.p;
    if (!w.next(lb[0], le[0])) { c.consumed = rec_start < end ? rec_start : end; return; }  // (a last line without '\n' leaves p at end + 1)
    if (lb[0] == le[0]) { c.consumed = w.p < end ? w.p : end; continue; }  // blank line between records
    bool full = true;
    for (int k = 1; k < 4; k++) full = full && w.next(lb[k], le[k]);
    // (a record is complete only when its quality line is terminated, or the text ends here for good)
    if (!full) { c.consumed = rec_start < end ? rec_start : end; return; }
    if (text[lb[0]] != '@') { c.bad = true; return; }
    c.n++;
    c.names += le[0] - lb[0] - 1; c.bases += le[1] - lb[1]; c.quals += le[3] - lb[3];
    c.consumed = w.p < end ? w.p : end;
  }
}

void tg_fastq_fill(const char* text, size_t begin, size_t end, uint64_t n, uint8_t* bases, uint64_t* offs, uint64_t base0,
                   uint8_t* names, uint64_t* name_offs, uint64_t name0, uint8_t* quals, uint64_t* qual_offs, uint64_t qual0) {
  LineWalker w{text, begin, end, true};
  size_t lb[4], le[4];
  uint64_t b = base0, nm = name0, q = qual0;
  for (uint64_t i = 0; i < n;) {
    if (!w.next(lb[0], le[0])) return;
    if (lb[0] == le[0]) continue;
    for (int k = 1; k < 4; k++) w.next(lb[k], le[k]);
    memcpy(names + nm, text + lb[0] + 1, le[0] - lb[0] - 1); nm += le[0] - lb[0] - 1;
    memcpy(bases + b, text + lb[1], le[1] - lb[1]); b += le[1] - lb[1];
    memcpy(quals + q, text + lb[3], le[3] - lb[3]); q += le[3] - lb[3];
    i++;
    offs[i] = b; name_offs[i] = nm; qual_offs[i] = q;  // offs / name_offs / qual_offs point at the segment's first record
  }
}

size_t tg_fastq_record_start(const char* text, size_t len, size_t p) { return fastq_record_start(text, len, p); }

size_t tg_fastq_skip(const char* text, size_t begin, size_t end, uint64_t k) {
  LineWalker w{text, begin, end, true};
  size_t lb, le, done = begin;
  for (uint64_t i = 0; i < k;) {
    if (!w.next(lb, le)) return done;
    if (lb == le) { done = w.p < end ? w.p : end; continue; }
    for (int q = 1; q < 4; q++)
      if (!w.next(lb, le)) return done;
    i++;
    done = w.p < end ? w.p : end;
  }
  return done;
}
