// host_bam.cpp -- BAM output: OutputFormat::Bam of align_reads_from_file
// (/root/reference/src/aligner.rs:41-47, 69-72, 98-101: noodles `bam::Writer::write_header`, `write_reference_sequences`,
// `write_sam_record`).  The reference builds a SAM record and lets the BAM writer encode it; this file does the same:
// the SAM text of tg_format_batch / tg_format_sam_header is encoded line by line into BAM records (SAM spec v1 section
// 4.2) and written as BGZF blocks (section 4.1; zlib raw deflate, 64 KB blocks, the 28-byte EOF marker on request).
// Parity note: the compressed bytes depend on the deflate implementation and the integer tag width on the encoder
// (the spec allows c/C/s/S/i/I for the same value; the smallest type that holds the value is used, as htslib does), so
// BAM parity is defined on the decoded records: they must decode to exactly the SAM text (tests/test_abi.py).
#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "tg_internal.h"
#include "host_text.h"

namespace {

void put16(std::string& s, uint32_t v) { char b[2] = {(char)(v & 255), (char)((v >> 8) & 255)}; s.append(b, 2); }
void put32(std::string& s, uint32_t v) {
  char b[4] = {(char)(v & 255), (char)((v >> 8) & 255), (char)((v >> 16) & 255), (char)((v >> 24) & 255)};
  s.append(b, 4);
}

// one BGZF block for data[0, n), n <= 0xff00
bool bgzf_block(const char* data, size_t n, std::string& out) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  // level 6 = what the reference's writer uses (noodles / flate2 default); TG_BAM_LEVEL=1..9 trades size for speed (the
  // BAM path is deflate bound: 2.3 M reads/s at level 6 on 16 threads)
  static const int level = []() { const char* e = getenv("TG_BAM_LEVEL"); const int v = e ? atoi(e) : 6; return v >= 1 && v <= 9 ? v : 6; }();
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
  std::vector<unsigned char> buf(deflateBound(&zs, (uLong)n) + 64);
  zs.next_in = (Bytef*)data;
  zs.avail_in = (uInt)n;
  zs.next_out = buf.data();
  zs.avail_out = (uInt)buf.size();
  const int rc = deflate(&zs, Z_FINISH);
  const size_t clen = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END || clen + 26 > 65536) return false;
  static const unsigned char hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
  out.append((const char*)hdr, 12);
  out.push_back('B'); out.push_back('C');
  put16(out, 2);
  put16(out, (uint32_t)(clen + 25));  // BSIZE = total block size - 1
  out.append((const char*)buf.data(), clen);
  put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)data, (uInt)n));
  put32(out, (uint32_t)n);
  return true;
}

// raw bytes -> BGZF blocks (independent blocks: compressed on the host's cores, concatenated in order)
bool bgzf_compress(const std::string& raw, std::string& out) {
  const size_t BLK = 0xff00;
  const size_t nb = (raw.size() + BLK - 1) / BLK;
  if (nb == 0) return true;
  uint32_t T = (uint32_t)std::min<size_t>(std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 64), nb);
  if (nb < 16) T = 1;
  std::vector<std::string> parts(T);
  std::vector<char> ok(T, 1);
  auto work = [&](uint32_t t) {
    const size_t b0 = nb * t / T, b1 = nb * (t + 1) / T;
    for (size_t b = b0; b < b1 && ok[t]; b++) {
      const size_t at = b * BLK;
      if (!bgzf_block(raw.data() + at, std::min(BLK, raw.size() - at), parts[t])) ok[t] = 0;
    }
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < T; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (uint32_t t = 0; t < T; t++) {
    if (!ok[t]) return false;
    out += parts[t];
  }
  return true;
}

const unsigned char BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                    0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// SAM spec 5.3
uint32_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

struct Field { const char* p; size_t n; };

// reference names of the header in @SQ order (build_sam_header collapses the two strands, src/aln_writer.rs:256-276)
std::vector<std::string> header_refs(const tg_index_host* ix, std::vector<uint32_t>* lens) {
  std::vector<std::string> seen;
  for (uint32_t i = 0; i < ix->hdr()->n_refs; i++) {
    const std::string& nm = ix->ref_names[i];
    bool dup = false;
    for (auto& s : seen) if (s == nm) { dup = true; break; }
    if (dup) continue;
    seen.push_back(nm);
    if (lens) {
      uint64_t v[4];
      tg_index_host_ref(ix, i, v);
      lens->push_back((uint32_t)v[2]);
    }
  }
  return seen;
}

// one SAM line (without the newline) -> one BAM record appended to `out`.  No allocation per line: the fields sit in a fixed
// array, the CIGAR words in a vector the caller keeps, the record is built in place and its length patched in at the end.
struct LineScratch {
  std::vector<uint32_t> cig;
  size_t last_ref = 0;
};
inline void set32(std::string& s, size_t at, uint32_t v) {
  s[at] = (char)(v & 255); s[at + 1] = (char)((v >> 8) & 255); s[at + 2] = (char)((v >> 16) & 255); s[at + 3] = (char)((v >> 24) & 255);
}
bool sam_line_to_bam(const char* line, size_t len, const std::vector<std::string>& refs, LineScratch& sc, std::string& out, std::string& err) {
  enum { MAXF = 64 };
  Field f[MAXF];
  size_t nf = 0, b = 0;
  while (nf < MAXF) {
    const char* tab = b < len ? (const char*)memchr(line + b, '\t', len - b) : nullptr;
    const size_t e = tab ? (size_t)(tab - line) : len;
    f[nf++] = Field{line + b, e - b};
    if (!tab) break;
    b = e + 1;
  }
  if (nf == MAXF && memchr(f[MAXF - 1].p, '\t', (size_t)(line + len - f[MAXF - 1].p))) { err = "SAM line with too many fields"; return false; }
  if (nf < 11) { err = "SAM line with fewer than 11 fields"; return false; }
  auto num = [](const Field& x) { long long v = 0; bool neg = false; size_t i = 0; if (x.n && x.p[0] == '-') { neg = true; i = 1; }
                                  for (; i < x.n; i++) v = v * 10 + (x.p[i] - '0'); return neg ? -v : v; };
  int32_t ref_id = -1;
  if (!(f[2].n == 1 && f[2].p[0] == '*')) {
    auto same = [&](size_t i) { return refs[i].size() == f[2].n && memcmp(refs[i].data(), f[2].p, f[2].n) == 0; };
    if (sc.last_ref < refs.size() && same(sc.last_ref)) ref_id = (int32_t)sc.last_ref;  // (neighbouring lines mostly share it)
    else
      for (size_t i = 0; i < refs.size(); i++)
        if (same(i)) { ref_id = (int32_t)i; sc.last_ref = i; break; }
    if (ref_id < 0) { err = "reference name not in the header"; return false; }
  }
  const int64_t pos = num(f[3]) - 1;
  // CIGAR
  std::vector<uint32_t>& cig = sc.cig;
  cig.clear();
  int64_t ref_len = 0;
  if (!(f[5].n == 1 && f[5].p[0] == '*')) {
    uint64_t v = 0;
    for (size_t i = 0; i < f[5].n; i++) {
      const char c = f[5].p[i];
      if (c >= '0' && c <= '9') { v = v * 10 + (uint64_t)(c - '0'); continue; }
      const char* ops = "MIDNSHP=X";
      const char* at = c ? strchr(ops, c) : nullptr;
      if (!at || v >= (1ull << 28)) { err = "bad CIGAR"; return false; }
      const uint32_t op = (uint32_t)(at - ops);
      cig.push_back((uint32_t)(v << 4) | op);
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += (int64_t)v;
      v = 0;
    }
  }
  const bool has_seq = !(f[9].n == 1 && f[9].p[0] == '*');
  const uint32_t l_seq = has_seq ? (uint32_t)f[9].n : 0u;
  const int64_t end = pos + (ref_len > 0 ? ref_len : 1);
  const uint32_t bin = reg2bin(pos, end);  // unmapped (pos -1): reg2bin(-1, 0) = 4680, as the spec asks
  if (f[0].n > 254) { err = "read name longer than 254 bytes"; return false; }
  const bool has_qual = !(f[10].n == 1 && f[10].p[0] == '*');
  if (has_qual && f[10].n != l_seq) { err = "SEQ and QUAL differ in length"; return false; }
  const size_t at0 = out.size();
  // fixed part, name, CIGAR, packed bases, qualities: sizes are known, written through a pointer
  const size_t fixed = 4 + 32 + f[0].n + 1 + 4 * cig.size() + (l_seq + 1) / 2 + l_seq;
  out.resize(at0 + fixed);
  {
    unsigned char* d = (unsigned char*)&out[at0 + 4];
    auto w16 = [&](uint32_t v) { d[0] = (unsigned char)(v & 255); d[1] = (unsigned char)((v >> 8) & 255); d += 2; };
    auto w32 = [&](uint32_t v) { d[0] = (unsigned char)(v & 255); d[1] = (unsigned char)((v >> 8) & 255); d[2] = (unsigned char)((v >> 16) & 255); d[3] = (unsigned char)((v >> 24) & 255); d += 4; };
    w32((uint32_t)ref_id);
    w32((uint32_t)(int32_t)pos);
    *d++ = (unsigned char)(f[0].n + 1);
    *d++ = (unsigned char)num(f[4]);
    w16(bin);
    w16((uint32_t)cig.size());
    w16((uint32_t)num(f[1]));
    w32(l_seq);
    w32(0xFFFFFFFFu);  // RNEXT '*'
    w32(0xFFFFFFFFu);  // PNEXT 0
    w32(0);            // TLEN
    memcpy(d, f[0].p, f[0].n); d += f[0].n;
    *d++ = 0;
    for (uint32_t c : cig) w32(c);
    // "=ACMGRSVTWYHKDBN", case-insensitive; anything else is N (15)
    static const struct Codes {
      unsigned char t[256];
      Codes() {
        memset(t, 15, sizeof(t));
        const char* codes = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; i++) { t[(unsigned char)codes[i]] = (unsigned char)i; if (codes[i] >= 'A' && codes[i] <= 'Z') t[(unsigned char)(codes[i] + 32)] = (unsigned char)i; }
      }
    } C;
    const unsigned char* sq = (const unsigned char*)f[9].p;
    uint32_t i = 0;
    for (; i + 1 < l_seq; i += 2) *d++ = (unsigned char)(C.t[sq[i]] << 4 | C.t[sq[i + 1]]);
    if (i < l_seq) *d++ = (unsigned char)(C.t[sq[i]] << 4);
    if (!has_qual) memset(d, 0xFF, l_seq);
    else for (uint32_t k = 0; k < l_seq; k++) d[k] = (unsigned char)(f[10].p[k] - 33);
  }
  for (size_t k = 11; k < nf; k++) {  // TAG:TYPE:VALUE
    const Field& t = f[k];
    if (t.n < 5 || t.p[2] != ':' || t.p[4] != ':') { err = "bad optional field"; return false; }
    out.append(t.p, 2);
    const Field val{t.p + 5, t.n - 5};
    switch (t.p[3]) {
      case 'A': out.push_back('A'); out.push_back(val.n ? val.p[0] : ' '); break;
      case 'Z': out.push_back('Z'); out.append(val.p, val.n); out.push_back('\0'); break;
      case 'i': {
        const long long v = num(val);
        if (v >= 0) {
          if (v <= 255) { out.push_back('C'); out.push_back((char)v); }
          else if (v <= 65535) { out.push_back('S'); put16(out, (uint32_t)v); }
          else { out.push_back('I'); put32(out, (uint32_t)v); }
        } else {
          if (v >= -128) { out.push_back('c'); out.push_back((char)v); }
          else if (v >= -32768) { out.push_back('s'); put16(out, (uint32_t)(int32_t)v); }
          else { out.push_back('i'); put32(out, (uint32_t)(int32_t)v); }
        }
        break;
      }
      default: err = "optional field type not produced by thermite"; return false;
    }
  }
  set32(out, at0, (uint32_t)(out.size() - at0 - 4));
  return true;
}

tg_status hand_over(const std::string& s, void** out, size_t* out_len) {
  *out = malloc(s.size() + 1);
  if (!*out) return tg_fail(TG_ERR_INTERNAL, "out of memory");
  memcpy(*out, s.data(), s.size());
  *out_len = s.size();
  return TG_OK;
}

}  // namespace

extern "C" {

tg_status tg_format_bam_header(const tg_index_host* ix, void** out, size_t* out_len) {
  if (!ix || !out || !out_len) return tg_fail(TG_ERR_INVALID, "null argument");
  try {
    char* text = nullptr;
    size_t tl = 0;
    tg_status st = tg_format_sam_header(ix, &text, &tl);
    if (st != TG_OK) return st;
    std::string raw("BAM\1", 4);
    put32(raw, (uint32_t)tl);
    raw.append(text, tl);
    free(text);
    std::vector<uint32_t> lens;
    const std::vector<std::string> refs = header_refs(ix, &lens);
    put32(raw, (uint32_t)refs.size());
    for (size_t i = 0; i < refs.size(); i++) {
      put32(raw, (uint32_t)refs[i].size() + 1);
      raw += refs[i];
      raw.push_back('\0');
      put32(raw, lens[i]);
    }
    std::string z;
    if (!bgzf_compress(raw, z)) return tg_fail(TG_ERR_INTERNAL, "deflate failed");
    return hand_over(z, out, out_len);
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

}  // extern "C" (closed for an internal C++ function)

// SAM lines (without header) -> BAM records in BGZF blocks (optionally followed by the end-of-file block).  The text is cut at
// line ends into one piece per thread; a thread encodes its piece and compresses it into blocks of its own (BGZF blocks are
// independent, a short last block per piece is legal), so the only serial step is joining the compressed pieces.
// max_threads = 1: the caller already runs one call per core (tg_align_files).
tg_status tg_sam_text_to_bam(const tg_index_host* ix, const char* sam, size_t sl, bool append_eof, std::string& z, unsigned max_threads) {
    const std::vector<std::string> refs = header_refs(ix, nullptr);
    uint32_t T = (uint32_t)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 64);
    if (max_threads) T = std::min<uint32_t>(T, max_threads);
    if (sl < (1u << 22)) T = 1;
    std::vector<size_t> cut(T + 1, sl);
    cut[0] = 0;
    for (uint32_t t = 1; t < T; t++) {
      size_t p = std::max(cut[t - 1], sl * t / T);
      while (p < sl && sam[p] != '\n') p++;
      cut[t] = p < sl ? p + 1 : sl;
    }
    std::vector<std::string> zparts(T), errs(T);
    auto work = [&](uint32_t t) {
      size_t p = cut[t];
      const size_t e = cut[t + 1];
      std::string raw;
      raw.reserve((e - p) * 3 / 4 + 64);
      LineScratch sc;
      while (p < e) {
        const char* nl = (const char*)memchr(sam + p, '\n', e - p);
        const size_t n = nl ? (size_t)(nl - (sam + p)) : e - p;
        if (n && !sam_line_to_bam(sam + p, n, refs, sc, raw, errs[t])) return;
        p += n + 1;
      }
      std::string& out = T == 1 ? z : zparts[t];
      const size_t BLK = 0xff00;
      out.reserve(out.size() + raw.size() / 2 + 64);
      for (size_t at = 0; at < raw.size(); at += BLK)
        if (!bgzf_block(raw.data() + at, std::min(BLK, raw.size() - at), out)) { errs[t] = "deflate failed"; return; }
    };
    if (T == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (uint32_t t = 0; t < T; t++) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
    for (uint32_t t = 0; t < T; t++)
      if (!errs[t].empty()) return tg_fail(TG_ERR_INTERNAL, "BAM encoding: " + errs[t]);
    if (T > 1) {
      size_t tot = z.size();
      for (uint32_t t = 0; t < T; t++) tot += zparts[t].size();
      z.reserve(tot + 28);
      for (uint32_t t = 0; t < T; t++) z += zparts[t];
    }
    if (append_eof) z.append((const char*)BGZF_EOF, 28);
    return TG_OK;
}

extern "C" {

tg_status tg_format_batch_bam(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                              const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                              int append_eof, void** out, size_t* out_len) {
  if (!out || !out_len) return tg_fail(TG_ERR_INVALID, "null argument");
  try {
    char* sam = nullptr;
    size_t sl = 0;
    tg_status st = tg_format_batch(ix, res, bases, offs, names, name_offs, quals, qual_offs, 1, &sam, &sl);
    if (st != TG_OK) return st;
    std::string z;
    st = tg_sam_text_to_bam(ix, sam, sl, append_eof != 0, z, 0);
    free(sam);
    if (st != TG_OK) return st;
    return hand_over(z, out, out_len);
  } catch (const std::exception& e) {
    return tg_fail(TG_ERR_INTERNAL, e.what());
  }
}

}  // extern "C"
