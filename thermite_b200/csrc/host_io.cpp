// host_io.cpp -- host-side ingest and output of libthermite_gpu: FASTQ text -> read batches, and flat
// alignment records -> PAF / SAM text.  Replaces the per-read loop of align_reads_from_file
// (reference src/aligner.rs:51-115) and the record builders of src/aln_writer.rs:47-116, 118-253, 256-358.
// SAM text follows noodles 0.1.0's writer as far as it can be recalled (not vendored: unpinned).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tg_internal.h"

namespace {

struct Out {
  std::string s;
  void num(uint64_t v) {
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(buf[--n]);
  }
  void snum(int64_t v) {
    if (v < 0) { s.push_back('-'); num((uint64_t)(-v)); } else num((uint64_t)v);
  }
};

// multimapq (src/aln_writer.rs:332-340)
unsigned mapq_of(uint64_t n) {
  if (n <= 1) return 255;
  if (n >= 5) return 0;
  return (unsigned)std::lround(-10.0f * std::log10(1.0f - 1.0f / (float)n));
}

// to_noodles_cigar (src/aln_writer.rs:279-323): Match and Subst both print as M and merge
void cigar(Out& o, const uint32_t* w, uint32_t n) {
  static const char sym[6] = {'M', 'M', 'D', 'I', 'S', 'N'};
  uint32_t i = 0;
  while (i < n) {
    uint32_t kind = w[i] & 7u, run = w[i] >> 3;
    if (kind <= TG_OP_SUBST) {
      uint64_t tot = run;
      while (i + 1 < n && (w[i + 1] & 7u) <= TG_OP_SUBST) { tot += w[i + 1] >> 3; i++; }
      o.num(tot);
      o.s.push_back('M');
    } else if (kind >= TG_OP_XCLIP) {
      // equal adjacent clips collapse into one entry carrying a single clip's length (:309-316, :292-294)
      while (i + 1 < n && w[i + 1] == w[i]) i++;
      o.num(run);
      o.s.push_back(sym[kind]);
    } else {
      o.num(run);
      o.s.push_back(sym[kind]);
    }
    i++;
  }
}

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
  Out o;
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

tg_status tg_format_batch(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                          const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                          int sam, char** out, size_t* out_len) {
  if (!ix || !res || !offs || !names || !name_offs || !out || !out_len) return tg_fail(TG_ERR_INVALID, "null argument");
  if (sam && (!quals || !qual_offs || !bases)) return tg_fail(TG_ERR_INVALID, "SAM output needs bases and qualities");
  Out o;
  o.s.reserve((size_t)res->n_reads * (sam ? 400 : 90));
  for (uint32_t r = 0; r < res->n_reads; r++) {
    const char* nm = (const char*)names + name_offs[r];
    size_t nm_len = name_offs[r + 1] - name_offs[r];
    size_t qn_len = nm_len;  // format_read_name: cut at the first space (src/aln_writer.rs:344-349)
    if (sam) for (size_t i = 0; i < nm_len; i++) if (nm[i] == ' ') { qn_len = i; break; }
    const uint64_t L = offs[r + 1] - offs[r];
    const uint32_t cnt = res->read_aln_count[r];
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
      const tg_aln& a = res->alns[res->read_aln_first[r] + i];
      const uint32_t* w = res->ops + a.ops_off;
      const std::string& rname = ix->ref_names[a.ref_id];
      if (!sam) {  // PafEntry (src/aln_writer.rs:47-116)
        uint64_t n_match = 0, n_match_gap = 0;
        for (uint32_t k = 0; k < a.ops_len; k++) {
          uint32_t kind = w[k] & 7u, run = w[k] >> 3;
          if (kind == TG_OP_MATCH) n_match += run;
          if (kind <= TG_OP_INS) n_match_gap += run;
          else if (kind == TG_OP_XCLIP) n_match_gap += 1;
        }
        o.s.append(nm, nm_len); o.s.push_back('\t');
        o.num(L); o.s.push_back('\t');
        o.num(a.xstart); o.s.push_back('\t');
        o.num(a.xend); o.s.push_back('\t');
        o.s.push_back(a.strand ? '+' : '-'); o.s.push_back('\t');
        o.s += rname; o.s.push_back('\t');
        o.num(a.ylen); o.s.push_back('\t');
        o.num(a.ystart); o.s.push_back('\t');
        o.num(a.yend); o.s.push_back('\t');
        o.num(n_match); o.s.push_back('\t');
        o.num(n_match_gap); o.s.push_back('\t');
        o.num(mapq_of(cnt)); o.s += "\t\n";
      } else {  // aln_to_sam_record (src/aln_writer.rs:118-238)
        unsigned flags = 0;
        if (!a.strand) flags |= 0x10;
        if (!a.primary) flags |= 0x100;
        uint64_t n_mis = 0;
        for (uint32_t k = 0; k < a.ops_len; k++) if ((w[k] & 7u) == TG_OP_SUBST) n_mis += w[k] >> 3;
        o.s.append(nm, qn_len); o.s.push_back('\t');
        o.num(flags); o.s.push_back('\t');
        o.s += rname; o.s.push_back('\t');
        o.num(a.ystart + 1); o.s.push_back('\t');
        o.num(mapq_of(cnt)); o.s.push_back('\t');
        cigar(o, w, a.ops_len);
        o.s += "\t*\t0\t0\t";
        const char* sq = (const char*)bases + offs[r];
        if (L == 0) o.s.push_back('*');
        else if (a.strand) o.s.append(sq, L);
        else for (uint64_t k = L; k-- > 0;) o.s.push_back(comp(sq[k]));
        o.s.push_back('\t');
        const char* ql = (const char*)quals + qual_offs[r];
        size_t qn = qual_offs[r + 1] - qual_offs[r];
        if (qn == 0) o.s.push_back('*');
        else if (a.strand) o.s.append(ql, qn);
        else for (size_t k = qn; k-- > 0;) o.s.push_back(ql[k]);
        o.s += "\tAS:i:"; o.snum(a.score);
        o.s += "\tNH:i:"; o.num(cnt);
        o.s += "\tHI:i:"; o.num(i + 1);
        o.s += "\tnM:i:"; o.num(n_mis);
        if (a.aln_type == TG_ALN_EXONIC) {
          uint32_t g = ix->tx_gene[a.tx_or_gene_idx];
          o.s += "\tTX:Z:"; o.s += ix->tx_ids[a.tx_or_gene_idx]; o.s += ",+"; o.num(a.tx_ystart); o.s.push_back(',');
          cigar(o, res->ops + a.tx_ops_off, a.tx_ops_len);
          o.s += "\tGX:Z:"; o.s += ix->gene_ids[g];
          o.s += "\tGN:Z:"; o.s += ix->gene_names[g];
          o.s += "\tRE:A:E";
        } else if (a.aln_type == TG_ALN_INTRONIC) {
          o.s += "\tGX:Z:"; o.s += ix->gene_ids[a.tx_or_gene_idx];
          o.s += "\tGN:Z:"; o.s += ix->gene_names[a.tx_or_gene_idx];
          o.s += "\tRE:A:N";
        } else {
          o.s += "\tRE:A:I";
        }
        o.s.push_back('\n');
      }
    }
  }
  *out = (char*)malloc(o.s.size() + 1);
  if (!*out) return tg_fail(TG_ERR_INTERNAL, "out of memory");
  memcpy(*out, o.s.data(), o.s.size());
  (*out)[o.s.size()] = 0;
  *out_len = o.s.size();
  return TG_OK;
}

tg_status tg_parse_fastq(const char* text, size_t len, uint32_t* n_reads, uint8_t** bases, uint64_t** offs, uint8_t** names,
                         uint64_t** name_offs, uint8_t** quals, uint64_t** qual_offs) {
  if (!text || !n_reads || !bases || !offs || !names || !name_offs || !quals || !qual_offs)
    return tg_fail(TG_ERR_INVALID, "null argument");
  std::vector<std::pair<size_t, size_t>> lines;  // [begin, end) without the line terminator
  size_t p = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && text[e] != '\n') e++;
    size_t ee = e;
    if (ee > p && text[ee - 1] == '\r') ee--;
    lines.push_back({p, ee});
    p = e + 1;
  }
  std::string b, nm, q;
  std::vector<uint64_t> bo(1, 0), no(1, 0), qo(1, 0);
  size_t i = 0;
  while (i < lines.size()) {
    if (lines[i].first == lines[i].second) { i++; continue; }
    if (i + 3 >= lines.size()) break;
    if (text[lines[i].first] != '@') return tg_fail(TG_ERR_IO, "FASTQ record does not start with '@'");
    nm.append(text + lines[i].first + 1, lines[i].second - lines[i].first - 1);
    b.append(text + lines[i + 1].first, lines[i + 1].second - lines[i + 1].first);
    q.append(text + lines[i + 3].first, lines[i + 3].second - lines[i + 3].first);
    no.push_back(nm.size()); bo.push_back(b.size()); qo.push_back(q.size());
    i += 4;
  }
  auto dup = [](const void* src, size_t n) {
    void* d = malloc(n ? n : 1);
    if (n) memcpy(d, src, n);
    return d;
  };
  *n_reads = (uint32_t)(bo.size() - 1);
  *bases = (uint8_t*)dup(b.data(), b.size());
  *names = (uint8_t*)dup(nm.data(), nm.size());
  *quals = (uint8_t*)dup(q.data(), q.size());
  *offs = (uint64_t*)dup(bo.data(), bo.size() * 8);
  *name_offs = (uint64_t*)dup(no.data(), no.size() * 8);
  *qual_offs = (uint64_t*)dup(qo.data(), qo.size() * 8);
  return TG_OK;
}

}  // extern "C"
