// tg_textfmt.h -- one read's SAM records as text, written by one thread (device: tg_paf.cu; host build for the CPU tests:
// hosttest.cpp).  The layout restates host_io.cpp's SAM branch (aln_to_sam_record, src/aln_writer.rs:118-238;
// unmapped_sam_record :241-253; to_noodles_cigar :279-323; format_read_name :344-349); the same routine counts a
// read's bytes (WRITE = false) and writes them (WRITE = true), so the two passes cannot disagree.
#pragma once
#include <stdint.h>

#include "tg_core.h"

struct TgTextParams {
  uint32_t n_reads;
  const uint8_t* bases;
  const uint64_t* offs;        // read offsets: L = offs[r + 1] - offs[r]
  const uint64_t* aln_first;   // wide records (tg_aln) of the batch
  const uint32_t* aln_count;
  const tg_aln* alns;
  const uint32_t* ops;
  const uint8_t* names;        // header lines without '@'
  const uint64_t* name_offs;
  const uint8_t* quals;
  const uint64_t* qual_offs;
  const char* ref_names;       // names of refs() back to back; then transcript ids, gene ids, gene names
  const uint32_t* ref_name_offs;
  const char* tx_ids;
  const uint32_t* tx_id_offs;
  const char* gene_ids;
  const uint32_t* gene_id_offs;
  const char* gene_names;
  const uint32_t* gene_name_offs;
  const uint32_t* tx_gene;
  unsigned long long* line_off;  // [n_reads + 1]: bytes of every read's lines, then their exclusive scan
  char* text;
  uint32_t mapq[6];            // multimapq (src/aln_writer.rs:332-340) of 0 .. 5 records
};

template <bool WRITE>
struct TgSink {
  char* d;
  unsigned long long n;
  TG_HD void ch(char c) {
    if (WRITE) *d++ = c; else n++;
  }
  TG_HD void num(unsigned long long v) {
    uint32_t k = 1;
    for (unsigned long long t = v; t >= 10ull; t /= 10ull) k++;
    if (WRITE) {
      for (uint32_t i = k; i-- > 0;) { d[i] = (char)('0' + (uint32_t)(v % 10ull)); v /= 10ull; }
      d += k;
    } else n += k;
  }
  TG_HD void bytes(const char* s, unsigned long long len) {
    if (WRITE) { for (unsigned long long i = 0; i < len; i++) d[i] = s[i]; d += len; } else n += len;
  }
  TG_HD void reversed(const char* s, unsigned long long len) {
    if (WRITE) { for (unsigned long long i = 0; i < len; i++) d[i] = s[len - 1 - i]; d += len; } else n += len;
  }
  TG_HD void revcomp(const char* s, unsigned long long len) {  // bio::alphabets::dna::revcomp keeps case; ACGTN is what reads carry
    if (WRITE) {
      for (unsigned long long i = 0; i < len; i++) {
        char c = s[len - 1 - i];
        switch (c) {
          case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
          case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
          default: break;
        }
        d[i] = c;
      }
      d += len;
    } else n += len;
  }
  template <int N>
  TG_HD void lit(const char (&s)[N]) {
    if (WRITE) { for (int i = 0; i < N - 1; i++) d[i] = s[i]; d += N - 1; } else n += N - 1;
  }
  // to_noodles_cigar: Match and Subst both print as M and merge; equal adjacent clips collapse into one entry
  TG_HD void cigar(const uint32_t* w, uint32_t cnt) {
    uint32_t i = 0;
    while (i < cnt) {
      const uint32_t kind = w[i] & 7u, run = w[i] >> 3;
      if (kind <= TG_OP_SUBST) {
        unsigned long long tot = run;
        while (i + 1 < cnt && (w[i + 1] & 7u) <= TG_OP_SUBST) { tot += w[i + 1] >> 3; i++; }
        num(tot); ch('M');
      } else {
        if (kind >= TG_OP_XCLIP) while (i + 1 < cnt && w[i + 1] == w[i]) i++;
        num(run);
        ch(kind == TG_OP_DEL ? 'D' : kind == TG_OP_INS ? 'I' : kind == TG_OP_XCLIP ? 'S' : 'N');
      }
      i++;
    }
  }
};

// the SAM lines of read r: returns their length; WRITE puts them at p.text + p.line_off[r]
template <bool WRITE>
TG_HD unsigned long long tg_sam_read(const TgTextParams& p, uint32_t r) {
  TgSink<WRITE> o;
  o.n = 0;
  o.d = WRITE ? p.text + p.line_off[r] : nullptr;
  const char* nm = (const char*)p.names + p.name_offs[r];
  unsigned long long qn_len = p.name_offs[r + 1] - p.name_offs[r];
  for (unsigned long long i = 0; i < qn_len; i++) if (nm[i] == ' ') { qn_len = i; break; }  // format_read_name
  const unsigned long long L = p.offs[r + 1] - p.offs[r], QL = p.qual_offs[r + 1] - p.qual_offs[r];
  const char* sq = (const char*)p.bases + p.offs[r];
  const char* ql = (const char*)p.quals + p.qual_offs[r];
  const uint32_t cnt = p.aln_count[r];
  if (cnt == 0) {  // unmapped_sam_record
    o.bytes(nm, qn_len);
    o.lit("\t4\t*\t0\t255\t*\t*\t0\t0\t");
    if (L) o.bytes(sq, L); else o.ch('*');
    o.ch('\t');
    if (QL) o.bytes(ql, QL); else o.ch('*');
    o.ch('\n');
    return WRITE ? 0 : o.n;
  }
  const unsigned long long first = p.aln_first[r];
  const uint32_t mq = p.mapq[cnt < 5u ? cnt : 5u];
  for (uint32_t i = 0; i < cnt; i++) {
    const tg_aln a = p.alns[first + i];
    const uint32_t* w = p.ops + a.ops_off;
    unsigned long long n_mis = 0;
    for (uint32_t k = 0; k < a.ops_len; k++) if ((w[k] & 7u) == TG_OP_SUBST) n_mis += w[k] >> 3;
    o.bytes(nm, qn_len); o.ch('\t');
    o.num((a.strand ? 0u : 0x10u) | (a.primary ? 0u : 0x100u)); o.ch('\t');
    o.bytes(p.ref_names + p.ref_name_offs[a.ref_id], p.ref_name_offs[a.ref_id + 1] - p.ref_name_offs[a.ref_id]); o.ch('\t');
    o.num(a.ystart + 1); o.ch('\t');
    o.num(mq); o.ch('\t');
    o.cigar(w, a.ops_len);
    o.lit("\t*\t0\t0\t");
    if (L == 0) o.ch('*'); else if (a.strand) o.bytes(sq, L); else o.revcomp(sq, L);
    o.ch('\t');
    if (QL == 0) o.ch('*'); else if (a.strand) o.bytes(ql, QL); else o.reversed(ql, QL);
    o.lit("\tAS:i:");
    if (a.score < 0) { o.ch('-'); o.num((unsigned long long)(-(long long)a.score)); } else o.num((unsigned long long)a.score);
    o.lit("\tNH:i:"); o.num(cnt);
    o.lit("\tHI:i:"); o.num(i + 1);
    o.lit("\tnM:i:"); o.num(n_mis);
    if (a.aln_type == TG_ALN_EXONIC) {
      const uint32_t t = a.tx_or_gene_idx, g = p.tx_gene[t];
      o.lit("\tTX:Z:"); o.bytes(p.tx_ids + p.tx_id_offs[t], p.tx_id_offs[t + 1] - p.tx_id_offs[t]);
      o.lit(",+"); o.num(a.tx_ystart); o.ch(',');
      o.cigar(p.ops + a.tx_ops_off, a.tx_ops_len);
      o.lit("\tGX:Z:"); o.bytes(p.gene_ids + p.gene_id_offs[g], p.gene_id_offs[g + 1] - p.gene_id_offs[g]);
      o.lit("\tGN:Z:"); o.bytes(p.gene_names + p.gene_name_offs[g], p.gene_name_offs[g + 1] - p.gene_name_offs[g]);
      o.lit("\tRE:A:E");
    } else if (a.aln_type == TG_ALN_INTRONIC) {
      const uint32_t g = a.tx_or_gene_idx;
      o.lit("\tGX:Z:"); o.bytes(p.gene_ids + p.gene_id_offs[g], p.gene_id_offs[g + 1] - p.gene_id_offs[g]);
      o.lit("\tGN:Z:"); o.bytes(p.gene_names + p.gene_name_offs[g], p.gene_name_offs[g + 1] - p.gene_name_offs[g]);
      o.lit("\tRE:A:N");
    } else {
      o.lit("\tRE:A:I");
    }
    o.ch('\n');
  }
  return WRITE ? 0 : o.n;
}

// host side: the index's names as back-to-back strings with offsets (what TgTextParams points at, on either side)
#include <string>
#include <vector>
struct TgTextTables {
  std::string ref_names, tx_ids, gene_ids, gene_names;
  std::vector<uint32_t> ref_name_offs, tx_id_offs, gene_id_offs, gene_name_offs, tx_gene;
  static void pack(const std::vector<std::string>& v, std::string& s, std::vector<uint32_t>& o) {
    s.clear(); o.assign(1, 0);
    for (const std::string& x : v) { s += x; o.push_back((uint32_t)s.size()); }
  }
  void build(const tg_index_host* ix) {
    pack(ix->ref_names, ref_names, ref_name_offs);
    pack(ix->tx_ids, tx_ids, tx_id_offs);
    pack(ix->gene_ids, gene_ids, gene_id_offs);
    pack(ix->gene_names, gene_names, gene_name_offs);
    tx_gene = ix->tx_gene;
  }
};
