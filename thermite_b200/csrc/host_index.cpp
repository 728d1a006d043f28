// host_index.cpp -- host-side index build for libthermite_gpu: FASTA + GTF -> one flat blob
// (4-bit packed both-strand text, full suffix array, flattened interval trees, packed transcripts).
// Replaces Index::create_from_files (reference src/index.rs:52-223); the FM/FMD structures of the
// reference (:103-111) are NOT built -- seeding on the GPU uses a k-mer table over the suffix array.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include "tg_internal.h"

static thread_local std::string g_tg_error;
void tg_set_error(const std::string& msg) { g_tg_error = msg; }
tg_status tg_fail(tg_status code, const std::string& msg) {
  g_tg_error = msg;
  return code;
}
extern "C" const char* tg_last_error(void) { return g_tg_error.c_str(); }
extern "C" const char* tg_version(void) { return "thermite_gpu 0.1 sm_100a"; }
extern "C" void tg_opts_default(tg_opts* o) {  // src/main.rs:115-132
  o->min_seed_len = 20;
  o->min_aln_score_percent = 0.66f;
  o->min_aln_score = 30;
  o->multimap_score_range = 1;
  o->intron_mode = 0;
}
extern "C" void tg_free(void* p) { free(p); }

// ------------------------------------------------------------------------------------------------
// SA-IS (Nong, Zhang & Chan 2009) over an integer string whose last symbol is the unique minimum.
// ------------------------------------------------------------------------------------------------
namespace {

template <typename S>
void sais_rec(const S* s, int32_t* sa, int32_t n, int32_t K) {
  if (n == 1) { sa[0] = 0; return; }
  if (n == 2) { sa[0] = 1; sa[1] = 0; return; }
  std::vector<uint8_t> stype(n);  // 1 = S-type
  stype[n - 1] = 1;
  for (int32_t i = n - 2; i >= 0; i--)
    stype[i] = (s[i] < s[i + 1] || (s[i] == s[i + 1] && stype[i + 1])) ? 1 : 0;
  auto is_lms = [&](int32_t i) { return i > 0 && stype[i] && !stype[i - 1]; };

  std::vector<int32_t> bkt(K + 1, 0);
  for (int32_t i = 0; i < n; i++) bkt[(int32_t)s[i] + 1]++;
  for (int32_t c = 0; c < K; c++) bkt[c + 1] += bkt[c];  // bkt[c] = start, bkt[c+1] = end
  std::vector<int32_t> pos(K);

  auto induce = [&]() {
    for (int32_t c = 0; c < K; c++) pos[c] = bkt[c];
    for (int32_t i = 0; i < n; i++) {
      int32_t j = sa[i] - 1;
      if (sa[i] > 0 && !stype[j]) sa[pos[s[j]]++] = j;
    }
    for (int32_t c = 0; c < K; c++) pos[c] = bkt[c + 1];
    for (int32_t i = n - 1; i >= 0; i--) {
      int32_t j = sa[i] - 1;
      if (sa[i] > 0 && stype[j]) sa[--pos[s[j]]] = j;
    }
  };

  // stage 1: sort LMS substrings
  std::fill(sa, sa + n, -1);
  for (int32_t c = 0; c < K; c++) pos[c] = bkt[c + 1];
  for (int32_t i = n - 1; i >= 1; i--)
    if (is_lms(i)) sa[--pos[s[i]]] = i;
  induce();

  // compact sorted LMS positions
  int32_t n1 = 0;
  for (int32_t i = 0; i < n; i++)
    if (is_lms(sa[i])) sa[n1++] = sa[i];
  // name them (names stored at sa[n1 + pos/2])
  std::fill(sa + n1, sa + n, -1);
  int32_t name = 0, prev = -1;
  for (int32_t i = 0; i < n1; i++) {
    int32_t p = sa[i];
    bool diff = false;
    if (prev < 0) diff = true;
    else {
      for (int32_t d = 0;; d++) {
        if (s[p + d] != s[prev + d] || stype[p + d] != stype[prev + d]) { diff = true; break; }
        if (d > 0 && (is_lms(p + d) || is_lms(prev + d))) break;
      }
    }
    if (diff) { name++; prev = p; }
    sa[n1 + p / 2] = name - 1;
  }
  std::vector<int32_t> s1(n1), sa1(n1);
  {
    int32_t j = 0;
    for (int32_t i = n1; i < n; i++)
      if (sa[i] >= 0) s1[j++] = sa[i];
  }
  if (name < n1) sais_rec<int32_t>(s1.data(), sa1.data(), n1, name);
  else
    for (int32_t i = 0; i < n1; i++) sa1[s1[i]] = i;

  // stage 3: place LMS suffixes in their final relative order and induce
  {
    int32_t j = 0;
    for (int32_t i = 1; i < n; i++)
      if (is_lms(i)) s1[j++] = i;  // s1 now maps reduced index -> text position
  }
  for (int32_t i = 0; i < n1; i++) sa1[i] = s1[sa1[i]];
  std::fill(sa, sa + n, -1);
  for (int32_t c = 0; c < K; c++) pos[c] = bkt[c + 1];
  for (int32_t i = n1 - 1; i >= 0; i--) {
    int32_t p = sa1[i];
    sa[--pos[s[p]]] = p;
  }
  induce();
}

}  // namespace

void tg_sais(const uint8_t* text, size_t n, int32_t* sa) {
  // map bytes to dense codes >= 1 and append a unique 0 sentinel so that a suffix which is a proper
  // prefix of another sorts first (plain lexicographic order, what divsufsort64 gives the reference).
  int code[256];
  std::fill(code, code + 256, 0);
  bool seen[256] = {false};
  for (size_t i = 0; i < n; i++) seen[text[i]] = true;
  int K = 1;
  for (int c = 0; c < 256; c++)
    if (seen[c]) code[c] = K++;
  std::vector<uint8_t> s(n + 1);
  for (size_t i = 0; i < n; i++) s[i] = (uint8_t)code[text[i]];
  s[n] = 0;
  std::vector<int32_t> tmp(n + 1);
  sais_rec<uint8_t>(s.data(), tmp.data(), (int32_t)(n + 1), K);
  std::memcpy(sa, tmp.data() + 1, n * sizeof(int32_t));
}

// ------------------------------------------------------------------------------------------------
// parsing
// ------------------------------------------------------------------------------------------------
namespace {

struct Chrom {
  std::string name;
  std::string seq;  // upper-cased
};

inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

inline int base_code(char c) {
  switch (c) {
    case '$': return TG_C_SENT;
    case 'A': return TG_C_A;
    case 'C': return TG_C_C;
    case 'G': return TG_C_G;
    case 'N': return TG_C_N;
    case 'T': return TG_C_T;
    default: return -1;
  }
}
inline char comp(char c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return c;
  }
}

// needletail semantics (src/index.rs:58-75): record id = header line, name = first space-delimited token
std::vector<Chrom> read_fasta(const char* text, size_t len) {
  std::vector<Chrom> out;
  size_t p = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && text[e] != '\n') e++;
    size_t ee = e;
    if (ee > p && text[ee - 1] == '\r') ee--;
    if (ee > p && text[p] == '>') {
      size_t sp = p + 1;
      while (sp < ee && text[sp] != ' ') sp++;
      out.push_back(Chrom{std::string(text + p + 1, sp - p - 1), std::string()});
    } else if (!out.empty()) {
      std::string& s = out.back().seq;
      size_t base = s.size();
      s.resize(base + (ee - p));
      for (size_t i = p; i < ee; i++) s[base + i - p] = up(text[i]);
    }
    p = e + 1;
  }
  return out;
}

struct GTx {
  std::string id, chrom, gene_id;
  bool fwd;
  std::vector<std::pair<uint64_t, uint64_t>> exons;  // 0-based half open
};
struct GGene {
  std::string id, name;
};

// value of `key "value";` inside a GTF attribute column ("" when absent)
std::string attr_value(const char* a, size_t n, const char* key) {
  size_t klen = strlen(key), i = 0;
  while (i < n) {
    while (i < n && (a[i] == ' ' || a[i] == ';' || a[i] == '\t')) i++;
    size_t ks = i;
    while (i < n && a[i] != ' ') i++;
    size_t ke = i;
    while (i < n && a[i] == ' ') i++;
    std::string val;
    if (i < n && a[i] == '"') {
      size_t vs = ++i;
      while (i < n && a[i] != '"') i++;
      val.assign(a + vs, i - vs);
      if (i < n) i++;
    } else {
      size_t vs = i;
      while (i < n && a[i] != ';') i++;
      val.assign(a + vs, i - vs);
    }
    if (ke - ks == klen && memcmp(a + ks, key, klen) == 0) return val;
  }
  return std::string();
}

// cellranger `transcriptome` crate behaviour the reference relies on (src/index.rs:116-204): genes and
// transcripts are numbered in file order, exons are 0-based half-open and ascending, gene name defaults
// to the gene id.
void read_gtf(const char* text, size_t len, std::vector<GGene>& genes, std::vector<GTx>& txs) {
  std::unordered_map<std::string, uint32_t> gene_of, tx_of;
  size_t p = 0;
  while (p < len) {
    size_t e = p;
    while (e < len && text[e] != '\n') e++;
    size_t ee = e;
    if (ee > p && text[ee - 1] == '\r') ee--;
    if (ee > p && text[p] != '#') {
      size_t col[10];
      int nc = 0;
      col[nc++] = p;
      for (size_t i = p; i < ee && nc < 9; i++)
        if (text[i] == '\t') col[nc++] = i + 1;
      if (nc == 9) {
        col[9] = ee + 1;
        auto field = [&](int k) { return std::string(text + col[k], col[k + 1] - 1 - col[k]); };
        std::string feat = field(2);
        const char* at = text + col[8];
        size_t an = ee - col[8];
        bool is_gene = feat == "gene", is_tx = feat == "transcript", is_exon = feat == "exon";
        if (is_gene || is_tx || is_exon) {
          std::string gid = attr_value(at, an, "gene_id");
          if (!gene_of.count(gid)) {
            std::string gname = attr_value(at, an, "gene_name");
            gene_of[gid] = (uint32_t)genes.size();
            genes.push_back(GGene{gid, gname.empty() ? gid : gname});
          }
          if (is_tx || is_exon) {
            std::string tid = attr_value(at, an, "transcript_id");
            auto it = tx_of.find(tid);
            uint32_t t;
            if (it == tx_of.end()) {
              t = (uint32_t)txs.size();
              tx_of[tid] = t;
              txs.push_back(GTx{tid, field(0), gid, field(6) != "-", {}});
            } else t = it->second;
            if (is_exon) txs[t].exons.push_back({std::stoull(field(3)) - 1, std::stoull(field(4))});
          }
        }
      }
    }
    p = e + 1;
  }
  for (auto& t : txs) std::sort(t.exons.begin(), t.exons.end());
}

// ------------------------------------------------------------------------------------------------
// array-based AVL interval tree with rust-bio's insertion rule (ties descend left) -- only its SHAPE
// matters: the device replays `find` (node, then right subtree, then left subtree) on these nodes.
// ------------------------------------------------------------------------------------------------
struct Avl {
  std::vector<TgTreeNode> nodes;
  std::vector<int32_t> height;
  int32_t root = -1;
  int32_t ht(int32_t i) const { return i < 0 ? 0 : height[i]; }
  void fix(int32_t i) {
    TgTreeNode& n = nodes[i];
    height[i] = 1 + std::max(ht(n.left), ht(n.right));
    n.max = n.end;
    if (n.left >= 0) n.max = std::max(n.max, nodes[n.left].max);
    if (n.right >= 0) n.max = std::max(n.max, nodes[n.right].max);
  }
  int32_t rot_left(int32_t i) {
    int32_t r = nodes[i].right;
    nodes[i].right = nodes[r].left;
    nodes[r].left = i;
    fix(i);
    fix(r);
    return r;
  }
  int32_t rot_right(int32_t i) {
    int32_t l = nodes[i].left;
    nodes[i].left = nodes[l].right;
    nodes[l].right = i;
    fix(i);
    fix(l);
    return l;
  }
  int32_t rebalance(int32_t i) {
    int32_t bal = ht(nodes[i].left) - ht(nodes[i].right);
    if (bal >= -1 && bal <= 1) { fix(i); return i; }
    if (bal < -1) {  // right heavy
      int32_t r = nodes[i].right;
      if (ht(nodes[r].left) > ht(nodes[r].right)) nodes[i].right = rot_right(r);
      return rot_left(i);
    }
    int32_t l = nodes[i].left;
    if (ht(nodes[l].right) > ht(nodes[l].left)) nodes[i].left = rot_left(l);
    return rot_right(i);
  }
  int32_t ins(int32_t at, int32_t fresh) {
    if (at < 0) return fresh;
    if (nodes[fresh].start <= nodes[at].start) nodes[at].left = ins(nodes[at].left, fresh);
    else nodes[at].right = ins(nodes[at].right, fresh);
    return rebalance(at);
  }
  void insert(uint32_t s, uint32_t e, uint32_t data) {
    nodes.push_back(TgTreeNode{s, e, e, data, -1, -1, 0u, 0u});
    height.push_back(1);
    root = ins(root, (int32_t)nodes.size() - 1);
  }
};

// sorted-by-start interval list with the tree's find() rank
std::vector<TgStab> stab_list(const Avl& t, uint64_t& maxlen) {
  std::vector<TgStab> out(t.nodes.size());
  maxlen = 0;
  std::vector<int32_t> stack;
  if (t.root >= 0) stack.push_back(t.root);
  uint32_t rank = 0;
  while (!stack.empty()) {  // node, then right subtree, then left subtree
    int32_t c = stack.back();
    stack.pop_back();
    const TgTreeNode& n = t.nodes[c];
    if (n.left >= 0) stack.push_back(n.left);
    if (n.right >= 0) stack.push_back(n.right);
    out[rank] = TgStab{n.start, n.end, n.data, rank};
    maxlen = std::max<uint64_t>(maxlen, (uint64_t)n.end - n.start);
    rank++;
  }
  std::stable_sort(out.begin(), out.end(), [](const TgStab& a, const TgStab& b) { return a.start < b.start; });
  return out;
}

struct BlobWriter {
  std::vector<uint8_t> buf;
  uint64_t reserve_section(size_t nbytes) {
    size_t off = (buf.size() + 255) & ~(size_t)255;
    buf.resize(off + nbytes, 0);
    return off;
  }
  template <typename T>
  uint64_t put(const std::vector<T>& v) {
    uint64_t off = reserve_section(v.size() * sizeof(T));
    if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
  uint64_t put_strings(const std::vector<std::string>& v) {
    std::vector<uint64_t> offs(v.size() + 1, 0);
    for (size_t i = 0; i < v.size(); i++) offs[i + 1] = offs[i] + v[i].size();
    uint64_t off = reserve_section(offs.size() * 8 + offs.back());
    memcpy(buf.data() + off, offs.data(), offs.size() * 8);
    uint8_t* p = buf.data() + off + offs.size() * 8;
    for (size_t i = 0; i < v.size(); i++) memcpy(p + offs[i], v[i].data(), v[i].size());
    return off;
  }
};

// 16 symbols per u64, first symbol in the most significant nibble
void pack4(const std::string& s, int (*code)(char), std::vector<uint64_t>& out, uint64_t base_pos) {
  for (size_t i = 0; i < s.size(); i++) {
    uint64_t p = base_pos + i;
    out[p >> 4] |= (uint64_t)code(s[i]) << ((15 - (p & 15)) * 4);
  }
}

void decode_strings(const uint8_t* blob, uint64_t off, size_t n, std::vector<std::string>& out) {
  const uint64_t* offs = (const uint64_t*)(blob + off);
  const char* p = (const char*)(blob + off + (n + 1) * 8);
  out.resize(n);
  for (size_t i = 0; i < n; i++) out[i].assign(p + offs[i], offs[i + 1] - offs[i]);
}

tg_status finish_host_index(tg_index_host* ix) {
  if (const char* why = tg_blob_validate_host(ix->blob.data(), ix->blob.size()))
    return tg_fail(TG_ERR_INVALID, std::string("not a usable thermite_gpu index blob: ") + why);
  const TgBlobHeader* h = ix->hdr();
  const uint8_t* b = ix->blob.data();
  decode_strings(b, h->off_ref_names, h->n_refs, ix->ref_names);
  decode_strings(b, h->off_tx_ids, h->n_txs, ix->tx_ids);
  decode_strings(b, h->off_gene_ids, h->n_genes, ix->gene_ids);
  decode_strings(b, h->off_gene_names, h->n_genes, ix->gene_names);
  ix->tx_gene.assign((const uint32_t*)(b + h->off_tx_gene), (const uint32_t*)(b + h->off_tx_gene) + h->n_txs);
  ix->tx_strand.assign((const uint32_t*)(b + h->off_tx_strand), (const uint32_t*)(b + h->off_tx_strand) + h->n_txs);
  return TG_OK;
}

// sa_device < 0: suffix array by SA-IS on the host; >= 0: on that GPU (csrc/tg_sa.cu)
tg_status build_index(const char* fasta, size_t fasta_len, const char* gtf, size_t gtf_len, int sa_device,
                      tg_index_host** out) {
  if (sa_device >= 0 && !g_tg_sa_device)
    return tg_fail(TG_ERR_CUDA, "this build has no GPU suffix-array builder (host test library)");
  std::vector<Chrom> chroms = read_fasta(fasta, fasta_len);
  if (chroms.empty()) return tg_fail(TG_ERR_IO, "no sequences in FASTA");
  // concatenated text: fwd $ revcomp $ per chromosome (src/index.rs:66-101)
  std::string text;
  std::vector<TgRef> refs;
  std::vector<std::string> ref_names;
  std::unordered_map<std::string, uint32_t> fwd_ref;
  {
    uint64_t total = 0;
    for (auto& c : chroms) total += 2 * (c.seq.size() + 1);
    if (total >= (1ull << 31) - 4096) return tg_fail(TG_ERR_CAPACITY, "concatenated text must be < 2^31 symbols");
    text.reserve(total);
  }
  for (auto& c : chroms) {
    // IUPAC ambiguity codes (real assemblies hold a few) are indexed as N: the reference's FM alphabet is "ACGNT"
    // (src/index.rs:108) and has no rank for them, so one such base would make the contig unusable there.  A documented
    // deviation; anything that is not a nucleotide code at all is still refused.
    for (char& ch : c.seq) {
      if (base_code(ch) >= 1) continue;
      if (strchr("RYSWKMBDHVU", ch)) ch = 'N';
      else return tg_fail(TG_ERR_IO, "reference sequence contains a symbol that is not a nucleotide code");
    }
    uint32_t s0 = (uint32_t)text.size();
    text += c.seq;
    text.push_back('$');
    fwd_ref[c.name] = (uint32_t)refs.size();
    refs.push_back(TgRef{s0, (uint32_t)text.size(), (uint32_t)c.seq.size(), 1u});
    ref_names.push_back(c.name);
    uint32_t s1 = (uint32_t)text.size();
    for (size_t i = c.seq.size(); i-- > 0;) text.push_back(comp(c.seq[i]));
    text.push_back('$');
    refs.push_back(TgRef{s1, (uint32_t)text.size(), (uint32_t)c.seq.size(), 0u});
    ref_names.push_back(c.name);
  }
  {  // name rank in byte order (String::cmp in filter_overlapping, src/aligner.rs:322-327)
    std::vector<std::string> sorted(ref_names);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    for (size_t i = 0; i < refs.size(); i++) {
      uint32_t rank = (uint32_t)(std::lower_bound(sorted.begin(), sorted.end(), ref_names[i]) - sorted.begin());
      refs[i].strand_rank |= rank << 1;
    }
  }
  const uint64_t T = text.size();

  // transcriptome (src/index.rs:115-220)
  std::vector<GGene> genes;
  std::vector<GTx> gtxs;
  read_gtf(gtf, gtf_len, genes, gtxs);
  Avl exon_tree, gene_tree;
  std::vector<uint64_t> tx_seq_off(1, 0);
  std::vector<uint32_t> tx_exon_off(1, 0), te_start, te_end, tx_gene, tx_strand;
  std::vector<std::string> tx_ids, gene_ids, gene_names;
  std::vector<std::string> tx_seqs;
  std::unordered_map<std::string, uint32_t> gene_of;
  for (size_t g = 0; g < genes.size(); g++) {
    gene_of[genes[g].id] = (uint32_t)g;
    gene_ids.push_back(genes[g].id);
    gene_names.push_back(genes[g].name);
  }
  std::vector<std::pair<uint64_t, uint64_t>> gene_span(genes.size(), {T, 0});  // :134
  for (size_t t = 0; t < gtxs.size(); t++) {
    GTx& tx = gtxs[t];
    if (tx.exons.empty()) return tg_fail(TG_ERR_IO, "transcript without exons: " + tx.id);
    auto fr = fwd_ref.find(tx.chrom);
    if (fr == fwd_ref.end()) return tg_fail(TG_ERR_IO, "GTF names a sequence that is not in the FASTA: " + tx.chrom);
    const Chrom& ch = chroms[fr->second / 2];
    const TgRef& r = refs[fr->second + (tx.fwd ? 0 : 1)];  // the transcript's own strand copy
    std::string seq;
    for (auto& e : tx.exons) {
      if (e.second > ch.seq.size() || e.first >= e.second) return tg_fail(TG_ERR_IO, "exon outside its sequence: " + tx.id);
      seq.append(ch.seq, e.first, e.second - e.first);
    }
    if (!tx.fwd) {
      std::string rc(seq.size(), 'N');
      for (size_t i = 0; i < seq.size(); i++) rc[i] = comp(seq[seq.size() - 1 - i]);
      seq.swap(rc);
    }
    auto to_concat = [&](uint64_t a, uint64_t b, uint64_t& s, uint64_t& e) {  // :149-179
      if (tx.fwd) { s = a + r.start_idx; e = b + r.start_idx; }
      else { s = (uint64_t)r.end_idx - 1 - b; e = (uint64_t)r.end_idx - 1 - a; }
    };
    uint64_t ts, te;
    to_concat(tx.exons.front().first, tx.exons.back().second, ts, te);
    uint32_t g = gene_of.at(tx.gene_id);
    gene_span[g].first = std::min(gene_span[g].first, ts);
    gene_span[g].second = std::max(gene_span[g].second, te);
    std::vector<std::pair<uint32_t, uint32_t>> ex;
    for (auto& e : tx.exons) {
      uint64_t s, en;
      to_concat(e.first, e.second, s, en);
      exon_tree.insert((uint32_t)s, (uint32_t)en, (uint32_t)t);  // GTF-ascending order (:182-183)
      ex.push_back({(uint32_t)s, (uint32_t)en});
    }
    if (!tx.fwd) std::reverse(ex.begin(), ex.end());  // :192-195
    for (auto& e : ex) { te_start.push_back(e.first); te_end.push_back(e.second); }
    tx_exon_off.push_back((uint32_t)te_start.size());
    tx_seq_off.push_back(tx_seq_off.back() + seq.size());
    tx_seqs.push_back(std::move(seq));
    tx_ids.push_back(tx.id);
    tx_gene.push_back(g);
    tx_strand.push_back(tx.fwd ? 1u : 0u);
  }
  for (size_t g = 0; g < genes.size(); g++)  // :208-213
    gene_tree.insert((uint32_t)gene_span[g].first, (uint32_t)gene_span[g].second, (uint32_t)g);

  // pack
  std::vector<uint64_t> text4(T / 16 + 4, 0), txseq4(tx_seq_off.back() / 16 + 4, 0);
  pack4(text, [](char c) { return base_code(c); }, text4, 0);

  // suffix array (divsufsort64 in the reference, src/index.rs:103-105)
  std::vector<uint32_t> sa(T);
  if (sa_device >= 0) {
    tg_status st = g_tg_sa_device(text4.data(), T, sa_device, sa.data(), nullptr, nullptr);
    if (st != TG_OK) return st;
  } else {
    tg_sais((const uint8_t*)text.data(), T, (int32_t*)sa.data());
  }
  for (size_t t = 0; t < tx_seqs.size(); t++) pack4(tx_seqs[t], [](char c) { return base_code(c); }, txseq4, tx_seq_off[t]);

  BlobWriter w;
  w.reserve_section(sizeof(TgBlobHeader));
  TgBlobHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = TG_BLOB_MAGIC;
  h.format_version = TG_BLOB_VERSION;
  h.text_len = T;
  h.n_refs = refs.size(); h.n_txs = gtxs.size(); h.n_genes = genes.size();
  h.n_exon_nodes = exon_tree.nodes.size(); h.n_gene_nodes = gene_tree.nodes.size();
  h.n_tx_exons = te_start.size(); h.txseq_len = tx_seq_off.back();
  h.exon_root = exon_tree.root; h.gene_root = gene_tree.root;
  h.off_text4 = w.put(text4);
  h.off_sa = w.put(sa);
  h.off_refs = w.put(refs);
  h.off_exon_nodes = w.put(exon_tree.nodes);
  h.off_gene_nodes = w.put(gene_tree.nodes);
  h.off_tx_seq_off = w.put(tx_seq_off);
  h.off_tx_exon_off = w.put(tx_exon_off);
  h.off_te_start = w.put(te_start);
  h.off_te_end = w.put(te_end);
  h.off_txseq4 = w.put(txseq4);
  h.off_exon_stab = w.put(stab_list(exon_tree, h.exon_maxlen));
  h.off_gene_stab = w.put(stab_list(gene_tree, h.gene_maxlen));
  h.device_bytes = (w.buf.size() + 255) & ~(uint64_t)255;
  h.off_ref_names = w.put_strings(ref_names);
  h.off_tx_ids = w.put_strings(tx_ids);
  h.off_gene_ids = w.put_strings(gene_ids);
  h.off_gene_names = w.put_strings(gene_names);
  h.off_tx_gene = w.put(tx_gene);
  h.off_tx_strand = w.put(tx_strand);
  w.reserve_section(0);
  h.nbytes = w.buf.size();
  memcpy(w.buf.data(), &h, sizeof(h));

  auto* ix = new tg_index_host();
  ix->blob.swap(w.buf);
  tg_status st = finish_host_index(ix);
  if (st != TG_OK) { delete ix; return st; }
  *out = ix;
  return TG_OK;
}

bool slurp(const char* path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

}  // namespace

tg_sa_device_fn g_tg_sa_device = nullptr;  // set by the static initialiser of tg_sa.cu in libthermite_gpu.so

const char* tg_blob_check_header(const TgBlobHeader& h, uint64_t avail, bool device_prefix_only) {
  if (avail < sizeof(TgBlobHeader)) return "shorter than its header";
  if (h.magic != TG_BLOB_MAGIC) return "wrong magic";
  if (h.format_version != TG_BLOB_VERSION) return "written by another format version (re-create the index)";
  if (h.device_bytes > h.nbytes || h.device_bytes < sizeof(TgBlobHeader)) return "device prefix larger than the blob";
  if (device_prefix_only ? h.device_bytes > avail : h.nbytes != avail) return "size does not match the header (truncated?)";
  if (h.text_len == 0 || h.text_len >= (1ull << 31)) return "text length out of range";
  if (h.n_refs == 0 || h.n_refs > (1ull << 28) || h.n_txs > (1ull << 28) || h.n_genes > (1ull << 28) ||
      h.n_exon_nodes > (1ull << 30) || h.n_gene_nodes > (1ull << 30) || h.n_tx_exons > (1ull << 30) || h.txseq_len >= (1ull << 40))
    return "counts out of range";
  // section = offset + count * element size, inside [header, limit), 8-byte aligned
  auto inside = [&](uint64_t off, uint64_t count, uint64_t elem, uint64_t limit) {
    return off >= sizeof(TgBlobHeader) && (off & 7) == 0 && off <= limit && count <= (limit - off) / elem;
  };
  const uint64_t dl = h.device_bytes;
  if (!inside(h.off_text4, h.text_len / 16 + 4, 8, dl)) return "text section out of bounds";
  if (!inside(h.off_sa, h.text_len, 4, dl)) return "suffix array out of bounds";
  if (!inside(h.off_refs, h.n_refs, sizeof(TgRef), dl)) return "refs out of bounds";
  if (!inside(h.off_exon_nodes, h.n_exon_nodes, sizeof(TgTreeNode), dl) || !inside(h.off_gene_nodes, h.n_gene_nodes, sizeof(TgTreeNode), dl))
    return "interval tree nodes out of bounds";
  if (!inside(h.off_exon_stab, h.n_exon_nodes, sizeof(TgStab), dl) || !inside(h.off_gene_stab, h.n_gene_nodes, sizeof(TgStab), dl))
    return "stab lists out of bounds";
  if (!inside(h.off_tx_seq_off, h.n_txs + 1, 8, dl) || !inside(h.off_tx_exon_off, h.n_txs + 1, 4, dl)) return "transcript tables out of bounds";
  if (!inside(h.off_te_start, h.n_tx_exons, 4, dl) || !inside(h.off_te_end, h.n_tx_exons, 4, dl)) return "exon tables out of bounds";
  if (!inside(h.off_txseq4, h.txseq_len / 16 + 4, 8, dl)) return "transcript sequences out of bounds";
  if (h.exon_root < -1 || h.exon_root >= (int64_t)h.n_exon_nodes || h.gene_root < -1 || h.gene_root >= (int64_t)h.n_gene_nodes)
    return "tree root out of range";
  if (!device_prefix_only) {
    const uint64_t hl = h.nbytes;
    if (!inside(h.off_ref_names, h.n_refs + 1, 8, hl) || !inside(h.off_tx_ids, h.n_txs + 1, 8, hl) ||
        !inside(h.off_gene_ids, h.n_genes + 1, 8, hl) || !inside(h.off_gene_names, h.n_genes + 1, 8, hl))
      return "string tables out of bounds";
    if (!inside(h.off_tx_gene, h.n_txs, 4, hl) || !inside(h.off_tx_strand, h.n_txs, 4, hl)) return "transcript metadata out of bounds";
  }
  return nullptr;
}

const char* tg_blob_validate_host(const uint8_t* b, uint64_t nbytes) {
  if (nbytes < sizeof(TgBlobHeader)) return "shorter than its header";
  TgBlobHeader h;
  memcpy(&h, b, sizeof(h));
  if (const char* why = tg_blob_check_header(h, nbytes, false)) return why;
  // string tables: offsets start at 0, never decrease, bytes inside the blob
  auto strings_ok = [&](uint64_t off, uint64_t n) {
    const uint64_t* o = (const uint64_t*)(b + off);
    const uint64_t room = h.nbytes - (off + (n + 1) * 8);
    if (o[0] != 0) return false;
    for (uint64_t i = 0; i < n; i++)
      if (o[i + 1] < o[i]) return false;
    return o[n] <= room;
  };
  if (!strings_ok(h.off_ref_names, h.n_refs) || !strings_ok(h.off_tx_ids, h.n_txs) || !strings_ok(h.off_gene_ids, h.n_genes) ||
      !strings_ok(h.off_gene_names, h.n_genes))
    return "corrupt string table";
  // refs: consecutive pieces of the text
  const TgRef* refs = (const TgRef*)(b + h.off_refs);
  uint64_t pos = 0;
  for (uint64_t i = 0; i < h.n_refs; i++) {
    if (refs[i].start_idx != pos || refs[i].end_idx != refs[i].start_idx + (uint64_t)refs[i].len + 1) return "corrupt refs table";
    pos = refs[i].end_idx;
  }
  if (pos != h.text_len) return "refs do not cover the text";
  // transcripts: offset tables monotone and ending at the section sizes; exons inside the text; gene indices in range
  const uint64_t* tso = (const uint64_t*)(b + h.off_tx_seq_off);
  const uint32_t* teo = (const uint32_t*)(b + h.off_tx_exon_off);
  const uint32_t* tg = (const uint32_t*)(b + h.off_tx_gene);
  if (tso[0] != 0 || teo[0] != 0 || tso[h.n_txs] != h.txseq_len || teo[h.n_txs] != h.n_tx_exons) return "corrupt transcript offsets";
  for (uint64_t t = 0; t < h.n_txs; t++)
    if (tso[t + 1] < tso[t] || teo[t + 1] < teo[t] || tg[t] >= h.n_genes) return "corrupt transcript table";
  const uint32_t* es = (const uint32_t*)(b + h.off_te_start);
  const uint32_t* ee = (const uint32_t*)(b + h.off_te_end);
  for (uint64_t e = 0; e < h.n_tx_exons; e++)
    if (es[e] > ee[e] || ee[e] > h.text_len) return "corrupt exon table";
  // suffix array entries are text positions (a permutation is not re-checked: that would cost a pass per load)
  const uint32_t* sa = (const uint32_t*)(b + h.off_sa);
  uint32_t bad = 0;
  for (uint64_t i = 0; i < h.text_len; i++) bad |= sa[i] >= h.text_len;
  if (bad) return "suffix array entry outside the text";
  // stab lists: data = transcript / gene index
  const TgStab* xs = (const TgStab*)(b + h.off_exon_stab);
  for (uint64_t i = 0; i < h.n_exon_nodes; i++)
    if (xs[i].data >= h.n_txs || xs[i].start > xs[i].end) return "corrupt exon stab list";
  const TgStab* gs = (const TgStab*)(b + h.off_gene_stab);
  for (uint64_t i = 0; i < h.n_gene_nodes; i++)
    if (gs[i].data >= h.n_genes || gs[i].start > gs[i].end) return "corrupt gene stab list";
  return nullptr;
}


extern "C" {

tg_status tg_index_host_create_from_memory(const char* fasta_text, size_t fasta_len, const char* gtf_text,
                                           size_t gtf_len, tg_index_host** out) {
  if (!fasta_text || !out || (!gtf_text && gtf_len)) return tg_fail(TG_ERR_INVALID, "null argument");
  TG_GUARD_BEGIN
  return build_index(fasta_text, fasta_len, gtf_text ? gtf_text : "", gtf_len, -1, out);
  TG_GUARD_END
}

tg_status tg_index_host_create_from_memory_gpu(const char* fasta_text, size_t fasta_len, const char* gtf_text,
                                               size_t gtf_len, int device, tg_index_host** out) {
  if (!fasta_text || !out || device < 0) return tg_fail(TG_ERR_INVALID, "null argument or negative device");
  TG_GUARD_BEGIN
  return build_index(fasta_text, fasta_len, gtf_text ? gtf_text : "", gtf_len, device, out);
  TG_GUARD_END
}

tg_status tg_index_host_create_from_files(const char* fasta_path, const char* gtf_path, tg_index_host** out) {
  if (!fasta_path || !gtf_path || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  TG_GUARD_BEGIN
  std::string fa, gtf;
  if (!slurp(fasta_path, fa)) return tg_fail(TG_ERR_IO, std::string("cannot read ") + fasta_path);
  if (!slurp(gtf_path, gtf)) return tg_fail(TG_ERR_IO, std::string("cannot read ") + gtf_path);
  return build_index(fa.data(), fa.size(), gtf.data(), gtf.size(), -1, out);
  TG_GUARD_END
}

tg_status tg_index_host_create_from_files_gpu(const char* fasta_path, const char* gtf_path, int device,
                                              tg_index_host** out) {
  if (!fasta_path || !gtf_path || !out || device < 0) return tg_fail(TG_ERR_INVALID, "null argument or negative device");
  TG_GUARD_BEGIN
  std::string fa, gtf;
  if (!slurp(fasta_path, fa)) return tg_fail(TG_ERR_IO, std::string("cannot read ") + fasta_path);
  if (!slurp(gtf_path, gtf)) return tg_fail(TG_ERR_IO, std::string("cannot read ") + gtf_path);
  return build_index(fa.data(), fa.size(), gtf.data(), gtf.size(), device, out);
  TG_GUARD_END
}

tg_status tg_index_host_blob(const tg_index_host* ix, const void** data, size_t* nbytes) {
  if (!ix || !data || !nbytes) return tg_fail(TG_ERR_INVALID, "null argument");
  *data = ix->blob.data();
  *nbytes = ix->blob.size();
  return TG_OK;
}

tg_status tg_index_host_from_blob(const void* data, size_t nbytes, tg_index_host** out) {
  if (!data || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  TG_GUARD_BEGIN
  auto* ix = new tg_index_host();
  ix->blob.assign((const uint8_t*)data, (const uint8_t*)data + nbytes);
  tg_status st = finish_host_index(ix);
  if (st != TG_OK) { delete ix; return st; }
  *out = ix;
  return TG_OK;
  TG_GUARD_END
}

tg_status tg_index_host_save(const tg_index_host* ix, const char* path) {
  if (!ix || !path) return tg_fail(TG_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return tg_fail(TG_ERR_IO, std::string("cannot write ") + path);
  size_t n = fwrite(ix->blob.data(), 1, ix->blob.size(), f);
  fclose(f);
  return n == ix->blob.size() ? TG_OK : tg_fail(TG_ERR_IO, "short write");
}

tg_status tg_index_host_load(const char* path, tg_index_host** out) {
  if (!path || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  TG_GUARD_BEGIN
  std::string s;
  if (!slurp(path, s)) return tg_fail(TG_ERR_IO, std::string("cannot read ") + path);
  return tg_index_host_from_blob(s.data(), s.size(), out);
  TG_GUARD_END
}

void tg_index_host_destroy(tg_index_host* ix) { delete ix; }

uint64_t tg_index_host_text_len(const tg_index_host* ix) { return ix->hdr()->text_len; }
uint32_t tg_index_host_n_refs(const tg_index_host* ix) { return (uint32_t)ix->hdr()->n_refs; }
uint32_t tg_index_host_n_txs(const tg_index_host* ix) { return (uint32_t)ix->hdr()->n_txs; }
uint32_t tg_index_host_n_genes(const tg_index_host* ix) { return (uint32_t)ix->hdr()->n_genes; }
const char* tg_index_host_ref(const tg_index_host* ix, uint32_t i, uint64_t* out4) {
  const TgRef& r = ((const TgRef*)(ix->blob.data() + ix->hdr()->off_refs))[i];
  out4[0] = r.start_idx; out4[1] = r.end_idx; out4[2] = r.len; out4[3] = r.strand_rank & 1;
  return ix->ref_names[i].c_str();
}
const char* tg_index_host_tx(const tg_index_host* ix, uint32_t i, uint64_t* out4) {
  const TgBlobHeader* h = ix->hdr();
  const uint64_t* so = (const uint64_t*)(ix->blob.data() + h->off_tx_seq_off);
  const uint32_t* eo = (const uint32_t*)(ix->blob.data() + h->off_tx_exon_off);
  out4[0] = ix->tx_gene[i]; out4[1] = ix->tx_strand[i]; out4[2] = eo[i + 1] - eo[i]; out4[3] = so[i + 1] - so[i];
  return ix->tx_ids[i].c_str();
}
const char* tg_index_host_gene_id(const tg_index_host* ix, uint32_t i) { return ix->gene_ids[i].c_str(); }
const char* tg_index_host_gene_name(const tg_index_host* ix, uint32_t i) { return ix->gene_names[i].c_str(); }
const uint32_t* tg_index_host_sa(const tg_index_host* ix) { return (const uint32_t*)(ix->blob.data() + ix->hdr()->off_sa); }
const uint64_t* tg_index_host_text4(const tg_index_host* ix) { return (const uint64_t*)(ix->blob.data() + ix->hdr()->off_text4); }

// ---- compact records -> wide records (include/thermite_gpu.h: tg_aln_c) ------------------------------------------------
tg_status tg_aln_expand(const tg_index_host* ix, const tg_aln_c* c, uint32_t read_len, tg_aln* out) {
  if (!ix || !c || !out) return tg_fail(TG_ERR_INVALID, "null argument");
  const TgBlobHeader* h = ix->hdr();
  if (!tg_expand_one((const TgRef*)(ix->blob.data() + h->off_refs), (uint32_t)h->n_refs,
                  (const uint64_t*)(ix->blob.data() + h->off_tx_seq_off), (uint32_t)h->n_txs, *c, read_len, *out))
    return tg_fail(TG_ERR_INVALID, "record does not belong to this index");
  return TG_OK;
}

tg_status tg_result_expand(const tg_index_host* ix, const tg_result_c* res, const uint64_t* offs, tg_aln* alns_out,
                           uint64_t* first_out, int n_threads) {
  if (!ix || !res || !offs || (!alns_out && res->alns_extent)) return tg_fail(TG_ERR_INVALID, "null argument");
  TG_GUARD_BEGIN
  const TgBlobHeader* h = ix->hdr();
  const TgRef* refs = (const TgRef*)(ix->blob.data() + h->off_refs);
  const uint64_t* tso = (const uint64_t*)(ix->blob.data() + h->off_tx_seq_off);
  const uint32_t n_refs = (uint32_t)h->n_refs, n_txs = (uint32_t)h->n_txs, n = res->n_reads;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if ((uint64_t)n < 65536) n_threads = 1;
  if (res->n_alns != res->alns_extent) memset(alns_out, 0, (size_t)res->alns_extent * sizeof(tg_aln));  // gaps between segments
  std::vector<int> bad(n_threads, 0);
  auto work = [&](int t) {
    const uint32_t r0 = (uint32_t)((uint64_t)n * t / n_threads), r1 = (uint32_t)((uint64_t)n * (t + 1) / n_threads);
    for (uint32_t r = r0; r < r1; r++) {
      const uint32_t f = res->read_aln_first[r], k = res->read_aln_count[r], L = (uint32_t)(offs[r + 1] - offs[r]);
      if (first_out) first_out[r] = f;
      if ((uint64_t)f + k > res->alns_extent) { if (k) bad[t] = 1; continue; }
      for (uint32_t i = 0; i < k; i++)
        if (!tg_expand_one(refs, n_refs, tso, n_txs, res->alns[f + i], L, alns_out[f + i])) bad[t] = 1;
    }
  };
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  for (int b : bad)
    if (b) return tg_fail(TG_ERR_INVALID, "result does not belong to this index");
  return TG_OK;
  TG_GUARD_END
}

}  // extern "C"
