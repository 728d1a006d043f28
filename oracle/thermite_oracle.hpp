// thermite_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
//
// A scalar, single-threaded C++17 restatement of the read-alignment hot path of
// 10XGenomics/thermite (reference files cited per function as `src/<file>.rs:<lines>`).
// Nothing in the shipped product (`thermite_b200/`, `include/`) may include, link or
// call this code: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
// `--impl reference` legs of `bench.py` use it, and only as the checker / CPU baseline.
//
// PARITY STATUS
//   pinned   : SwgExtend::extend (src/swg.rs:249-317 KATs), extend_left_right
//              (src/aligner.rs:603-639), filter_overlapping (src/aligner.rs:472-601),
//              lift_mem_to_tx / lift_tx_to_gx (src/txome.rs:168-340).
//   UNPINNED : everything that lives in un-vendored crates -- rust-bio 0.37.1 FMD-index SMEM
//              search + sampled-SA locate + AVL IntervalTree iteration order, libdivsufsort
//              suffix order, cellranger `transcriptome` GTF order, needletail parsing, noodles
//              SAM text.  Those are restated from their published algorithms (SURVEY.md
//              Appendix A) and anchored on the reference's call sites; the reference has no
//              test or golden vector for them and cannot be compiled here (no cargo/rustc).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// bio::alignment::pairwise::MIN_SCORE (rust-bio 0.37.1) -- used by src/swg.rs:1,67,76-90
static const int32_t MIN_SCORE = -858993459;

// bio::alignment::AlignmentOperation
enum OpKind : uint8_t { Match = 0, Subst = 1, Del = 2, Ins = 3, Xclip = 4, Yclip = 5 };
struct Op {
  uint8_t kind;
  uint64_t n;  // 1 for Match/Subst/Del/Ins; the clip length for Xclip/Yclip
  bool operator==(const Op& o) const { return kind == o.kind && n == o.n; }
};

// bio::alignment::Alignment (mode is always Custom on this path)
struct Alignment {
  int32_t score = 0;
  size_t ystart = 0, xstart = 0, yend = 0, xend = 0, ylen = 0, xlen = 0;
  std::vector<Op> operations;
};

struct ReferencePanic : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// src/swg.rs:6-241
class SwgExtend {
 public:
  SwgExtend(size_t max_band_width, int gap_open, int gap_extend, int match, int mismatch);
  Alignment extend(const uint8_t* x, size_t xlen, const uint8_t* y, size_t ylen, size_t band_width,
                   int32_t x_drop);
  uint64_t cells = 0;  // DP cells visited by the reference loops (src/swg.rs:80,119); GCUPS unit
 private:
  std::vector<Op> trace_path(size_t i, size_t j, size_t len, size_t band_width) const;
  void set_trace(size_t j, size_t i, uint8_t op);
  uint8_t get_trace(size_t j, size_t i) const;
  std::vector<int32_t> D, C, R;
  std::vector<uint8_t> trace;
  int go, ge, ma, mi;
  size_t max_band_width;
};

// src/index.rs:383-399
struct Mem {
  size_t ref_idx, query_idx, len;
  bool operator==(const Mem& o) const {
    return ref_idx == o.ref_idx && query_idx == o.query_idx && len == o.len;
  }
};
struct Ref {
  std::string name;
  bool has_seq;  // forward strand keeps its sequence (src/index.rs:80,95)
  std::vector<uint8_t> seq;
  bool strand;
  size_t len, start_idx, end_idx;
};

// src/txome.rs:10-48
struct Exon {
  size_t start, end, tx_idx;
  size_t len() const { return end - start; }
};
struct Tx {
  std::string id, chrom;
  bool strand;
  std::vector<Exon> exons;
  std::vector<uint8_t> seq;
  size_t gene_idx;
};
struct Gene {
  std::string id, name;
};

// bio::data_structures::interval_tree::IntervalTree<usize, usize> (AVL)
class IntervalTree {
 public:
  void insert(size_t start, size_t end, size_t data);
  // values of all intervals intersecting [start,end) in the iterator's order (node, right, left)
  std::vector<size_t> find(size_t start, size_t end) const;
  // every entry with its rank in the un-pruned node->right->left traversal
  struct Flat { size_t start, end, data, rank; };
  std::vector<Flat> flatten() const;
  struct Node {
    size_t start, end, data, max;
    int height;
    std::unique_ptr<Node> left, right;
  };
 private:
  std::unique_ptr<Node> root;
};

struct Txome {
  std::vector<Gene> genes;
  std::vector<Tx> txs;
  IntervalTree exon_to_tx, gene_intervals;
};

// src/txome.rs:55-69
enum AlnTypeKind : uint8_t { Exonic = 0, Intronic = 1, Intergenic = 2 };
struct GenomeAlignment {
  Alignment gx_aln;
  AlnTypeKind aln_type = Intergenic;
  Alignment tx_aln;    // Exonic only
  size_t tx_idx = 0;   // Exonic only
  size_t gene_idx = 0; // Intronic only
  std::string ref_name;
  size_t ref_id = 0;   // index into Index::refs of the hit's Ref (not in the reference struct;
                       // (ref_name, strand) identify it uniquely)
  bool strand = true;
  bool primary = false;
};

// src/aligner.rs:452-464 ; defaults src/main.rs:115-132
struct AlignOpts {
  size_t min_seed_len = 20;
  float min_aln_score_percent = 0.66f;
  int32_t min_aln_score = 30;
  size_t multimap_score_range = 1;
  bool intron_mode = false;
};

struct FastaRecord {
  std::string id;  // whole header line without '>'
  std::vector<uint8_t> seq;
};
struct FastqRecord {
  std::string id;
  std::vector<uint8_t> seq, qual;
};
std::vector<FastaRecord> parse_fasta(const std::string& text);
std::vector<FastqRecord> parse_fastq(const std::string& text);
std::string read_file(const std::string& path);

// Per-call work counters (for roofline units; not part of the reference)
struct Counters {
  uint64_t swg_cells = 0, swg_calls = 0, occ_lookups = 0, fmd_ext = 0, sa_locates = 0, hits = 0;
};

// src/index.rs:40-44
class Index {
 public:
  // src/index.rs:52-223
  static Index create(const std::vector<FastaRecord>& fasta, const std::string& gtf_text,
                      size_t sa_sampling_rate = 32, size_t occ_sampling_rate = 128);
  static Index create_from_files(const std::string& ref_path, const std::string& annot_path,
                                 size_t sa_sampling_rate = 32, size_t occ_sampling_rate = 128);
  std::vector<Mem> all_smems(const uint8_t* query, size_t qlen, size_t min_seed_len) const;
  // definition-based SMEMs straight from the text (cross-check for the FMD restatement)
  std::vector<Mem> all_smems_brute(const uint8_t* query, size_t qlen, size_t min_seed_len) const;
  size_t idx_to_ref(size_t idx) const;                        // src/index.rs:287-290
  std::vector<uint8_t> seq_slice(size_t start, size_t end) const;  // src/index.rs:304-323

  std::vector<Ref> refs;
  Txome txome;
  // FM structures
  std::vector<uint8_t> bwt;
  std::vector<size_t> less;            // 256+1 entries
  std::vector<uint32_t> occ_samples;   // [checkpoint][6] for $ A C G N T
  size_t occ_rate = 128, sa_rate = 32;
  std::vector<uint32_t> sa_samples;    // SA[r] for r % sa_rate == 0
  std::map<size_t, size_t> sa_extra;   // rows whose BWT byte is the sentinel
  std::vector<uint8_t> text;           // kept ONLY for all_smems_brute / tests (reference drops it)
  std::vector<uint32_t> full_sa;       // kept ONLY for tests (suffix-order checks)
  mutable Counters counters;       // totals (single-thread runs update them directly)
  static Counters& tl();           // per-thread counters used while aligning; folded into `counters` by fold_tl()
  void fold_tl() const;

  size_t occ(size_t r, uint8_t a) const;
  size_t sa_get(size_t r) const;
};

// suffix array of `text` in plain byte-lexicographic order (what divsufsort64 returns,
// src/index.rs:103-105); prefix doubling, independent of the product's SA-IS.
std::vector<uint32_t> suffix_array(const std::vector<uint8_t>& text);

// src/txome.rs:77-160
bool intersect(size_t a0, size_t a1, size_t b0, size_t b1);
Mem lift_mem_to_tx(const Mem& mem, const Tx& tx);
Alignment lift_tx_to_gx(const Alignment& tx_aln, const Tx& tx);

// src/aligner.rs:123-449
std::vector<GenomeAlignment> align_read(const Index& index, const uint8_t* read, size_t len,
                                        const AlignOpts& opts);
GenomeAlignment align_seed_hit(const Index& index, const std::vector<uint8_t>& read, const Mem& hit,
                               SwgExtend& swg, size_t band_width, int32_t x_drop);
std::vector<GenomeAlignment> filter_overlapping(std::vector<GenomeAlignment> alns);
Alignment extend_left_right(const uint8_t* ref_seq, size_t ref_len, const Mem& hit,
                            const uint8_t* read, size_t read_len, SwgExtend& swg, size_t band_width,
                            int32_t x_drop);
void extend_seed_match(const uint8_t* ref_seq, size_t ref_len, Mem& hit, const uint8_t* read,
                       size_t read_len);

// src/aln_writer.rs
uint8_t multimapq(size_t n);                                        // :332-340
std::string cigar_string(const std::vector<Op>& ops);               // :279-323
std::string paf_line(const std::string& query_name, size_t query_len, const GenomeAlignment& aln,
                     size_t multimap);                              // :47-116
std::string sam_header(const Index& index);                         // :256-276 (noodles text: unpinned)
std::string sam_line(const Index& index, const FastqRecord& rec, const GenomeAlignment& aln,
                     size_t multimap, size_t hit_index);            // :118-238
std::string sam_unmapped_line(const FastqRecord& rec);              // :241-253
// src/aligner.rs:22-120 (PAF / SAM text of a whole FASTQ)
std::string align_fastq(const Index& index, const std::vector<FastqRecord>& reads,
                        const AlignOpts& opts, bool sam);

std::vector<uint8_t> revcomp(const uint8_t* s, size_t n);  // bio::alphabets::dna::revcomp

}  // namespace orc
