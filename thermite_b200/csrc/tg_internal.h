// tg_internal.h -- layouts shared by the host index builder, the C ABI glue and the CUDA kernels.
// Product code: must never include anything from oracle/.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/thermite_gpu.h"

// ------------------------------------------------------------------------------------------------
// Symbol codes.  Text and transcripts use 0..5 in the byte order of the reference's FM alphabet
// ("$ACGNT", src/index.rs:108 + the '$' separators of :76,91) so that packed words compare
// lexicographically.  Reads additionally use 7 for any byte outside ACGNT (matches nothing; the
// reference's FM index has no such symbol) and 15 as end-of-read padding.
// ------------------------------------------------------------------------------------------------
enum : uint8_t { TG_C_SENT = 0, TG_C_A = 1, TG_C_C = 2, TG_C_G = 3, TG_C_N = 4, TG_C_T = 5, TG_C_OTHER = 7, TG_C_PAD = 15 };

// One interval-tree node of the flattened AVL tree (rust-bio IntervalTree shape; src/index.rs:135,182,208).
// Arrays of these reproduce `find()` by running the same stack DFS on the device.
struct alignas(16) TgTreeNode {  // 32 bytes: two 128-bit loads on the device
  uint32_t start, end, max, data;
  int32_t left, right;
  uint32_t pad0, pad1;
};

// One interval of a stab list: sorted by start; `rank` is the interval's position in the un-pruned traversal of the
// rust-bio AVL tree (node, right subtree, left subtree), so results ordered by rank reproduce find()'s order.
struct alignas(16) TgStab {
  uint32_t start, end, data, rank;
};

struct TgRef {
  uint32_t start_idx, end_idx, len;
  uint32_t strand_rank;  // bit0 = strand (1 forward), bits 1.. = rank of the name in byte order (filter_overlapping key)
};

#define TG_BLOB_MAGIC 0x3330424947544854ull /* "THTGIB03" */

// Position-independent header at offset 0 of the index blob.  All off_* are byte offsets from the blob
// start, 256-byte aligned.
struct TgBlobHeader {
  uint64_t magic, nbytes;
  uint64_t text_len;       // T = 2 * (sum chrom len + n_chrom)
  uint64_t n_refs, n_txs, n_genes, n_exon_nodes, n_gene_nodes, n_tx_exons, txseq_len;
  int64_t exon_root, gene_root;
  uint64_t device_bytes;   // prefix of the blob the GPU needs (everything before the host-only metadata)
  uint64_t off_text4, off_sa, off_refs, off_exon_nodes, off_gene_nodes, off_tx_seq_off, off_tx_exon_off,
      off_te_start, off_te_end, off_txseq4;
  // host-only metadata (string tables: u64 offsets[n+1] followed by bytes)
  uint64_t off_ref_names, off_tx_ids, off_gene_ids, off_gene_names, off_tx_gene, off_tx_strand;
  // interval lists sorted by start for warp-cooperative stabbing (rank = position in the tree's find() order)
  uint64_t off_exon_stab, off_gene_stab, exon_maxlen, gene_maxlen;
  uint64_t format_version;  // TG_BLOB_VERSION: bumped whenever a section layout or a device struct changes
  uint64_t reserved[3];
};
#define TG_BLOB_VERSION 4ull

// Checks a header against the number of bytes that are really there (`avail`; the whole blob on the host, the device
// prefix on the GPU side): magic, version, every section inside the blob, counts consistent with each other.  A
// truncated, stale or foreign file must fail here and never reach decode_strings or a kernel.  Returns an error text or
// nullptr.  (Section contents -- string offsets, exon tables -- are checked by tg_blob_validate_host.)
const char* tg_blob_check_header(const TgBlobHeader& h, uint64_t avail, bool device_prefix_only);
// The host-side walk over the section contents (monotone string tables, exon / transcript offset tables, refs).
const char* tg_blob_validate_host(const uint8_t* blob, uint64_t nbytes);

// Device-side view (plain pointers into the device copy of the blob).
struct TgIndexDev {
  const uint64_t* text4;
  const uint32_t* sa;
  const TgRef* refs;
  const TgTreeNode* exon_nodes;
  const TgTreeNode* gene_nodes;
  const TgStab* exon_stab;
  const TgStab* gene_stab;
  uint32_t n_exon_stab, n_gene_stab, exon_maxlen, gene_maxlen;
  const uint64_t* tx_seq_off;
  const uint32_t* tx_exon_off;
  const uint32_t* te_start;
  const uint32_t* te_end;
  const uint64_t* txseq4;
  uint64_t text_len;
  uint32_t n_refs, n_txs;
  int32_t exon_root, gene_root;
};

// k-mer table slot (16 B, one 128-bit load).  tag == 0: empty.
struct TgSlot {
  uint32_t tag;
  uint32_t lo;    // first suffix-array row of the k-mer, or the text position itself when count == 1
  uint32_t count;
  uint32_t pad;
};

struct tg_index_host {
  std::vector<uint8_t> blob;
  const TgBlobHeader* hdr() const { return (const TgBlobHeader*)blob.data(); }
  // decoded metadata
  std::vector<std::string> ref_names, tx_ids, gene_ids, gene_names;
  std::vector<uint32_t> tx_gene, tx_strand;
};

// compact record -> wide record (include/thermite_gpu.h: tg_aln_c); false when an index is out of range
inline bool tg_expand_one(const TgRef* refs, uint32_t n_refs, const uint64_t* tx_seq_off, uint32_t n_txs, const tg_aln_c& c,
                              uint32_t read_len, tg_aln& a) {
  if (c.ref_id >= n_refs) return false;
  const TgRef& r = refs[c.ref_id];
  a.ystart = c.ystart; a.yend = c.yend; a.ylen = r.len;
  a.score = c.score; a.ref_id = c.ref_id;
  a.xstart = c.xstart; a.xend = c.xend; a.xlen = read_len;
  a.tx_or_gene_idx = c.tx_or_gene_idx;
  a.ops_off = c.ops_off; a.ops_len = c.ops_len;
  a.aln_type = c.aln_type; a.primary = c.primary; a.strand = (uint8_t)(r.strand_rank & 1u); a.pad = 0;
  if (c.aln_type == TG_ALN_EXONIC) {
    if (c.tx_or_gene_idx >= n_txs) return false;
    a.tx_ystart = c.tx_ystart; a.tx_yend = c.tx_yend;
    a.tx_ylen = tx_seq_off[c.tx_or_gene_idx + 1] - tx_seq_off[c.tx_or_gene_idx];
    a.tx_score = c.score; a.tx_xstart = c.xstart; a.tx_xend = c.xend;
    a.tx_ops_off = c.ops_off + c.ops_len; a.tx_ops_len = c.tx_ops_len;
  } else {
    a.tx_ystart = 0; a.tx_yend = 0; a.tx_ylen = 0; a.tx_score = 0; a.tx_xstart = 0; a.tx_xend = 0;
    a.tx_ops_off = 0; a.tx_ops_len = 0;
  }
  return true;
}

void tg_set_error(const std::string& msg);
tg_status tg_fail(tg_status code, const std::string& msg);

// Nothing throws across the C ABI: every extern "C" entry point that can allocate or start threads runs inside these.
#define TG_GUARD_BEGIN try {
#define TG_GUARD_END                                                             \
  }                                                                              \
  catch (const std::exception& e) { return tg_fail(TG_ERR_INTERNAL, e.what()); } \
  catch (...) { return tg_fail(TG_ERR_INTERNAL, "unknown exception"); }

// thermite_gpu.cu, used by tg_multi.cpp: one shard of a multi-GPU batch on one context.  The records (compact) go straight
// into the caller's pinned result segment; first indices / operation offsets are rebased on the device.
struct TgHostSegment {
  uint32_t* first;        // [n] of the shard
  uint32_t* count;        // [n]
  tg_aln_c* alns;         // segment start
  uint32_t* ops;          // segment start
  size_t alns_cap, ops_cap;                    // segment capacity in elements
  unsigned long long first_base, ops_base;     // index of the segment start inside the whole result's pools
};
struct TgShardStat {
  bool overflow;                     // the segment was too small: need_* hold what the shard needs, nothing else is valid
  uint64_t n_alns, n_ops, need_alns, need_ops;
  uint64_t swg_cells, swg_extensions, seed_hits, n_smems;
  double wall_ms;
  float seed_ms, extend_ms, dp_ms;
};
tg_status tg_ctx_align_segment(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads,
                               const TgHostSegment& seg, TgShardStat* stat);
int tg_ctx_device(const tg_ctx* ctx);

// host_batcher.cpp: the micro-batcher over any batch aligner (tg_align_batch in the product; a stand-in in the host test)
typedef tg_status (*tg_batch_backend_fn)(void* user, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads,
                                         tg_result* out);
tg_status tg_batcher_create_backend(tg_batch_backend_fn fn, void* user, uint32_t max_batch_reads, uint32_t max_wait_us,
                                    tg_batcher** out);

// host_index.cpp
// GPU suffix-array builder (csrc/tg_sa.cu), registered when that file is linked in; nullptr in the host test library.
typedef tg_status (*tg_sa_device_fn)(const uint64_t* text4, uint64_t text_len, int device, uint32_t* sa_out, float* ms,
                                     uint32_t* steps);
extern tg_sa_device_fn g_tg_sa_device;
void tg_sais(const uint8_t* text, size_t n, int32_t* sa);  // plain byte-lexicographic suffix order
