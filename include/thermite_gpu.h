/* thermite_gpu.h -- C ABI of libthermite_gpu.so: the B200 (sm_100a) replacement for the read-alignment hot
 * path of 10XGenomics/thermite.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * What each entry point replaces in the reference (paths under /root/reference):
 *   tg_index_host_create_*   Index::create_from_files              src/index.rs:52-223
 *   tg_index_host_create_*_gpu, tg_suffix_array_gpu   the same with divsufsort64 (src/index.rs:103-105) on the GPU
 *   tg_index_host_save/load  bincode .tai (de)serialisation        src/main.rs:37-43, 63-67; src/wrapper.rs:31-37
 *   tg_index_host_* getters  Index::refs() / Index::txome()        src/index.rs:293-300 (used by src/aln_writer.rs:179-213,257)
 *   tg_index_create*         (new) upload / adopt the flat index in HBM; one replica per GPU
 *   tg_ctx_create            AlignOpts + per-read SwgExtend scratch  src/aligner.rs:452-464, :140-141
 *   tg_align_batch           align_read, batched                   src/aligner.rs:123-190 (+ :198-449, src/swg.rs, src/txome.rs:82-160,
 *                                                                   Index::all_smems src/index.rs:228-255)
 *   tg_format_bam_header / tg_format_batch_bam   bam::Writer path   src/aligner.rs:41-47, 69-72, 98-101
 *   tg_batcher_*             ThermiteAligner::align_read from many threads  src/wrapper.rs:20-27, :72
 *   tg_seed_batch            Index::all_smems, batched             src/index.rs:228-255
 *   tg_swg_extend_batch      SwgExtend::extend, batched            src/swg.rs:31-207
 *   tg_format_paf / _sam     PafEntry / aln_to_sam_record          src/aln_writer.rs:47-116, 118-253 ; src/aligner.rs:54-115
 *
 * Every function returns tg_status (0 = ok, < 0 = error); tg_last_error() gives the thread-local message.
 * Nothing throws or aborts across the ABI.  There is no CPU fallback: without a CUDA device every device
 * entry point fails with TG_ERR_CUDA.
 *
 * Threading: a tg_index is immutable and may be shared by any number of contexts/threads.  A tg_ctx owns a
 * CUDA stream plus scratch and is NOT thread-safe: use one per host thread / per GPU (the reference's
 * ThermiteAligner is Clone + Send with an Arc<Index>, src/wrapper.rs:20-27).
 *
 * Scale limits of this build (the reference uses usize throughout, src/index.rs:383-388); every one is reported as
 * TG_ERR_CAPACITY / TG_ERR_INVALID, never as a wrong result:
 *   - both-strand concatenated text < 2^31 symbols (32-bit suffix array and text positions): about 1.07 Gbp of reference,
 *     i.e. any single human chromosome or a mouse-sized transcriptome, not a whole human genome in one index;
 *   - reads <= TG_MAX_READ_LEN (512) symbols; min_seed_len <= 32;
 *   - one batch: < 2^32 reads, seed hits, records and operation words (split larger batches);
 *   - the k-mer table takes 16 B x the next power of two above twice the number of distinct k-mers (4.3 GB for chr21,
 *     shared by all contexts of an index that use the same min_seed_len).
 */
#ifndef THERMITE_GPU_H
#define THERMITE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t tg_status;
enum {
  TG_OK = 0,
  TG_ERR_INVALID = -1,   /* bad argument (null pointer, read too long, x_drop < band width, ...) */
  TG_ERR_IO = -2,        /* file could not be read / parsed */
  TG_ERR_CUDA = -3,      /* CUDA runtime error or no device */
  TG_ERR_CAPACITY = -4,  /* a documented fixed limit was exceeded (see TG_MAX_*) */
  TG_ERR_INTERNAL = -5
};

#define TG_MAX_READ_LEN 512u        /* longest read accepted by tg_align_batch / tg_seed_batch */
#define TG_MAX_SEED_LEN 32u         /* largest min_seed_len (k-mer table key is k 4-bit symbols) */
#define TG_MAX_ALNS_PER_READ 4096u  /* near-best alignments (score >= running max - range) kept per read before the final filters */

const char* tg_last_error(void);
/* Library / build identification: "thermite_gpu <ver> sm_100a". */
const char* tg_version(void);

/* ---------------------------------------------------------------------------------------------------
 * Alignment operations: one u32 word = kind | (run << 3).
 *   kind 0 Match, 1 Subst, 2 Del, 3 Ins : run = number of consecutive unit operations of that kind
 *   kind 4 Xclip (soft clip), 5 Yclip (intron): run = clip length; every clip is its own word
 * Expanding the words in order gives bio::alignment::Alignment::operations one for one.
 * ------------------------------------------------------------------------------------------------- */
enum { TG_OP_MATCH = 0, TG_OP_SUBST = 1, TG_OP_DEL = 2, TG_OP_INS = 3, TG_OP_XCLIP = 4, TG_OP_YCLIP = 5 };
enum { TG_ALN_EXONIC = 0, TG_ALN_INTRONIC = 1, TG_ALN_INTERGENIC = 2 };

/* GenomeAlignment (src/txome.rs:55-69) as a flat record. */
typedef struct tg_aln {
  uint64_t ystart, yend, ylen;            /* gx_aln: chromosome coordinates, forward-strand orientation; ylen = chromosome length */
  uint64_t tx_ystart, tx_yend, tx_ylen;   /* Exonic only: tx_aln in transcript coordinates; tx_ylen = transcript length */
  int32_t score;                          /* gx_aln.score */
  uint32_t ref_id;                        /* index into refs() of the hit's Ref => ref_name and strand */
  uint32_t xstart, xend, xlen;            /* read coordinates; xlen = read length */
  uint32_t tx_or_gene_idx;                /* Exonic: tx_idx; Intronic: gene_idx; Intergenic: 0xFFFFFFFF */
  int32_t tx_score;                       /* Exonic only */
  uint32_t tx_xstart, tx_xend;            /* Exonic only */
  uint32_t ops_off, ops_len;              /* gx_aln.operations: words [ops_off, ops_off+ops_len) of tg_result.ops */
  uint32_t tx_ops_off, tx_ops_len;        /* Exonic only: tx_aln.operations */
  uint8_t aln_type;                       /* TG_ALN_* */
  uint8_t primary;
  uint8_t strand;                         /* 1 = forward */
  uint8_t pad;
} tg_aln;

/* AlignOpts (src/aligner.rs:452-464); tg_opts_default() gives src/main.rs:115-132 / src/wrapper.rs:40-46. */
typedef struct tg_opts {
  uint32_t min_seed_len;          /* -k, default 20 */
  float min_aln_score_percent;    /* -s, default 0.66 */
  int32_t min_aln_score;          /* --min-aln-score, default 30 */
  uint32_t multimap_score_range;  /* --multimap-score-range, default 1 */
  uint32_t intron_mode;           /* --intron-mode, default 0 */
} tg_opts;
void tg_opts_default(tg_opts* out);

/* ---------------------------------------------------------------------------------------------------
 * Host-side flat index
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_index_host tg_index_host;

/* FASTA (+ GTF) -> concatenated both-strand text, suffix array, flattened transcriptome.  The .fai the
 * reference needs next to the FASTA is not required. */
tg_status tg_index_host_create_from_files(const char* fasta_path, const char* gtf_path, tg_index_host** out);
tg_status tg_index_host_create_from_memory(const char* fasta_text, size_t fasta_len, const char* gtf_text,
                                           size_t gtf_len, tg_index_host** out);
/* Same index, but the suffix array -- divsufsort64 in the reference (src/index.rs:103-105), the dominant cost of index
 * creation -- is built on GPU `device` by prefix doubling over the unresolved suffixes (csrc/tg_sa.cu).
 * The blob is byte-identical to the one the host-only functions produce.  Needs ~46 B of HBM per text symbol. */
tg_status tg_index_host_create_from_files_gpu(const char* fasta_path, const char* gtf_path, int device,
                                              tg_index_host** out);
tg_status tg_index_host_create_from_memory_gpu(const char* fasta_text, size_t fasta_len, const char* gtf_text,
                                               size_t gtf_len, int device, tg_index_host** out);
/* The suffix-array builder on its own.  text4 = HOST pointer to text_len symbols as 4-bit codes ($ACGNT = 0..5), 16 per
 * u64 with the first symbol in the most significant nibble, followed by at least one zero word; sa_out = HOST buffer of
 * text_len u32.  Order: plain lexicographic, a suffix that is a proper prefix of another one first (what divsufsort64
 * returns for the reference's `$`-joined text).  device_ms / n_steps (optional): device time without the copies and
 * the number of sort steps taken. */
tg_status tg_suffix_array_gpu(const uint64_t* text4, uint64_t text_len, int device, uint32_t* sa_out, float* device_ms,
                              uint32_t* n_steps);
/* One contiguous, position-independent blob (what is saved to disk, uploaded, or broadcast over NCCL). */
tg_status tg_index_host_blob(const tg_index_host* ix, const void** data, size_t* nbytes);
tg_status tg_index_host_from_blob(const void* data, size_t nbytes, tg_index_host** out);
tg_status tg_index_host_save(const tg_index_host* ix, const char* path);
tg_status tg_index_host_load(const char* path, tg_index_host** out);
void tg_index_host_destroy(tg_index_host* ix);

/* refs() / txome() accessors for the writers.  Returned strings live as long as the index. */
uint64_t tg_index_host_text_len(const tg_index_host* ix);
uint32_t tg_index_host_n_refs(const tg_index_host* ix);
uint32_t tg_index_host_n_txs(const tg_index_host* ix);
uint32_t tg_index_host_n_genes(const tg_index_host* ix);
/* out4 = { start_idx, end_idx, len, strand } */
const char* tg_index_host_ref(const tg_index_host* ix, uint32_t i, uint64_t* out4);
/* out4 = { gene_idx, strand, n_exons, seq_len } ; returns transcript id */
const char* tg_index_host_tx(const tg_index_host* ix, uint32_t i, uint64_t* out4);
const char* tg_index_host_gene_id(const tg_index_host* ix, uint32_t i);
const char* tg_index_host_gene_name(const tg_index_host* ix, uint32_t i);
/* Suffix array (text_len u32 entries) -- exposed so callers can expand tg_seed records into Mems. */
const uint32_t* tg_index_host_sa(const tg_index_host* ix);
/* Packed both-strand text (text_len 4-bit codes, 16 per u64, text_len / 16 + 4 words): the input of tg_suffix_array_gpu. */
const uint64_t* tg_index_host_text4(const tg_index_host* ix);

/* ---------------------------------------------------------------------------------------------------
 * Device index (HBM-resident replica)
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_index tg_index;
tg_status tg_index_create(const tg_index_host* ix, int device, tg_index** out);
/* Adopt a blob that already sits in device memory (e.g. after an NCCL broadcast).  The memory stays owned
 * by the caller and must outlive the index. */
tg_status tg_index_create_from_device_blob(const void* device_blob, size_t nbytes, int device, tg_index** out);
void tg_index_destroy(tg_index* ix);

/* ---------------------------------------------------------------------------------------------------
 * Alignment context: options, stream, k-mer table for opts.min_seed_len, scratch.
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_ctx tg_ctx;
tg_status tg_ctx_create(const tg_index* ix, const tg_opts* opts, tg_ctx** out);
void tg_ctx_destroy(tg_ctx* ctx);
/* tg_align_batch copies the bases of a large batch in chunks of this many reads (0 = default: an eighth of the batch, at
 * least 262,144 reads) and seeds every chunk
 * as soon as it has landed; at the other end the records of the reads that are finished after the second round travel to
 * the host while the late rounds run.  Results do not depend on the chunk size. */
void tg_ctx_set_chunk_reads(tg_ctx* ctx, uint32_t reads);
/* Host result buffers of tg_align_batch / tg_align_batch_compact.  n = 1 (default): a result is valid until the next call
 * on the context.  n = 2: two buffer sets used alternately, so a result stays valid until the SECOND next call -- a
 * consumer (writer thread) may still read batch k while batch k + 1 is being aligned. */
tg_status tg_ctx_set_result_buffers(tg_ctx* ctx, int n);
/* The CUDA stream (cudaStream_t) all of the context's work is enqueued on. */
void* tg_ctx_stream(tg_ctx* ctx);
/* Device time (CUDA events on the context's stream) of the seeding and extension kernels of the last
 * tg_align_batch* / tg_seed_batch call, in milliseconds. */
void tg_ctx_last_kernel_ms(const tg_ctx* ctx, float* seed_ms, float* extend_ms);
/* Cell accounting.  By default an extension stops as soon as no later DP cell can STRICTLY exceed the running
 * maximum (bound: best cell of the column + symbols of x still unread).  Score, end cell, traceback and therefore
 * every output record are unchanged, but fewer cells are visited than the reference's loops visit.  With `on` = 1
 * every column the reference runs is run, so tg_result.swg_cells / *cells equal the reference's cell count. */
void tg_ctx_set_exact_cell_count(tg_ctx* ctx, int on);
/* Execution strategy of the hit loop.  on = 1 (default): speculative round pipeline -- every round evaluates a batch of
 * consecutive hits of each unfinished read under the read's current (band_width, x_drop) with thread-per-hit control
 * kernels and a warp-per-task extension kernel, then replays the reference's serial accept / narrow logic over the
 * batch; hits evaluated under a state that an accepted hit changed are re-submitted under predicted states
 * (csrc/tg_rounds.h).  Reads it
 * cannot hold (more than 12 transcripts on one seed, ~100k hits) run on the single-warp kernel.  on = 0: every read on
 * the single-warp kernel.  Both produce identical records. */
void tg_ctx_set_round_pipeline(tg_ctx* ctx, int on);
/* Device time of the banded-SWG kernels alone (the DP sections of all rounds, CUDA events) in the last tg_align_batch*
 * call, in milliseconds.  Part of extend_ms above. */
float tg_ctx_last_dp_ms(const tg_ctx* ctx);
/* Number of kernels the last tg_align_batch* / tg_seed_batch call launched on the context's stream. */
uint64_t tg_ctx_last_kernel_launches(const tg_ctx* ctx);
/* Measurement aid for the seeding roofline: rate of independent random 16-B loads (one 32-B sector each) over the
 * context's own k-mer table, in GB/s of sectors, best of `reps` launches (CUDA events). */
tg_status tg_bench_random_gather(tg_ctx* ctx, uint64_t n_loads, int reps, double* sector_gbs, float* best_ms);
/* Measurement aid for the SWG roofline: lane-operations per second of dependency-free streams of the two instructions
 * the DP inner loop consists of, VIADDMNMX (max(a + b, c)) and VIMNMX3 (max(a, b, c)), best of `reps` launches. */
tg_status tg_bench_int_peak(tg_ctx* ctx, int reps, double* viaddmnmx_lane_ops, double* vimnmx3_lane_ops);
/* Size of the context's k-mer table in bytes. */
uint64_t tg_ctx_kmer_table_bytes(const tg_ctx* ctx);

typedef struct tg_result {
  uint32_t n_reads;
  uint64_t n_alns, n_ops;
  const uint64_t* read_aln_first; /* [n_reads] index of the read's first record in alns */
  const uint32_t* read_aln_count; /* [n_reads] number of records (0 = unmapped); records are in output order */
  const tg_aln* alns;             /* [n_alns] */
  const uint32_t* ops;            /* [n_ops] */
  /* work counters of this batch */
  uint64_t swg_cells;             /* DP cells as the reference's loops visit them (src/swg.rs:80,119) */
  uint64_t swg_extensions;        /* non-trivial SwgExtend::extend calls */
  uint64_t seed_hits;             /* Mems handed to align_seed_hit */
  uint64_t n_smems;
} tg_result;

/* align_read for n_reads reads.  `bases` = concatenated ASCII reads (any case), read r = bases[offs[r], offs[r+1]).
 * HOST buffers in, HOST result out (owned by ctx, valid until the next call on ctx or tg_ctx_destroy).
 * The call does H2D, all kernels, D2H, and synchronises the context's stream. */
tg_status tg_align_batch(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result* out);
/* Same, but inputs are DEVICE pointers and the result pointers are DEVICE pointers (counters and n_* are
 * still filled on the host after a stream sync). */
tg_status tg_align_batch_device(tg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n_reads,
                                uint64_t total_bases, uint32_t max_read_len, tg_result* out);

/* ---------------------------------------------------------------------------------------------------
 * Compact results.  tg_aln spends 104 B on a record; 64 of them repeat what the caller already holds (ylen = length of
 * refs()[ref_id], tx_ylen = length of txome().txs[tx_idx], xlen = the read's length, strand = refs()[ref_id].strand) or
 * what the record says twice (tx_score / tx_xstart / tx_xend equal score / xstart / xend of an Exonic record, and the
 * transcript operations follow the genome operations directly).  tg_aln_c is the same GenomeAlignment
 * (src/txome.rs:55-69) in 40 B; tg_aln_expand gives back the wide form bit for bit.  The compact calls move 70 B per
 * read to the host instead of 141 B, which is what bounds the end-to-end rate when 8 GPUs share one host.
 * Limits: reads <= TG_MAX_READ_LEN (16-bit read coordinates and score), chromosomes < 2^32 bp.
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_aln_c {
  uint32_t ystart, yend;          /* gx_aln: chromosome coordinates, forward-strand orientation */
  uint32_t tx_ystart, tx_yend;    /* Exonic only: tx_aln in transcript coordinates */
  uint32_t ref_id;                /* index into refs() => ref_name, strand, ylen */
  uint32_t tx_or_gene_idx;        /* Exonic: tx_idx; Intronic: gene_idx; Intergenic: 0xFFFFFFFF */
  uint32_t ops_off;               /* gx_aln.operations = ops[ops_off, ops_off + ops_len); tx_aln.operations = the tx_ops_len words behind them */
  int16_t score;
  uint16_t xstart, xend;          /* read coordinates */
  uint16_t ops_len, tx_ops_len;
  uint8_t aln_type;               /* TG_ALN_* */
  uint8_t primary;
} tg_aln_c;

/* Like tg_result with compact records.  The pools are SEGMENTED: tg_align_batch_compact fills one segment,
 * tg_multi_align_batch one per GPU (every GPU's copies land at their final place, nothing is merged afterwards), so
 * alns / ops may hold unused gaps between segments: n_alns / n_ops count what is in use, *_extent is one past the last
 * used element.  Read r owns alns[read_aln_first[r] .. + read_aln_count[r]) in output order, r in the caller's order. */
typedef struct tg_result_c {
  uint32_t n_reads;
  uint32_t n_segments;
  uint64_t n_alns, n_ops;
  uint64_t alns_extent, ops_extent;
  const uint32_t* read_aln_first; /* [n_reads] */
  const uint32_t* read_aln_count; /* [n_reads] */
  const tg_aln_c* alns;           /* [alns_extent] */
  const uint32_t* ops;            /* [ops_extent] */
  uint64_t swg_cells, swg_extensions, seed_hits, n_smems;  /* as in tg_result */
} tg_result_c;

/* tg_align_batch with compact records (HOST buffers in, HOST result out, owned by ctx until the next call). */
tg_status tg_align_batch_compact(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result_c* out);
/* One compact record -> the wide record.  read_len = length of the read the record belongs to.  The wide record's
 * ops_off / tx_ops_off index the same `ops` array. */
tg_status tg_aln_expand(const tg_index_host* ix, const tg_aln_c* c, uint32_t read_len, tg_aln* out);
/* All records of a compact result -> wide records (alns_out[i] for i < alns_extent; gaps are zeroed) and u64 firsts
 * (first_out[n_reads], may be NULL), on n_threads host threads (0 = all cores).  offs = the offsets given to the align call. */
tg_status tg_result_expand(const tg_index_host* ix, const tg_result_c* res, const uint64_t* offs, tg_aln* alns_out,
                           uint64_t* first_out, int n_threads);

/* ---------------------------------------------------------------------------------------------------
 * Several GPUs of one host behind one call: the reference's driver is one process that writes its output in read order
 * (src/main.rs:45-83 -> src/aligner.rs:22-120, serial loop :54-115).  A tg_multi holds one index replica, one context and
 * one host thread per device.  tg_multi_create uploads the index to devices[0] and replicates it to the other GPUs with
 * ONE NCCL broadcast over NVLink (libnccl.so.2 is loaded at run time; without it: peer-to-peer copies) -- the only
 * collective on the path.  tg_multi_align_batch cuts the batch into contiguous shards [g*N/G, (g+1)*N/G),
 * runs them concurrently and returns ONE result in the caller's read order.
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_multi tg_multi;
tg_status tg_multi_create(const tg_index_host* ix, const int* devices, int n_devices, const tg_opts* opts, tg_multi** out);
void tg_multi_destroy(tg_multi* m);
int tg_multi_n_devices(const tg_multi* m);
/* Like tg_ctx_set_result_buffers: with n = 2 a result stays valid until the second next tg_multi_align_batch. */
tg_status tg_multi_set_result_buffers(tg_multi* m, int n);
/* "nccl" or "peer-copy": how the index replicas were made; *ms = device time of the broadcast (may be NULL). */
const char* tg_multi_replication(const tg_multi* m, float* ms);
/* Context of device slot g (e.g. for tg_ctx_set_exact_cell_count); NULL when g is out of range. */
tg_ctx* tg_multi_ctx(tg_multi* m, int g);
/* HOST buffers in, HOST result out (owned by m until the next call).  Records are identical to tg_align_batch_compact's
 * over the whole batch on one GPU. */
tg_status tg_multi_align_batch(tg_multi* m, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_result_c* out);
/* Timing of the last tg_multi_align_batch for device slot g: host wall time of the shard's call and the device times
 * tg_ctx_last_kernel_ms / tg_ctx_last_dp_ms report (any pointer may be NULL). */
tg_status tg_multi_last_timing(const tg_multi* m, int g, double* wall_ms, float* seed_ms, float* extend_ms, float* dp_ms);

/* ---------------------------------------------------------------------------------------------------
 * Per-read calls from many host threads: ThermiteAligner::align_read (src/wrapper.rs:20-27, :72).
 * The reference's embedding callers hold one clone per worker thread and align one read per call.  A tg_batcher lets
 * any number of threads do that against ONE context: reads are queued and a dispatcher thread runs tg_align_batch as
 * soon as max_batch_reads are waiting, the oldest queued read is max_wait_us old, or no new read has arrived for
 * max_wait_us / 4 (clamped to 20..100 us): blocking callers are served as soon as they have all resubmitted, callers
 * that stream tickets build large batches.  All tg_batcher_* calls except create / destroy are thread-safe; while the
 * batcher lives, `ctx` must not be used for anything else.  Records are identical to tg_align_batch's.
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_batcher tg_batcher;
typedef struct tg_read_alns {   /* the Vec<GenomeAlignment> of one align_read call (src/aligner.rs:123) */
  uint32_t n_alns;              /* 0 = unmapped */
  uint32_t n_ops;
  tg_aln* alns;                 /* [n_alns] in output order; ops_off / tx_ops_off index `ops` below */
  uint32_t* ops;                /* [n_ops] */
} tg_read_alns;
tg_status tg_batcher_create(tg_ctx* ctx, uint32_t max_batch_reads, uint32_t max_wait_us, tg_batcher** out);
/* Queue one read (the bytes are copied) and get a ticket; tg_batcher_wait blocks until that read's batch is done and
 * hands over its records (free with tg_read_alns_free).  A ticket is waited for exactly once. */
tg_status tg_batcher_submit(tg_batcher* b, const uint8_t* read, uint32_t len, uint64_t* ticket);
tg_status tg_batcher_wait(tg_batcher* b, uint64_t ticket, tg_read_alns* out);
/* submit + wait: the drop-in for a blocking align_read. */
tg_status tg_batcher_align_read(tg_batcher* b, const uint8_t* read, uint32_t len, tg_read_alns* out);
void tg_read_alns_free(tg_read_alns* r);
/* Reads served, batches run and the largest batch so far (any pointer may be NULL). */
tg_status tg_batcher_stats(tg_batcher* b, uint64_t* n_reads, uint64_t* n_batches, uint32_t* largest_batch);
/* Serves what is still queued, frees results nobody waited for.  No other call may be in flight on `b`. */
void tg_batcher_destroy(tg_batcher* b);

/* Seeding only: the SMEMs of every read in the order Index::all_smems returns them, one record per SMEM:
 * occurrences are suffix-array rows [sa_lo, sa_lo+count) visited from the LAST row to the first; when
 * `direct` is set the single occurrence position is sa_lo itself. */
typedef struct tg_seed {
  uint32_t query_idx, len, sa_lo, count;
  uint32_t direct, pad;
} tg_seed;
typedef struct tg_seed_result {
  uint32_t n_reads;
  uint64_t n_seeds;
  const uint64_t* read_seed_first; /* [n_reads] */
  const uint32_t* read_seed_count; /* [n_reads] */
  const tg_seed* seeds;
} tg_seed_result;
tg_status tg_seed_batch(tg_ctx* ctx, const uint8_t* bases, const uint64_t* offs, uint32_t n_reads, tg_seed_result* out);

/* SwgExtend::extend for n independent (x, y) pairs with unit scoring (gap_open -1, gap_extend -1, match +1,
 * mismatch -1; src/aligner.rs:140).  Bytes are compared as given.  Requires x_drop[t] >= band_width[t] (with a
 * smaller x_drop the reference reads stale columns or panics; see DESIGN.md "Q4").  HOST buffers in/out.
 * ops_off has n+1 entries; task t's words are ops[ops_off[t], ops_off[t+1]).  *cells = DP cells visited. */
tg_status tg_swg_extend_batch(tg_ctx* ctx, const uint8_t* xs, const uint64_t* xoff, const uint8_t* ys,
                              const uint64_t* yoff, uint32_t n, const uint32_t* band_width, const int32_t* x_drop,
                              int32_t* score, uint32_t* xend, uint32_t* yend, uint64_t* ops_off, uint32_t* ops,
                              uint64_t ops_cap, uint64_t* cells, float* kernel_ms);

/* ---------------------------------------------------------------------------------------------------
 * Writers: text identical to the reference's PAF / SAM output for the records of one batch.
 * names/quals: concatenated with n+1 offsets (like bases).  The returned buffer is malloc'ed; free with tg_free.
 * ------------------------------------------------------------------------------------------------- */
tg_status tg_format_sam_header(const tg_index_host* ix, char** out, size_t* out_len);
tg_status tg_format_batch(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                          const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals,
                          const uint64_t* qual_offs, int sam, char** out, size_t* out_len);
/* BAM (OutputFormat::Bam, src/aligner.rs:41-47, 69-72, 98-101): the records of tg_format_sam_header / tg_format_batch's
 * SAM text encoded as BAM and BGZF-compressed.  tg_format_bam_header = magic, header text and reference list
 * (write_header + write_reference_sequences); tg_format_batch_bam = the alignment records of one batch, optionally
 * followed by the BGZF end-of-file block.  header ++ batches (the last with append_eof = 1) is a complete BAM file.
 * Compressed bytes and integer tag widths are encoder choices: parity is on the decoded records. */
tg_status tg_format_bam_header(const tg_index_host* ix, void** out, size_t* out_len);
tg_status tg_format_batch_bam(const tg_index_host* ix, const tg_result* res, const uint8_t* bases, const uint64_t* offs,
                              const uint8_t* names, const uint64_t* name_offs, const uint8_t* quals, const uint64_t* qual_offs,
                              int append_eof, void** out, size_t* out_len);
/* FASTQ text -> concatenated bases / names / quals with offsets (needletail::parse_fastx_file, src/aligner.rs:51-55).
 * All six outputs are malloc'ed; free each with tg_free. */
tg_status tg_parse_fastq(const char* text, size_t len, uint32_t* n_reads, uint8_t** bases, uint64_t** offs,
                         uint8_t** names, uint64_t** name_offs, uint8_t** quals, uint64_t** qual_offs);
void tg_free(void* p);

/* ---------------------------------------------------------------------------------------------------
 * Streaming ingest and the file-to-file driver (align_reads_from_file, src/aligner.rs:22-120).
 *
 * tg_fastq_reader = needletail::parse_fastx_file for FASTQ (src/aligner.rs:51-55): plain text, gzip (any number of
 * members) or BGZF, told apart by the first bytes; "-" reads stdin.  Instead of one record at a time it hands out
 * BATCHES: tg_fastq_next parses up to max_reads records on all host cores straight into buffers owned by the reader
 * (bases and offsets page-locked when a CUDA device is present) and returns n_reads == 0 at the end of the input.  A
 * batch stays valid until the THIRD next call (three buffer sets alternate: one being filled, one being aligned, one
 * being written).  BGZF blocks are inflated in parallel, plain gzip members by one thread.  Record rules as
 * tg_parse_fastq: four lines per record, blank lines between records skipped, '\r' dropped, a truncated last record
 * dropped.  tg_fastq_format: 0 plain, 1 gzip, 2 BGZF.
 * ------------------------------------------------------------------------------------------------- */
typedef struct tg_fastq_reader tg_fastq_reader;
typedef struct tg_read_batch {
  uint32_t n_reads;
  uint32_t pad;
  const uint8_t* bases;  const uint64_t* offs;       /* [n_reads + 1] */
  const uint8_t* names;  const uint64_t* name_offs;  /* header line without '@' */
  const uint8_t* quals;  const uint64_t* qual_offs;
} tg_read_batch;
tg_status tg_fastq_open(const char* path, tg_fastq_reader** out);
tg_status tg_fastq_next(tg_fastq_reader* r, uint32_t max_reads, tg_read_batch* out);
int tg_fastq_format(const tg_fastq_reader* r);
void tg_fastq_close(tg_fastq_reader* r);

/* PAF lines written on the GPU (PafEntry, src/aln_writer.rs:47-116; csrc/tg_paf.cu): the batch is aligned with its records left in
 * HBM and two kernels turn them into text, so only names go up and only text comes back; byte-identical to tg_format_batch.
 * `ctx` must live on `device` and is used by the formatter (one call at a time).  *text points into a page-locked buffer
 * owned by the formatter, valid until the second next call (two buffers alternate); counters (may be NULL) receives the
 * tg_result counters of the batch (its pointers are device pointers). */
typedef struct tg_paf tg_paf;
tg_status tg_paf_create(const tg_index_host* ix, tg_ctx* ctx, int device, tg_paf** out);
/* The same formatter writing SAM records instead (aln_to_sam_record / unmapped_sam_record, src/aln_writer.rs:118-253; one
 * thread per read runs csrc/tg_textfmt.h): the batch's qualities go up as well, every read prints (unmapped ones as flag 4);
 * byte-identical to tg_format_batch(sam = 1).  No header lines (tg_format_sam_header). */
tg_status tg_sam_create(const tg_index_host* ix, tg_ctx* ctx, int device, tg_paf** out);
tg_status tg_paf_align_batch(tg_paf* f, const tg_read_batch* batch, const char** text, size_t* text_len, tg_result* counters);
/* The same call returning while the text is still on its way home (its copy runs on a stream of the formatter's own, so the
 * next batch can be aligned meanwhile): *text must not be read before tg_paf_wait(f, *text) has returned (text = NULL waits
 * for every copy in flight).  tg_align_files hands the wait to its writer stage. */
tg_status tg_paf_align_batch_async(tg_paf* f, const tg_read_batch* batch, const char** text, size_t* text_len, tg_result* counters);
tg_status tg_paf_wait(tg_paf* f, const char* text);
void tg_paf_destroy(tg_paf* f);
/* The device a context lives on. */
int tg_ctx_device(const tg_ctx* ctx);

/* align_reads_from_file (src/aligner.rs:22-120): query files -> PAF (output_fmt 0), SAM (1) or BAM (2) at out_path
 * ("-" = stdout), records in input order.  Exactly one of ctx / multi is given (one GPU, or the reads of every batch
 * sharded over the GPUs of a tg_multi).  Three overlapped stages: reader (inflate + parse batch k + 1), aligner (batch k
 * on the GPU), writers (format batch k - 1 on all host cores and write it); no stage holds more than three batches of
 * batch_reads reads (0 = 1 Mi).  The call switches ctx / multi to two alternating result sets.  PAF and SAM on one GPU are written
 * on the device (tg_paf_* / tg_sam_create; TG_PAF_HOST=1 in the environment keeps the host writers); a read of more than
 * TG_MAX_READ_LEN bases ends the run with TG_ERR_INVALID naming it (the batches before it are in the output). */
typedef struct tg_file_stats {
  uint64_t n_reads, n_alns, n_batches, bytes_out;
  double read_ms, align_ms, write_ms;  /* busy time of each stage */
  double format_ms;                    /* the part of write_ms spent turning records into text */
  double wall_ms;
  uint64_t warm_reads;                 /* reads of the first two batches and the time until they were written: the context's */
  double warm_ms;                      /* buffers grow there; (n_reads - warm_reads) / (wall_ms - warm_ms) is the steady rate */
} tg_file_stats;
tg_status tg_align_files(const tg_index_host* ix, tg_ctx* ctx, tg_multi* multi, const char* const* query_paths, int n_paths,
                         const char* out_path, int output_fmt, uint32_t batch_reads, tg_file_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* THERMITE_GPU_H */
