//! ffi.rs -- `extern "C"` declarations of include/thermite_gpu.h for the Rust host side.
//! UNCOMPILED here (no cargo/rustc in the image).  Field order and widths mirror the C header exactly.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub type tg_status = i32;
#[repr(C)] pub struct tg_index_host { _p: [u8; 0] }
#[repr(C)] pub struct tg_index { _p: [u8; 0] }
#[repr(C)] pub struct tg_ctx { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tg_opts {
    pub min_seed_len: u32,
    pub min_aln_score_percent: f32,
    pub min_aln_score: i32,
    pub multimap_score_range: u32,
    pub intron_mode: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tg_aln {
    pub ystart: u64, pub yend: u64, pub ylen: u64,
    pub tx_ystart: u64, pub tx_yend: u64, pub tx_ylen: u64,
    pub score: i32, pub ref_id: u32,
    pub xstart: u32, pub xend: u32, pub xlen: u32,
    pub tx_or_gene_idx: u32, pub tx_score: i32,
    pub tx_xstart: u32, pub tx_xend: u32,
    pub ops_off: u32, pub ops_len: u32, pub tx_ops_off: u32, pub tx_ops_len: u32,
    pub aln_type: u8, pub primary: u8, pub strand: u8, pub pad: u8,
}

#[repr(C)]
pub struct tg_result {
    pub n_reads: u32,
    pub n_alns: u64,
    pub n_ops: u64,
    pub read_aln_first: *const u64,
    pub read_aln_count: *const u32,
    pub alns: *const tg_aln,
    pub ops: *const u32,
    pub swg_cells: u64,
    pub swg_extensions: u64,
    pub seed_hits: u64,
    pub n_smems: u64,
}

extern "C" {
    pub fn tg_last_error() -> *const c_char;
    pub fn tg_opts_default(out: *mut tg_opts);
    pub fn tg_index_host_create_from_files(fasta: *const c_char, gtf: *const c_char, out: *mut *mut tg_index_host) -> tg_status;
    pub fn tg_index_host_save(ix: *const tg_index_host, path: *const c_char) -> tg_status;
    pub fn tg_index_host_load(path: *const c_char, out: *mut *mut tg_index_host) -> tg_status;
    pub fn tg_index_host_destroy(ix: *mut tg_index_host);
    pub fn tg_index_host_n_refs(ix: *const tg_index_host) -> u32;
    pub fn tg_index_host_ref(ix: *const tg_index_host, i: u32, out4: *mut u64) -> *const c_char;
    pub fn tg_index_host_tx(ix: *const tg_index_host, i: u32, out4: *mut u64) -> *const c_char;
    pub fn tg_index_host_gene_id(ix: *const tg_index_host, i: u32) -> *const c_char;
    pub fn tg_index_host_gene_name(ix: *const tg_index_host, i: u32) -> *const c_char;
    pub fn tg_index_create(ix: *const tg_index_host, device: c_int, out: *mut *mut tg_index) -> tg_status;
    pub fn tg_index_create_from_device_blob(blob: *const c_void, nbytes: usize, device: c_int, out: *mut *mut tg_index) -> tg_status;
    pub fn tg_index_destroy(ix: *mut tg_index);
    pub fn tg_ctx_create(ix: *const tg_index, opts: *const tg_opts, out: *mut *mut tg_ctx) -> tg_status;
    pub fn tg_ctx_destroy(ctx: *mut tg_ctx);
    pub fn tg_align_batch(ctx: *mut tg_ctx, bases: *const u8, offs: *const u64, n_reads: u32, out: *mut tg_result) -> tg_status;
    /// chunk size of the input copies inside tg_align_batch (copies overlap the seeding kernels)
    pub fn tg_ctx_set_chunk_reads(ctx: *mut tg_ctx, reads: u32);
    /// 1 = tg_result.swg_cells equals the reference's DP cell count (no early stop of extensions)
    pub fn tg_ctx_set_exact_cell_count(ctx: *mut tg_ctx, on: c_int);
    pub fn tg_ctx_last_kernel_ms(ctx: *const tg_ctx, seed_ms: *mut f32, extend_ms: *mut f32);
    pub fn tg_ctx_last_dp_ms(ctx: *const tg_ctx) -> f32;
    /// Index::create_from_files with divsufsort64 (src/index.rs:103-105) replaced by the GPU suffix-array builder
    pub fn tg_index_host_create_from_files_gpu(fasta: *const c_char, gtf: *const c_char, device: c_int, out: *mut *mut tg_index_host) -> tg_status;
    /// ThermiteAligner::align_read from many threads (src/wrapper.rs:20-27, :72): micro-batcher over one context
    pub fn tg_batcher_create(ctx: *mut tg_ctx, max_batch_reads: u32, max_wait_us: u32, out: *mut *mut tg_batcher) -> tg_status;
    pub fn tg_batcher_submit(b: *mut tg_batcher, read: *const u8, len: u32, ticket: *mut u64) -> tg_status;
    pub fn tg_batcher_wait(b: *mut tg_batcher, ticket: u64, out: *mut tg_read_alns) -> tg_status;
    pub fn tg_batcher_align_read(b: *mut tg_batcher, read: *const u8, len: u32, out: *mut tg_read_alns) -> tg_status;
    pub fn tg_read_alns_free(r: *mut tg_read_alns);
    pub fn tg_batcher_stats(b: *mut tg_batcher, n_reads: *mut u64, n_batches: *mut u64, largest_batch: *mut u32) -> tg_status;
    pub fn tg_batcher_destroy(b: *mut tg_batcher);
}

#[repr(C)]
pub struct tg_batcher { _private: [u8; 0] }

/// the Vec<GenomeAlignment> of one align_read call; `alns` and `ops` are one block owned by the caller until tg_read_alns_free
#[repr(C)]
pub struct tg_read_alns {
    pub n_alns: u32,
    pub n_ops: u32,
    pub alns: *mut tg_aln,
    pub ops: *mut u32,
}
