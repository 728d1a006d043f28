//! index_gpu.rs -- the crate's `Index` (src/index.rs:28-33) over libthermite_gpu.  UNCOMPILED here (no cargo in the image).
//!
//! The FM-index fields (`sa: SampledSuffixArray<..>`) disappear: seeding runs on the GPU against the flat blob.  What the
//! rest of the crate still reads from an `Index` -- `refs()` (src/index.rs:298-300: names, lengths, strands for PAF / SAM
//! headers and records) and `txome()` (src/index.rs:303-305: transcript ids, gene ids / names for the TX / GX / GN tags of
//! aln_writer.rs:179-213) -- is rebuilt once from the `tg_index_host_*` getters.  Exon tables and sequences stay inside the
//! blob: nothing on the Rust side uses them once align_seed_hit runs on the device.
use crate::ffi::*;
use crate::txome::{Gene, Tx, Txome};
use anyhow::{bail, Result};
use bio::data_structures::interval_tree::IntervalTree;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::sync::Mutex;

/// src/index.rs:391-399 without the sequence (it lives 4-bit packed in the blob).
#[derive(Debug)]
pub struct Ref {
    pub name: String,
    pub seq: Option<Vec<u8>>, // always None: kept so that code matching on it still compiles
    pub strand: bool,
    pub len: usize,
    pub start_idx: usize,
    pub end_idx: usize,
}

pub struct Index {
    host: *mut tg_index_host,
    refs: Vec<Ref>,
    txome: Txome,
    device: Mutex<HashMap<i32, usize>>, // device ordinal -> *mut tg_index (one replica per GPU)
}
unsafe impl Send for Index {}
unsafe impl Sync for Index {}

fn check(st: tg_status) -> Result<()> {
    if st == TG_OK { return Ok(()); }
    bail!("thermite_gpu: {}", unsafe { CStr::from_ptr(tg_last_error()) }.to_string_lossy())
}
unsafe fn cstr(p: *const std::os::raw::c_char) -> String { CStr::from_ptr(p).to_string_lossy().into_owned() }

impl Index {
    fn from_host(host: *mut tg_index_host) -> Index {
        unsafe {
            let mut refs = Vec::new();
            for i in 0..tg_index_host_n_refs(host) {
                let mut v = [0u64; 4]; // { start_idx, end_idx, len, strand }
                let name = cstr(tg_index_host_ref(host, i, v.as_mut_ptr()));
                refs.push(Ref { name, seq: None, strand: v[3] != 0, len: v[2] as usize, start_idx: v[0] as usize, end_idx: v[1] as usize });
            }
            let genes = (0..tg_index_host_n_genes(host))
                .map(|g| Gene { id: cstr(tg_index_host_gene_id(host, g)), name: cstr(tg_index_host_gene_name(host, g)) })
                .collect();
            let mut txs = Vec::new();
            for t in 0..tg_index_host_n_txs(host) {
                let mut v = [0u64; 4]; // { gene_idx, strand, n_exons, seq_len }
                let id = cstr(tg_index_host_tx(host, t, v.as_mut_ptr()));
                txs.push(Tx { id, chrom: String::new(), strand: v[1] != 0, exons: Vec::new(), seq: Vec::new(), gene_idx: v[0] as usize });
            }
            let txome = Txome { genes, txs, exon_to_tx: IntervalTree::new(), gene_intervals: IntervalTree::new() };
            Index { host, refs, txome, device: Mutex::new(HashMap::new()) }
        }
    }
    /// src/index.rs:52-57.  `sa_sampling_rate` / `occ_sampling_rate` have no meaning here (full suffix array in HBM).
    /// `gpu`: build the suffix array on that device (tg_sa.cu); None = SA-IS on the host.  Same blob either way.
    pub fn create_from_files(ref_path: &str, annot_path: &str, gpu: Option<i32>) -> Result<Index> {
        let (r, a) = (CString::new(ref_path)?, CString::new(annot_path)?);
        let mut h = std::ptr::null_mut();
        unsafe {
            match gpu {
                Some(d) => check(tg_index_host_create_from_files_gpu(r.as_ptr(), a.as_ptr(), d, &mut h))?,
                None => check(tg_index_host_create_from_files(r.as_ptr(), a.as_ptr(), &mut h))?,
            }
        }
        Ok(Index::from_host(h))
    }
    /// Replaces bincode::serialize_into / deserialize_from of src/main.rs:37-43, 63-67 (the file is the flat blob).
    pub fn save(&self, path: &str) -> Result<()> { check(unsafe { tg_index_host_save(self.host, CString::new(path)?.as_ptr()) }) }
    pub fn load(path: &str) -> Result<Index> {
        let mut h = std::ptr::null_mut();
        check(unsafe { tg_index_host_load(CString::new(path)?.as_ptr(), &mut h) })?;
        Ok(Index::from_host(h))
    }
    pub fn refs(&self) -> &[Ref] { &self.refs }
    pub fn txome(&self) -> &Txome { &self.txome }
    pub fn ref_names(&self) -> Vec<String> { self.refs.iter().map(|r| r.name.clone()).collect() }
    pub fn host(&self) -> *const tg_index_host { self.host }
    /// The replica of the index on one GPU (uploaded on first use).
    pub fn on_device(&self, device: i32) -> Result<*mut tg_index> {
        let mut m = self.device.lock().unwrap();
        if let Some(p) = m.get(&device) { return Ok(*p as *mut tg_index); }
        let mut d = std::ptr::null_mut();
        check(unsafe { tg_index_create(self.host, device, &mut d) })?;
        m.insert(device, d as usize);
        Ok(d)
    }
}
impl Drop for Index {
    fn drop(&mut self) {
        unsafe {
            for (_, p) in self.device.lock().unwrap().drain() { tg_index_destroy(p as *mut tg_index); }
            tg_index_host_destroy(self.host);
        }
    }
}

/// `thermite align` (src/main.rs:45-83 -> src/aligner.rs:22-120): the whole read loop runs inside the library
/// (tg_align_files: gz-transparent reader, aligner and writers overlapped), on one GPU or on all of `devices`.
pub fn align_reads_from_file(index: &Index, queries: &[String], output: &str, output_fmt: i32, opts: &tg_opts, devices: &[i32]) -> Result<tg_file_stats> {
    let paths: Vec<CString> = queries.iter().map(|q| CString::new(q.as_str()).unwrap()).collect();
    let ptrs: Vec<*const std::os::raw::c_char> = paths.iter().map(|p| p.as_ptr()).collect();
    let out = CString::new(output)?;
    let mut stats: tg_file_stats = unsafe { std::mem::zeroed() };
    unsafe {
        if devices.len() > 1 {
            let mut m = std::ptr::null_mut();
            check(tg_multi_create(index.host(), devices.as_ptr(), devices.len() as i32, opts, &mut m))?;
            let st = tg_align_files(index.host(), std::ptr::null_mut(), m, ptrs.as_ptr(), ptrs.len() as i32, out.as_ptr(), output_fmt, 0, &mut stats);
            tg_multi_destroy(m);
            check(st)?;
        } else {
            let mut ctx = std::ptr::null_mut();
            check(tg_ctx_create(index.on_device(*devices.first().unwrap_or(&0))?, opts, &mut ctx))?;
            let st = tg_align_files(index.host(), ctx, std::ptr::null_mut(), ptrs.as_ptr(), ptrs.len() as i32, out.as_ptr(), output_fmt, 0, &mut stats);
            tg_ctx_destroy(ctx);
            check(st)?;
        }
    }
    Ok(stats)
}
