//! aligner_gpu.rs -- how src/aligner.rs keeps its public signature while the work goes to the GPU.
//! UNCOMPILED here.  `align_read` stays (`index`, `read`, `opts`) -> Vec<GenomeAlignment>; a batched entry point
//! amortises the launch and is what align_reads_from_file (src/aligner.rs:22-120) calls per FASTQ chunk.
use crate::ffi::*;
use crate::txome::{AlnType, GenomeAlignment};
use bio::alignment::{Alignment, AlignmentMode, AlignmentOperation};

fn expand_ops(words: &[u32]) -> Vec<AlignmentOperation> {
    let mut v = Vec::new();
    for &w in words {
        let (kind, run) = (w & 7, (w >> 3) as usize);
        match kind {
            0 => v.extend(std::iter::repeat(AlignmentOperation::Match).take(run)),
            1 => v.extend(std::iter::repeat(AlignmentOperation::Subst).take(run)),
            2 => v.extend(std::iter::repeat(AlignmentOperation::Del).take(run)),
            3 => v.extend(std::iter::repeat(AlignmentOperation::Ins).take(run)),
            4 => v.push(AlignmentOperation::Xclip(run)),
            _ => v.push(AlignmentOperation::Yclip(run)),
        }
    }
    v
}

/// Batched align_read: reads[i] -> Vec<GenomeAlignment>, identical to calling the CPU align_read per read.
pub unsafe fn align_reads_gpu(ctx: *mut tg_ctx, ref_names: &[String], reads: &[&[u8]]) -> anyhow::Result<Vec<Vec<GenomeAlignment>>> {
    let mut bases = Vec::new();
    let mut offs = vec![0u64];
    for r in reads { bases.extend_from_slice(r); offs.push(bases.len() as u64); }
    let mut res: tg_result = std::mem::zeroed();
    let st = tg_align_batch(ctx, bases.as_ptr(), offs.as_ptr(), reads.len() as u32, &mut res);
    if st != 0 { anyhow::bail!("thermite_gpu: {}", std::ffi::CStr::from_ptr(tg_last_error()).to_string_lossy()); }
    let alns = std::slice::from_raw_parts(res.alns, res.n_alns as usize);
    let ops = std::slice::from_raw_parts(res.ops, res.n_ops as usize);
    let first = std::slice::from_raw_parts(res.read_aln_first, reads.len());
    let count = std::slice::from_raw_parts(res.read_aln_count, reads.len());
    let mut out = Vec::with_capacity(reads.len());
    for r in 0..reads.len() {
        let mut v = Vec::with_capacity(count[r] as usize);
        for a in &alns[first[r] as usize..first[r] as usize + count[r] as usize] {
            let gx = Alignment { score: a.score, ystart: a.ystart as usize, xstart: a.xstart as usize, yend: a.yend as usize,
                xend: a.xend as usize, ylen: a.ylen as usize, xlen: a.xlen as usize,
                operations: expand_ops(&ops[a.ops_off as usize..(a.ops_off + a.ops_len) as usize]), mode: AlignmentMode::Custom };
            let aln_type = match a.aln_type {
                0 => AlnType::Exonic { tx_idx: a.tx_or_gene_idx as usize, tx_aln: Alignment { score: a.tx_score,
                        ystart: a.tx_ystart as usize, xstart: a.tx_xstart as usize, yend: a.tx_yend as usize, xend: a.tx_xend as usize,
                        ylen: a.tx_ylen as usize, xlen: a.xlen as usize,
                        operations: expand_ops(&ops[a.tx_ops_off as usize..(a.tx_ops_off + a.tx_ops_len) as usize]), mode: AlignmentMode::Custom } },
                1 => AlnType::Intronic { gene_idx: a.tx_or_gene_idx as usize },
                _ => AlnType::Intergenic,
            };
            v.push(GenomeAlignment { gx_aln: gx, aln_type, ref_name: ref_names[a.ref_id as usize].clone(),
                strand: a.strand != 0, primary: a.primary != 0 });
        }
        out.push(v);
    }
    Ok(out)
}
