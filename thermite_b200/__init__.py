"""thermite_b200 -- B200 (sm_100a) implementation of thermite's read-alignment hot path behind the
reference's aligner API.  Compute lives in csrc/libthermite_gpu.so (hand-written CUDA); this package is
the host-side mirror of the reference interface.  See DESIGN.md / INTEGRATION.md."""
from .api import (ALN_C_DTYPE, ALN_DTYPE, SEED_DTYPE, MultiAligner, compact_arrays, expand_result, AlignOpts, Aligner, AlignResult, Alignment, GenomeAlignment, Index,
                  OutputFormat, FastqReader, ThermiteAligner, ThermiteError, align_reads_from_file, bam_header, expand_ops, lib, parse_fastq, sam_header, suffix_array_gpu)

__all__ = ["ALN_C_DTYPE", "MultiAligner", "compact_arrays", "expand_result", "ALN_DTYPE", "SEED_DTYPE", "AlignOpts", "Aligner", "AlignResult", "Alignment", "GenomeAlignment", "Index",
           "OutputFormat", "FastqReader", "ThermiteAligner", "ThermiteError", "align_reads_from_file", "bam_header", "expand_ops", "lib", "parse_fastq", "sam_header", "suffix_array_gpu"]
