"""Host-side mirror of thermite's aligner API over the C ABI of libthermite_gpu.so (include/thermite_gpu.h).

Names follow the reference crate (paths under /root/reference):
  Index.create_from_files / refs / txome      src/index.rs:52-57, 293-300
  AlignOpts                                   src/aligner.rs:452-464 (defaults src/main.rs:115-132)
  Aligner.align_reads (batched align_read)    src/aligner.rs:123-190
  GenomeAlignment / AlnType                   src/txome.rs:55-69
  align_reads_from_file                       src/aligner.rs:22-120
  OutputFormat                                src/aln_writer.rs:16-21

The CUDA library is the only compute path: importing works without a GPU (so the host-side logic and the
symbol table can be tested), but every device call raises ThermiteError when the library or a device is
missing.  Nothing here falls back to a CPU implementation.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.environ.get("THERMITE_GPU_LIB", os.path.join(_CSRC, "libthermite_gpu.so"))
_LIB = None

TG_MAX_READ_LEN = 512

ALN_DTYPE = np.dtype([
    ("ystart", "<u8"), ("yend", "<u8"), ("ylen", "<u8"),
    ("tx_ystart", "<u8"), ("tx_yend", "<u8"), ("tx_ylen", "<u8"),
    ("score", "<i4"), ("ref_id", "<u4"),
    ("xstart", "<u4"), ("xend", "<u4"), ("xlen", "<u4"),
    ("tx_or_gene_idx", "<u4"), ("tx_score", "<i4"),
    ("tx_xstart", "<u4"), ("tx_xend", "<u4"),
    ("ops_off", "<u4"), ("ops_len", "<u4"), ("tx_ops_off", "<u4"), ("tx_ops_len", "<u4"),
    ("aln_type", "u1"), ("primary", "u1"), ("strand", "u1"), ("pad", "u1"),
])
# tg_aln_c (40 B): the same record without what refs() / txome() / the read already say
ALN_C_DTYPE = np.dtype([
    ("ystart", "<u4"), ("yend", "<u4"), ("tx_ystart", "<u4"), ("tx_yend", "<u4"),
    ("ref_id", "<u4"), ("tx_or_gene_idx", "<u4"), ("ops_off", "<u4"),
    ("score", "<i2"), ("xstart", "<u2"), ("xend", "<u2"), ("ops_len", "<u2"), ("tx_ops_len", "<u2"),
    ("aln_type", "u1"), ("primary", "u1"),
])
SEED_DTYPE = np.dtype([("query_idx", "<u4"), ("len", "<u4"), ("sa_lo", "<u4"), ("count", "<u4"),
                       ("direct", "<u4"), ("pad", "<u4")])
OP_NAMES = ["Match", "Subst", "Del", "Ins", "Xclip", "Yclip"]
ALN_TYPES = ["Exonic", "Intronic", "Intergenic"]


class ThermiteError(RuntimeError):
    pass


class _Opts(C.Structure):
    _fields_ = [("min_seed_len", C.c_uint32), ("min_aln_score_percent", C.c_float), ("min_aln_score", C.c_int32),
                ("multimap_score_range", C.c_uint32), ("intron_mode", C.c_uint32)]


class _Result(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_alns", C.c_uint64), ("n_ops", C.c_uint64),
                ("read_aln_first", C.c_void_p), ("read_aln_count", C.c_void_p), ("alns", C.c_void_p),
                ("ops", C.c_void_p), ("swg_cells", C.c_uint64), ("swg_extensions", C.c_uint64),
                ("seed_hits", C.c_uint64), ("n_smems", C.c_uint64)]


class _ResultC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_segments", C.c_uint32), ("n_alns", C.c_uint64), ("n_ops", C.c_uint64),
                ("alns_extent", C.c_uint64), ("ops_extent", C.c_uint64),
                ("read_aln_first", C.c_void_p), ("read_aln_count", C.c_void_p), ("alns", C.c_void_p),
                ("ops", C.c_void_p), ("swg_cells", C.c_uint64), ("swg_extensions", C.c_uint64),
                ("seed_hits", C.c_uint64), ("n_smems", C.c_uint64)]


class _SeedResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_seeds", C.c_uint64), ("read_seed_first", C.c_void_p),
                ("read_seed_count", C.c_void_p), ("seeds", C.c_void_p)]


# every symbol include/thermite_gpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "tg_last_error", "tg_version", "tg_opts_default",
    "tg_index_host_create_from_files", "tg_index_host_create_from_memory", "tg_index_host_blob",
    "tg_index_host_from_blob", "tg_index_host_save", "tg_index_host_load", "tg_index_host_destroy",
    "tg_index_host_text_len", "tg_index_host_n_refs", "tg_index_host_n_txs", "tg_index_host_n_genes",
    "tg_index_host_ref", "tg_index_host_tx", "tg_index_host_gene_id", "tg_index_host_gene_name", "tg_index_host_sa",
    "tg_format_bam_header", "tg_format_batch_bam",
    "tg_batcher_create", "tg_batcher_submit", "tg_batcher_wait", "tg_batcher_align_read", "tg_read_alns_free",
    "tg_batcher_stats", "tg_batcher_destroy",
    "tg_index_host_text4", "tg_index_host_create_from_files_gpu", "tg_index_host_create_from_memory_gpu", "tg_suffix_array_gpu",
    "tg_index_create", "tg_index_create_from_device_blob", "tg_index_destroy",
    "tg_ctx_create", "tg_ctx_destroy", "tg_ctx_set_chunk_reads", "tg_ctx_set_result_buffers", "tg_ctx_stream", "tg_ctx_last_kernel_ms", "tg_ctx_last_kernel_launches", "tg_ctx_last_dp_ms", "tg_bench_random_gather", "tg_bench_int_peak", "tg_ctx_kmer_table_bytes", "tg_ctx_set_exact_cell_count", "tg_ctx_set_round_pipeline",
    "tg_align_batch", "tg_align_batch_device", "tg_seed_batch", "tg_swg_extend_batch",
    "tg_format_sam_header", "tg_format_batch", "tg_parse_fastq", "tg_free",
    "tg_align_batch_compact", "tg_aln_expand", "tg_result_expand",
    "tg_multi_create", "tg_multi_destroy", "tg_multi_set_result_buffers", "tg_multi_n_devices", "tg_multi_replication", "tg_multi_ctx",
    "tg_multi_align_batch", "tg_multi_last_timing",
    "tg_fastq_open", "tg_fastq_next", "tg_fastq_format", "tg_fastq_close", "tg_align_files",
    "tg_paf_create", "tg_sam_create", "tg_paf_align_batch", "tg_paf_align_batch_async", "tg_paf_wait", "tg_paf_destroy", "tg_ctx_device",
]


def lib():
    """Load libthermite_gpu.so (built by __graft_entry__.build() / `make -C thermite_b200/csrc`)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ThermiteError(f"{LIB_PATH} is missing: build it with `make -C thermite_b200/csrc` "
                                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tg_last_error.restype = C.c_char_p
        L.tg_version.restype = C.c_char_p
        L.tg_index_host_text_len.restype = C.c_uint64
        for f in ("tg_index_host_ref", "tg_index_host_tx", "tg_index_host_gene_id", "tg_index_host_gene_name"):
            getattr(L, f).restype = C.c_char_p
        L.tg_index_host_sa.restype = C.c_void_p
        L.tg_index_host_text4.restype = C.c_void_p
        L.tg_batcher_destroy.restype = None
        L.tg_read_alns_free.restype = None
        L.tg_ctx_stream.restype = C.c_void_p
        L.tg_ctx_kmer_table_bytes.restype = C.c_uint64
        L.tg_ctx_last_kernel_launches.restype = C.c_uint64
        L.tg_ctx_last_dp_ms.restype = C.c_float
        L.tg_ctx_last_kernel_ms.restype = None
        L.tg_ctx_set_exact_cell_count.restype = None
        L.tg_ctx_set_round_pipeline.restype = None
        L.tg_ctx_set_chunk_reads.restype = None
        L.tg_free.restype = None
        L.tg_multi_destroy.restype = None
        L.tg_fastq_close.restype = None
        L.tg_paf_destroy.restype = None
        L.tg_multi_replication.restype = C.c_char_p
        L.tg_multi_ctx.restype = C.c_void_p
        _LIB = L
    return _LIB


def _check(st):
    if st != 0:
        raise ThermiteError(f"thermite_gpu error {st}: {lib().tg_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def expand_ops(words):
    """RLE words -> [(kind_name, n)], one entry per bio AlignmentOperation."""
    out = []
    for w in np.asarray(words, dtype=np.uint32).tolist():
        k, n = w & 7, w >> 3
        if k <= 3:
            out.extend([(OP_NAMES[k], 1)] * n)
        else:
            out.append((OP_NAMES[k], n))
    return out


@dataclass
class AlignOpts:
    """src/aligner.rs:452-464; defaults = src/main.rs:115-132."""
    min_seed_len: int = 20
    min_aln_score_percent: float = 0.66
    min_aln_score: int = 30
    multimap_score_range: int = 1
    intron_mode: bool = False

    def _c(self):
        return _Opts(self.min_seed_len, self.min_aln_score_percent, self.min_aln_score, self.multimap_score_range,
                     int(self.intron_mode))


@dataclass
class Ref:
    """src/index.rs:391-399 (the sequence itself stays in the packed index)."""
    name: str
    strand: bool
    len: int
    start_idx: int
    end_idx: int


@dataclass
class Tx:
    id: str
    strand: bool
    gene_idx: int
    n_exons: int
    seq_len: int


@dataclass
class Gene:
    id: str
    name: str


class Txome:
    def __init__(self, genes, txs):
        self.genes, self.txs = genes, txs


def suffix_array_gpu(text4: np.ndarray, text_len: int, device: int = 0):
    """Suffix array of a packed text on the GPU (csrc/tg_sa.cu; divsufsort64 at src/index.rs:103-105).
    `text4` needs at least one zero word behind the text.  Returns (sa, device_ms, sort_steps)."""
    text4 = np.ascontiguousarray(text4, np.uint64)
    if len(text4) < text_len // 16 + 2:
        raise ThermiteError("text4 needs a zero word behind the text")
    if len(text4) < text_len // 16 + 4:
        text4 = np.concatenate([text4, np.zeros(4, np.uint64)])
    sa = np.empty(text_len, np.uint32)
    ms, steps = C.c_float(0), C.c_uint32(0)
    _check(lib().tg_suffix_array_gpu(_p(text4), C.c_uint64(text_len), C.c_int(device), _p(sa), C.byref(ms), C.byref(steps)))
    return sa, ms.value, steps.value


class Index:
    """Flat host index + (lazily) its HBM replica.  src/index.rs:40-44."""

    def __init__(self, handle):
        self._h = handle
        self._dev = {}
        self._blob_keepalive = None

    # -- construction ---------------------------------------------------------------------------------------
    @classmethod
    def create_from_files(cls, ref_path: str, annot_path: str, sa_sampling_rate: int = 32,
                          occ_sampling_rate: int = 128, sa_device: Optional[int] = None) -> "Index":
        """src/index.rs:52-57.  The sampling rates are accepted for signature parity and ignored: the GPU
        index keeps the full suffix array and has no Occ table.  `sa_device` = GPU that builds the suffix array
        (divsufsort64 in the reference, src/index.rs:103-105); None = SA-IS on the host.  Same blob either way."""
        h = C.c_void_p()
        if sa_device is None:
            _check(lib().tg_index_host_create_from_files(ref_path.encode(), annot_path.encode(), C.byref(h)))
        else:
            _check(lib().tg_index_host_create_from_files_gpu(ref_path.encode(), annot_path.encode(),
                                                             C.c_int(sa_device), C.byref(h)))
        return cls(h)

    @classmethod
    def create_from_memory(cls, fasta_text: bytes, gtf_text: bytes, sa_device: Optional[int] = None) -> "Index":
        h = C.c_void_p()
        if sa_device is None:
            _check(lib().tg_index_host_create_from_memory(fasta_text, C.c_size_t(len(fasta_text)), gtf_text,
                                                          C.c_size_t(len(gtf_text)), C.byref(h)))
        else:
            _check(lib().tg_index_host_create_from_memory_gpu(fasta_text, C.c_size_t(len(fasta_text)), gtf_text,
                                                              C.c_size_t(len(gtf_text)), C.c_int(sa_device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_blob(cls, blob: np.ndarray) -> "Index":
        blob = np.ascontiguousarray(blob, np.uint8)
        h = C.c_void_p()
        _check(lib().tg_index_host_from_blob(_p(blob), C.c_size_t(blob.nbytes), C.byref(h)))
        return cls(h)

    def save(self, path: str):
        """The `.tai` of this implementation (src/main.rs:37-43)."""
        _check(lib().tg_index_host_save(self._h, path.encode()))

    @classmethod
    def load(cls, path: str) -> "Index":
        h = C.c_void_p()
        _check(lib().tg_index_host_load(path.encode(), C.byref(h)))
        return cls(h)

    def blob(self) -> np.ndarray:
        """Zero-copy view of the position-independent index blob."""
        d, n = C.c_void_p(), C.c_size_t()
        _check(lib().tg_index_host_blob(self._h, C.byref(d), C.byref(n)))
        return np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint8)), shape=(n.value,))

    def __del__(self):
        try:
            for d in self._dev.values():
                lib().tg_index_destroy(d)
            lib().tg_index_host_destroy(self._h)
        except Exception:
            pass

    # -- accessors (src/index.rs:293-300) ---------------------------------------------------------------------
    def text_len(self) -> int:
        return lib().tg_index_host_text_len(self._h)

    def refs(self) -> List[Ref]:
        out = []
        for i in range(lib().tg_index_host_n_refs(self._h)):
            v = (C.c_uint64 * 4)()
            name = lib().tg_index_host_ref(self._h, i, v).decode()
            out.append(Ref(name, bool(v[3]), int(v[2]), int(v[0]), int(v[1])))
        return out

    def txome(self) -> Txome:
        L = lib()
        genes = [Gene(L.tg_index_host_gene_id(self._h, i).decode(), L.tg_index_host_gene_name(self._h, i).decode())
                 for i in range(L.tg_index_host_n_genes(self._h))]
        txs = []
        for i in range(L.tg_index_host_n_txs(self._h)):
            v = (C.c_uint64 * 4)()
            tid = L.tg_index_host_tx(self._h, i, v).decode()
            txs.append(Tx(tid, bool(v[1]), int(v[0]), int(v[2]), int(v[3])))
        return Txome(genes, txs)

    def suffix_array(self) -> np.ndarray:
        n = self.text_len()
        p = lib().tg_index_host_sa(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,))

    def text4(self) -> np.ndarray:
        """The packed both-strand text: 4-bit codes ($ACGNT = 0..5), 16 per u64, first symbol in the top nibble."""
        n = self.text_len()
        p = lib().tg_index_host_text4(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n // 16 + 4,))

    # -- device replica -----------------------------------------------------------------------------------------
    def device_index(self, device: int = 0):
        if device not in self._dev:
            d = C.c_void_p()
            _check(lib().tg_index_create(self._h, device, C.byref(d)))
            self._dev[device] = d
        return self._dev[device]

    def adopt_device_blob(self, device_ptr: int, nbytes: int, device: int, keepalive=None):
        """Use a blob that is already in HBM (e.g. received by an NCCL broadcast) as this index's replica."""
        d = C.c_void_p()
        _check(lib().tg_index_create_from_device_blob(C.c_void_p(device_ptr), C.c_size_t(nbytes), device, C.byref(d)))
        self._dev[device] = d
        self._blob_keepalive = keepalive
        return d


@dataclass
class Alignment:
    """bio::alignment::Alignment fields used on this path."""
    score: int
    ystart: int
    xstart: int
    yend: int
    xend: int
    ylen: int
    xlen: int
    operations: list


@dataclass
class GenomeAlignment:
    """src/txome.rs:55-69"""
    gx_aln: Alignment
    aln_type: str
    ref_name: str
    strand: bool
    primary: bool
    tx_aln: Optional[Alignment] = None
    tx_idx: Optional[int] = None
    gene_idx: Optional[int] = None


class AlignResult:
    """Flat records of one batch (host copies).  read r owns alns[first[r] : first[r] + count[r]]."""

    def __init__(self, first, count, alns, ops, counters, ref_names=None):
        self.first, self.count, self.alns, self.ops = first, count, alns, ops
        self.counters = counters
        self._ref_names = ref_names

    def __len__(self):
        return len(self.first)

    def read_alignments(self, r: int) -> List[GenomeAlignment]:
        out = []
        for a in self.alns[int(self.first[r]): int(self.first[r]) + int(self.count[r])]:
            ops = expand_ops(self.ops[int(a["ops_off"]): int(a["ops_off"]) + int(a["ops_len"])])
            gx = Alignment(int(a["score"]), int(a["ystart"]), int(a["xstart"]), int(a["yend"]), int(a["xend"]),
                           int(a["ylen"]), int(a["xlen"]), ops)
            g = GenomeAlignment(gx, ALN_TYPES[int(a["aln_type"])],
                                self._ref_names[int(a["ref_id"])] if self._ref_names else str(int(a["ref_id"])),
                                bool(a["strand"]), bool(a["primary"]))
            if a["aln_type"] == 0:
                tops = expand_ops(self.ops[int(a["tx_ops_off"]): int(a["tx_ops_off"]) + int(a["tx_ops_len"])])
                g.tx_aln = Alignment(int(a["tx_score"]), int(a["tx_ystart"]), int(a["tx_xstart"]), int(a["tx_yend"]),
                                     int(a["tx_xend"]), int(a["tx_ylen"]), int(a["xlen"]), tops)
                g.tx_idx = int(a["tx_or_gene_idx"])
            elif a["aln_type"] == 1:
                g.gene_idx = int(a["tx_or_gene_idx"])
            out.append(g)
        return out


class Aligner:
    """One alignment context (CUDA stream + k-mer table + scratch) on one GPU.  Not thread-safe; create one per
    host thread / per GPU (the reference's ThermiteAligner is Clone + Send, src/wrapper.rs:20-27)."""

    def __init__(self, index: Index, opts: AlignOpts = None, device: int = 0):
        self.index = index
        self.opts = opts or AlignOpts()
        self.device = device
        dix = index.device_index(device)
        h = C.c_void_p()
        o = self.opts._c()
        _check(lib().tg_ctx_create(dix, C.byref(o), C.byref(h)))
        self._h = h
        self._ref_names = [r.name for r in index.refs()]

    def close(self):
        """Destroy the context (its k-mer table stays while another context of the index uses it)."""
        if getattr(self, "_owned", True) and getattr(self, "_h", None):
            lib().tg_ctx_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def _view(cls, handle, index, opts, device, ref_names):
        """An Aligner over a context somebody else owns (MultiAligner.context)."""
        a = cls.__new__(cls)
        a._h, a.index, a.opts, a.device, a._ref_names, a._owned = C.c_void_p(handle), index, opts, device, ref_names, False
        return a

    def stream_ptr(self) -> int:
        return lib().tg_ctx_stream(self._h)

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        lib().tg_ctx_last_kernel_ms(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_exact_cell_count(self, on: bool):
        """True: visit every DP column the reference visits (counters match it); False (default): stop extensions early
        when provably nothing later can win -- identical records, fewer cells."""
        lib().tg_ctx_set_exact_cell_count(self._h, int(on))

    def set_round_pipeline(self, on: bool):
        """True (default): round pipeline (thread-per-read control + warp-per-task extension); False: single-warp kernel."""
        lib().tg_ctx_set_round_pipeline(self._h, int(on))

    def set_chunk_reads(self, reads: int):
        """Chunk size of align_reads on host buffers (copies of one chunk overlap the kernels of the next)."""
        lib().tg_ctx_set_chunk_reads(self._h, C.c_uint32(reads))

    def last_dp_ms(self) -> float:
        """Device time of the banded-SWG kernels in the last align call (part of the extend time)."""
        return float(lib().tg_ctx_last_dp_ms(self._h))

    def random_gather_gbs(self, n_loads: int = 1 << 26, reps: int = 5) -> float:
        """HBM random-access yardstick: GB/s of 32-B sectors for independent random 16-B loads over the k-mer table."""
        g, ms = C.c_double(), C.c_float()
        _check(lib().tg_bench_random_gather(self._h, C.c_uint64(n_loads), reps, C.byref(g), C.byref(ms)))
        return float(g.value)

    def int_peak_lane_ops(self, reps: int = 5):
        """ALU-pipe yardstick: (VIADDMNMX, VIMNMX3) lane-operations per second, dependency-free streams."""
        a, b = C.c_double(), C.c_double()
        _check(lib().tg_bench_int_peak(self._h, reps, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def last_kernel_launches(self) -> int:
        return int(lib().tg_ctx_last_kernel_launches(self._h))

    def kmer_table_bytes(self) -> int:
        return lib().tg_ctx_kmer_table_bytes(self._h)

    @staticmethod
    def _counters(res):
        return dict(swg_cells=res.swg_cells, swg_extensions=res.swg_extensions, seed_hits=res.seed_hits,
                    n_smems=res.n_smems, n_alns=res.n_alns, n_ops=res.n_ops)

    def align_reads_raw(self, bases_ptr: int, offs_ptr: int, n_reads: int) -> _Result:
        """tg_align_batch on raw HOST pointers; returns the C struct (pointers owned by the context)."""
        res = _Result()
        _check(lib().tg_align_batch(self._h, C.c_void_p(bases_ptr), C.c_void_p(offs_ptr), n_reads, C.byref(res)))
        return res

    def align_reads_device_raw(self, d_bases_ptr: int, d_offs_ptr: int, n_reads: int, total_bases: int,
                               max_read_len: int) -> _Result:
        """tg_align_batch_device: inputs and results stay in HBM."""
        res = _Result()
        _check(lib().tg_align_batch_device(self._h, C.c_void_p(d_bases_ptr), C.c_void_p(d_offs_ptr), n_reads,
                                           C.c_uint64(total_bases), max_read_len, C.byref(res)))
        return res

    def align_reads(self, bases, offs) -> AlignResult:
        """Batched align_read (src/aligner.rs:123-190).  bases: concatenated ASCII reads; offs: uint64[n+1]."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        res = self.align_reads_raw(bases.ctypes.data, offs.ctypes.data, n)
        first = np.ctypeslib.as_array(C.cast(res.read_aln_first, C.POINTER(C.c_uint64)), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        count = np.ctypeslib.as_array(C.cast(res.read_aln_count, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        alns = np.zeros(res.n_alns, ALN_DTYPE)
        if res.n_alns:
            C.memmove(alns.ctypes.data, res.alns, res.n_alns * ALN_DTYPE.itemsize)
        ops = np.zeros(res.n_ops, np.uint32)
        if res.n_ops:
            C.memmove(ops.ctypes.data, res.ops, res.n_ops * 4)
        return AlignResult(first, count, alns, ops, self._counters(res), self._ref_names)

    def align_reads_compact_raw(self, bases_ptr: int, offs_ptr: int, n_reads: int) -> _ResultC:
        """tg_align_batch_compact on raw HOST pointers: 40-byte records (tg_aln_c), u32 first indices."""
        res = _ResultC()
        _check(lib().tg_align_batch_compact(self._h, C.c_void_p(bases_ptr), C.c_void_p(offs_ptr), n_reads, C.byref(res)))
        return res

    def align_reads_compact(self, bases, offs) -> "AlignResult":
        """align_reads through the compact call; the records are expanded back to the wide form (tg_result_expand)."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        res = self.align_reads_compact_raw(bases.ctypes.data, offs.ctypes.data, n)
        return expand_result(self.index, res, offs, self._ref_names)

    def align_read(self, read: bytes) -> List[GenomeAlignment]:
        """align_read for one read (src/aligner.rs:123)."""
        b = np.frombuffer(read, np.uint8)
        r = self.align_reads(b, np.array([0, len(b)], np.uint64))
        return r.read_alignments(0)

    def seed_reads(self, bases, offs):
        """Index::all_smems, batched (src/index.rs:228-255) -> (seeds, first, count)."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        res = _SeedResult()
        _check(lib().tg_seed_batch(self._h, _p(bases), _p(offs), n, C.byref(res)))
        first = np.ctypeslib.as_array(C.cast(res.read_seed_first, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        count = np.ctypeslib.as_array(C.cast(res.read_seed_count, C.POINTER(C.c_uint32)), shape=(n,)).copy()
        seeds = np.zeros(res.n_seeds, SEED_DTYPE)
        if res.n_seeds:
            C.memmove(seeds.ctypes.data, res.seeds, res.n_seeds * SEED_DTYPE.itemsize)
        return seeds, first, count

    def all_smems(self, read: bytes):
        """Index::all_smems for one read -> [(ref_idx, query_idx, len)] in the reference's order."""
        b = np.frombuffer(read, np.uint8)
        seeds, first, count = self.seed_reads(b, np.array([0, len(b)], np.uint64))
        sa = self.index.suffix_array()
        out = []
        for s in seeds[int(first[0]): int(first[0]) + int(count[0])]:
            if s["direct"]:
                out.append((int(s["sa_lo"]), int(s["query_idx"]), int(s["len"])))
            else:
                for rk in range(int(s["count"]) - 1, -1, -1):
                    out.append((int(sa[int(s["sa_lo"]) + rk]), int(s["query_idx"]), int(s["len"])))
        return out

    def swg_extend_batch(self, xs, xoff, ys, yoff, bw, x_drop):
        """SwgExtend::extend for independent pairs (src/swg.rs:31-207)."""
        n = len(bw)
        xs = np.ascontiguousarray(xs, np.uint8); ys = np.ascontiguousarray(ys, np.uint8)
        xoff = np.ascontiguousarray(xoff, np.uint64); yoff = np.ascontiguousarray(yoff, np.uint64)
        bw = np.ascontiguousarray(bw, np.uint32); x_drop = np.ascontiguousarray(x_drop, np.int32)
        score = np.zeros(n, np.int32); xend = np.zeros(n, np.uint32); yend = np.zeros(n, np.uint32)
        ops_off = np.zeros(n + 1, np.uint64)
        cap = int(len(xs) + len(ys) + 4 * n + 16)
        ops = np.zeros(cap, np.uint32)
        cells, ms = C.c_uint64(), C.c_float()
        _check(lib().tg_swg_extend_batch(self._h, _p(xs), _p(xoff), _p(ys), _p(yoff), n, _p(bw), _p(x_drop), _p(score),
                                         _p(xend), _p(yend), _p(ops_off), _p(ops), C.c_uint64(cap), C.byref(cells),
                                         C.byref(ms)))
        return dict(score=score, xend=xend, yend=yend, ops_off=ops_off, ops=ops[: int(ops_off[n])].copy(),
                    cells=cells.value, kernel_ms=ms.value, dp_ms=self.last_dp_ms())

    def format_result_raw(self, res: _Result, bases, offs, names, name_offs, quals, qual_offs, sam: bool) -> bytes:
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().tg_format_batch(self.index._h, C.byref(res), _p(bases), _p(offs), _p(names), _p(name_offs),
                                     _p(quals) if quals is not None else None,
                                     _p(qual_offs) if qual_offs is not None else None, int(sam), C.byref(out),
                                     C.byref(n)))
        s = C.string_at(out, n.value)
        lib().tg_free(out)
        return s


    def format_result_bam_raw(self, res: _Result, bases, offs, names, name_offs, quals, qual_offs, eof: bool = False) -> bytes:
        """BGZF-compressed BAM records of one batch (src/aligner.rs:69-72, 98-101)."""
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().tg_format_batch_bam(self.index._h, C.byref(res), _p(bases), _p(offs), _p(names), _p(name_offs),
                                         _p(quals), _p(qual_offs), int(eof), C.byref(out), C.byref(n)))
        s = C.string_at(out, n.value)
        lib().tg_free(out)
        return s


def compact_arrays(res: _ResultC):
    """Zero-copy numpy views of a compact result (valid until the next call on its context / tg_multi)."""
    n = res.n_reads
    first = np.ctypeslib.as_array(C.cast(res.read_aln_first, C.POINTER(C.c_uint32)), shape=(n,)) if n else np.zeros(0, np.uint32)
    count = np.ctypeslib.as_array(C.cast(res.read_aln_count, C.POINTER(C.c_uint32)), shape=(n,)) if n else np.zeros(0, np.uint32)
    na, no = int(res.alns_extent), int(res.ops_extent)
    alns = (np.ctypeslib.as_array(C.cast(res.alns, C.POINTER(C.c_uint8)), shape=(na * ALN_C_DTYPE.itemsize,)).view(ALN_C_DTYPE)
            if na else np.zeros(0, ALN_C_DTYPE))
    ops = np.ctypeslib.as_array(C.cast(res.ops, C.POINTER(C.c_uint32)), shape=(no,)) if no else np.zeros(0, np.uint32)
    return first, count, alns, ops


def expand_result(index: "Index", res: _ResultC, offs, ref_names=None, n_threads: int = 0) -> "AlignResult":
    """tg_result_expand: compact records -> AlignResult with wide records (host copies)."""
    offs = np.ascontiguousarray(offs, np.uint64)
    n = res.n_reads
    alns = np.zeros(int(res.alns_extent), ALN_DTYPE)
    first = np.zeros(n, np.uint64)
    if n:
        _check(lib().tg_result_expand(index._h, C.byref(res), _p(offs), _p(alns), _p(first), n_threads))
    _, count, _, ops = compact_arrays(res)
    counters = dict(swg_cells=res.swg_cells, swg_extensions=res.swg_extensions, seed_hits=res.seed_hits,
                    n_smems=res.n_smems, n_alns=res.n_alns, n_ops=res.n_ops, n_segments=res.n_segments)
    return AlignResult(first, count.copy(), alns, ops.copy(), counters, ref_names)


class MultiAligner:
    """Several GPUs of one host behind one call (tg_multi_*): the index is replicated with one NCCL broadcast, a batch is
    cut into contiguous shards, one per GPU, and comes back as ONE result in read order -- the reference's serial output
    order (src/aligner.rs:54-115).  `devices` may list a GPU more than once (two contexts on one GPU)."""

    def __init__(self, index: Index, opts: AlignOpts = None, devices=(0,)):
        self.index = index
        self.opts = opts or AlignOpts()
        self.devices = list(devices)
        h = C.c_void_p()
        o = self.opts._c()
        dv = (C.c_int * len(self.devices))(*self.devices)
        _check(lib().tg_multi_create(index._h, dv, len(self.devices), C.byref(o), C.byref(h)))
        self._h = h
        self._ref_names = [r.name for r in index.refs()]

    def close(self):
        if self._h:
            lib().tg_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def replication(self):
        """("nccl" | "peer-copy" | "single", milliseconds)"""
        ms = C.c_float()
        how = lib().tg_multi_replication(self._h, C.byref(ms)).decode()
        return how, ms.value

    def context(self, g: int) -> "Aligner":
        """The context of device slot g as an Aligner (owned by this object; not to be used while align_reads runs)."""
        return Aligner._view(lib().tg_multi_ctx(self._h, g), self.index, self.opts, self.devices[g], self._ref_names)

    def set_exact_cell_count(self, on: bool):
        for g in range(len(self.devices)):
            self.context(g).set_exact_cell_count(on)

    def align_reads_raw(self, bases_ptr: int, offs_ptr: int, n_reads: int) -> _ResultC:
        res = _ResultC()
        _check(lib().tg_multi_align_batch(self._h, C.c_void_p(bases_ptr), C.c_void_p(offs_ptr), n_reads, C.byref(res)))
        return res

    def align_reads(self, bases, offs) -> "AlignResult":
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        res = self.align_reads_raw(bases.ctypes.data, offs.ctypes.data, len(offs) - 1)
        return expand_result(self.index, res, offs, self._ref_names)

    def last_timing(self):
        """Per device slot: dict(wall_ms, seed_ms, extend_ms, dp_ms) of the last call."""
        out = []
        for g in range(len(self.devices)):
            w, a, b, d = C.c_double(), C.c_float(), C.c_float(), C.c_float()
            _check(lib().tg_multi_last_timing(self._h, g, C.byref(w), C.byref(a), C.byref(b), C.byref(d)))
            out.append(dict(wall_ms=w.value, seed_ms=a.value, extend_ms=b.value, dp_ms=d.value))
        return out


class OutputFormat:
    """src/aln_writer.rs:16-21"""
    Bam = "bam"
    Sam = "sam"
    Paf = "paf"


def parse_fastq(text: bytes):
    """needletail::parse_fastx_file for FASTQ text -> (bases, offs, names, name_offs, quals, qual_offs)."""
    L = lib()
    n = C.c_uint32()
    ptrs = [C.c_void_p() for _ in range(6)]
    _check(L.tg_parse_fastq(text, C.c_size_t(len(text)), C.byref(n), *[C.byref(p) for p in ptrs]))
    n = n.value

    def take(ptr, ctype, count, dtype):
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(count, 1),))[:count].astype(dtype, copy=True)
        return a
    offs = take(ptrs[1], C.c_uint64, n + 1, np.uint64)
    name_offs = take(ptrs[3], C.c_uint64, n + 1, np.uint64)
    qual_offs = take(ptrs[5], C.c_uint64, n + 1, np.uint64)
    bases = take(ptrs[0], C.c_uint8, int(offs[n]), np.uint8)
    names = take(ptrs[2], C.c_uint8, int(name_offs[n]), np.uint8)
    quals = take(ptrs[4], C.c_uint8, int(qual_offs[n]), np.uint8)
    for p in ptrs:
        L.tg_free(p)
    return bases, offs, names, name_offs, quals, qual_offs


def sam_header(index: Index) -> bytes:
    out, n = C.c_void_p(), C.c_size_t()
    _check(lib().tg_format_sam_header(index._h, C.byref(out), C.byref(n)))
    s = C.string_at(out, n.value)
    lib().tg_free(out)
    return s


def bam_header(index: Index) -> bytes:
    """BGZF block(s) with the BAM magic, header text and reference list (src/aligner.rs:41-47)."""
    out, n = C.c_void_p(), C.c_size_t()
    _check(lib().tg_format_bam_header(index._h, C.byref(out), C.byref(n)))
    s = C.string_at(out, n.value)
    lib().tg_free(out)
    return s


class _ReadAlns(C.Structure):
    _fields_ = [("n_alns", C.c_uint32), ("n_ops", C.c_uint32), ("alns", C.c_void_p), ("ops", C.c_void_p)]


class ThermiteAligner:
    """The reference's embedding wrapper (src/wrapper.rs:20-27): `align_read` for ONE read, callable from any number of
    host threads at once (the reference hands each worker thread its own clone).  All callers share one GPU context:
    reads are micro-batched by the library (`tg_batcher_*`, csrc/host_batcher.cpp).  `submit` / `wait` let one thread keep
    many reads in flight.  Records are those of `Aligner.align_reads`.  Defaults are the wrapper's (src/wrapper.rs:40-46)."""

    def __init__(self, index: Index, opts: "AlignOpts" = None, device: int = 0, max_batch_reads: int = 1 << 16,
                 max_wait_us: int = 200):
        self._aligner = Aligner(index, opts, device)
        h = C.c_void_p()
        _check(lib().tg_batcher_create(self._aligner._h, C.c_uint32(max_batch_reads), C.c_uint32(max_wait_us), C.byref(h)))
        self._h = h

    def opts(self) -> "AlignOpts":
        return self._aligner.opts

    def _take(self, ra: _ReadAlns) -> AlignResult:
        alns = np.zeros(ra.n_alns, ALN_DTYPE)
        ops = np.zeros(ra.n_ops, np.uint32)
        if ra.n_alns:
            C.memmove(alns.ctypes.data, ra.alns, ra.n_alns * ALN_DTYPE.itemsize)
        if ra.n_ops:
            C.memmove(ops.ctypes.data, ra.ops, ra.n_ops * 4)
        lib().tg_read_alns_free(C.byref(ra))
        return AlignResult(np.zeros(1, np.uint64), np.array([len(alns)], np.uint32), alns, ops, {}, self._aligner._ref_names)

    def align_read_raw(self, read: bytes) -> AlignResult:
        ra = _ReadAlns()
        _check(lib().tg_batcher_align_read(self._h, read, C.c_uint32(len(read)), C.byref(ra)))
        return self._take(ra)

    def align_read(self, read: bytes) -> List[GenomeAlignment]:
        """src/wrapper.rs:72 / src/aligner.rs:123 (blocking; thread-safe; ctypes releases the GIL while it waits)."""
        return self.align_read_raw(read).read_alignments(0)

    def submit(self, read: bytes) -> int:
        t = C.c_uint64()
        _check(lib().tg_batcher_submit(self._h, read, C.c_uint32(len(read)), C.byref(t)))
        return t.value

    def wait_raw(self, ticket: int) -> AlignResult:
        ra = _ReadAlns()
        _check(lib().tg_batcher_wait(self._h, C.c_uint64(ticket), C.byref(ra)))
        return self._take(ra)

    def wait(self, ticket: int) -> List[GenomeAlignment]:
        return self.wait_raw(ticket).read_alignments(0)

    def stats(self):
        r, b, l = C.c_uint64(), C.c_uint64(), C.c_uint32()
        _check(lib().tg_batcher_stats(self._h, C.byref(r), C.byref(b), C.byref(l)))
        return dict(reads=r.value, batches=b.value, largest_batch=l.value)

    def close(self):
        if self._h:
            lib().tg_batcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("pad", C.c_uint32), ("bases", C.c_void_p), ("offs", C.c_void_p), ("names", C.c_void_p),
                ("name_offs", C.c_void_p), ("quals", C.c_void_p), ("qual_offs", C.c_void_p)]


class FileStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_alns", C.c_uint64), ("n_batches", C.c_uint64), ("bytes_out", C.c_uint64),
                ("read_ms", C.c_double), ("align_ms", C.c_double), ("write_ms", C.c_double), ("format_ms", C.c_double),
                ("wall_ms", C.c_double), ("warm_reads", C.c_uint64), ("warm_ms", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        if self.n_batches > 2 and self.wall_ms > self.warm_ms:
            d["steady_reads_per_s"] = (self.n_reads - self.warm_reads) / ((self.wall_ms - self.warm_ms) / 1e3)
        return d


class FastqReader:
    """needletail::parse_fastx_file for FASTQ (src/aligner.rs:51-55) as a batch reader: plain, gzip or BGZF input, batches
    of up to `max_reads` reads parsed on all host cores into buffers owned by the reader (`tg_fastq_*`, csrc/host_stream.cpp).
    Iterating yields (bases, offs, names, name_offs, quals, qual_offs) as numpy COPIES."""
    FORMATS = ("plain", "gzip", "bgzf")

    def __init__(self, path: str, max_reads: int = 1 << 20):
        h = C.c_void_p()
        _check(lib().tg_fastq_open(path.encode(), C.byref(h)))
        self._h = h
        self.max_reads = max_reads

    @property
    def format(self) -> str:
        return self.FORMATS[lib().tg_fastq_format(self._h)]

    def next_raw(self) -> _ReadBatch:
        b = _ReadBatch()
        _check(lib().tg_fastq_next(self._h, C.c_uint32(self.max_reads), C.byref(b)))
        return b

    def __iter__(self):
        while True:
            b = self.next_raw()
            n = b.n_reads
            if n == 0:
                return

            def take(ptr, ctype, count, dtype):
                if count == 0:
                    return np.zeros(0, dtype)
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)
            offs = take(b.offs, C.c_uint64, n + 1, np.uint64)
            name_offs = take(b.name_offs, C.c_uint64, n + 1, np.uint64)
            qual_offs = take(b.qual_offs, C.c_uint64, n + 1, np.uint64)
            yield (take(b.bases, C.c_uint8, int(offs[n]), np.uint8), offs, take(b.names, C.c_uint8, int(name_offs[n]), np.uint8),
                   name_offs, take(b.quals, C.c_uint8, int(qual_offs[n]), np.uint8), qual_offs)

    def close(self):
        if self._h:
            lib().tg_fastq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_reads_from_file(index: Index, query_paths, output_path: str, output_fmt: str, align_opts: AlignOpts,
                          device: int = 0, batch_reads: int = 1 << 20, devices=None) -> dict:
    """src/aligner.rs:22-120: align FASTQ files (plain, .gz or BGZF) and write PAF, SAM or BAM, records in input order.
    Runs natively (`tg_align_files`, csrc/host_stream.cpp): reader, aligner and writers are three overlapped stages and no
    stage holds a whole file.  `devices` = several GPUs of this process (reads of every batch sharded, `MultiAligner`);
    else one GPU (`device`).  Returns the pipeline's statistics (reads, records, bytes, busy time per stage, wall time)."""
    fmt = {OutputFormat.Paf: 0, OutputFormat.Sam: 1, OutputFormat.Bam: 2}[output_fmt]
    if isinstance(query_paths, (str, bytes)):
        query_paths = [query_paths]
    paths = (C.c_char_p * len(query_paths))(*[q.encode() if isinstance(q, str) else q for q in query_paths])
    stats = FileStats()
    if devices is not None and len(devices) > 1:
        aligner = MultiAligner(index, align_opts, devices)
        ctx, multi = None, aligner._h
    else:
        aligner = Aligner(index, align_opts, device if devices is None else devices[0])
        ctx, multi = aligner._h, None
    _check(lib().tg_align_files(index._h, ctx, multi, paths, len(query_paths), output_path.encode(), fmt, C.c_uint32(batch_reads),
                                C.byref(stats)))
    return stats.as_dict()


# ---- multi-GPU plumbing (reads shard; index replicated; ordered merge on the host) ------------------------------
def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a batch owned by `rank` ([g*N/G, (g+1)*N/G))."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def merge_shards(shards):
    """Concatenate per-rank results (each (first, count, alns, ops), given in rank order) into one result whose
    reads are in the original order -- the reference's serial output order (src/aligner.rs:54-115)."""
    firsts, counts, alns, ops = [], [], [], []
    a_base = o_base = 0
    for first, count, a, o in shards:
        a = a.copy()
        a["ops_off"] += np.uint32(o_base)
        ex = a["aln_type"] == 0
        a["tx_ops_off"][ex] += np.uint32(o_base)
        firsts.append(first.astype(np.uint64) + np.uint64(a_base))
        counts.append(count)
        alns.append(a)
        ops.append(o)
        a_base += len(a)
        o_base += len(o)
    return (np.concatenate(firsts) if firsts else np.zeros(0, np.uint64),
            np.concatenate(counts) if counts else np.zeros(0, np.uint32),
            np.concatenate(alns) if alns else np.zeros(0, ALN_DTYPE),
            np.concatenate(ops) if ops else np.zeros(0, np.uint32))


def broadcast_index(index, rank: int, device=None):
    """Replicate the flat index with ONE torch.distributed broadcast (NCCL over NVLink on GPUs, gloo on CPU).
    Returns (Index, device blob tensor or None).  On GPU ranks the received blob is adopted in place as the HBM
    replica; no other collective is ever issued on the data path."""
    import torch
    import torch.distributed as dist
    on_gpu = device is not None and str(device).startswith("cuda")
    dev = torch.device(device) if on_gpu else torch.device("cpu")
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        host = torch.from_numpy(index.blob())
        n[0] = host.numel()
    dist.broadcast(n, 0)
    if rank == 0:
        buf = host.to(dev) if on_gpu else host
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, 0)
    if rank != 0:
        index = Index.from_blob(buf.cpu().numpy())
    if on_gpu:
        index.adopt_device_blob(buf.data_ptr(), buf.numel(), dev.index or 0, keepalive=buf)
        return index, buf
    return index, None
