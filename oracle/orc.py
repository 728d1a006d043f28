"""ctypes binding of the CPU ORACLE (liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The oracle restates 10XGenomics/thermite's CPU path (src/aligner.rs, src/swg.rs, src/txome.rs,
src/index.rs + recalled rust-bio behaviour); see oracle/thermite_oracle.hpp for the parity status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# identical field order/layout to `tg_aln` (include/thermite_gpu.h) -- asserted in tests
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

OP_NAMES = ["Match", "Subst", "Del", "Ins", "Xclip", "Yclip"]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("thermite_oracle.cpp", "oracle_capi.cpp", "thermite_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_last_error.restype = C.c_char_p
        L.orc_sizeof_aln.restype = C.c_size_t
        for f in ("orc_index_create_files", "orc_index_create_mem", "orc_align_batch"):
            getattr(L, f).restype = C.c_void_p
        L.orc_index_text_len.restype = C.c_uint64
        L.orc_index_ref.restype = C.c_char_p
        L.orc_index_tx.restype = C.c_char_p
        L.orc_index_gene_id.restype = C.c_char_p
        L.orc_index_gene_name.restype = C.c_char_p
        L.orc_interval_find.restype = C.c_uint64
        L.orc_all_smems.restype = C.c_int64
        L.orc_swg_extend_batch.restype = C.c_int64
        L.orc_swg_extend_batch_mt.restype = C.c_int64
        L.orc_filter_overlapping.restype = C.c_int64
        L.orc_result_n_alns.restype = C.c_uint64
        L.orc_result_n_ops.restype = C.c_uint64
        L.orc_result_seconds.restype = C.c_double
        L.orc_align_fastq_text.restype = C.c_void_p
        assert L.orc_sizeof_aln() == ALN_DTYPE.itemsize, (L.orc_sizeof_aln(), ALN_DTYPE.itemsize)
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def err():
    return lib().orc_last_error().decode()


def expand_ops(words):
    """RLE words -> list of (kind_name, n) one entry per reference AlignmentOperation."""
    out = []
    for w in np.asarray(words, dtype=np.uint32).tolist():
        k, n = w & 7, w >> 3
        if k <= 3:
            out.extend([(OP_NAMES[k], 1)] * n)
        else:
            out.append((OP_NAMES[k], n))
    return out


def swg_extend(x: bytes, y: bytes, bw: int, x_drop: int, max_bw=None):
    """-> (score, xend, yend, rle_ops ndarray, cells) or raises RuntimeError when the reference panics."""
    L = lib()
    cap = 2 * (len(x) + len(y)) + 8
    ops = np.zeros(cap, np.uint32)
    score, xend, yend, n_ops = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    cells = C.c_uint64()
    st = L.orc_swg_extend(x, C.c_uint64(len(x)), y, C.c_uint64(len(y)), C.c_uint64(bw if max_bw is None else max_bw),
                          C.c_uint64(bw), C.c_int32(x_drop), C.byref(score), C.byref(xend), C.byref(yend), _p(ops),
                          C.c_uint32(cap), C.byref(n_ops), C.byref(cells))
    if st != 0:
        raise RuntimeError("reference panics: " + err())
    return score.value, xend.value, yend.value, ops[: n_ops.value].copy(), cells.value


def swg_extend_batch(xs, xoff, ys, yoff, bw, x_drop, n_threads=1):
    """xs/ys: uint8 arrays of concatenated sequences; xoff/yoff: uint64[n+1]; bw uint32[n]; x_drop int32[n]."""
    L = lib()
    n = len(bw)
    xs = np.ascontiguousarray(xs, np.uint8); ys = np.ascontiguousarray(ys, np.uint8)
    xoff = np.ascontiguousarray(xoff, np.uint64); yoff = np.ascontiguousarray(yoff, np.uint64)
    bw = np.ascontiguousarray(bw, np.uint32); x_drop = np.ascontiguousarray(x_drop, np.int32)
    score = np.zeros(n, np.int32); xend = np.zeros(n, np.uint32); yend = np.zeros(n, np.uint32)
    ops_off = np.zeros(n + 1, np.uint64)
    cap = int(len(xs) + len(ys) + 4 * n + 16)
    ops = np.zeros(cap, np.uint32)
    cells = C.c_uint64()
    if n_threads > 1:
        tot = L.orc_swg_extend_batch_mt(_p(xs), _p(xoff), _p(ys), _p(yoff), C.c_uint64(n), _p(bw), _p(x_drop), _p(score),
                                        _p(xend), _p(yend), _p(ops_off), _p(ops), C.c_uint64(cap), C.byref(cells), C.c_int(n_threads))
    else:
        tot = L.orc_swg_extend_batch(_p(xs), _p(xoff), _p(ys), _p(yoff), C.c_uint64(n), _p(bw), _p(x_drop), _p(score),
                                     _p(xend), _p(yend), _p(ops_off), _p(ops), C.c_uint64(cap), C.byref(cells))
    if tot < 0:
        raise RuntimeError("reference panics: " + err())
    assert tot <= cap
    return dict(score=score, xend=xend, yend=yend, ops_off=ops_off, ops=ops[:tot].copy(), cells=cells.value)


class AlignResult:
    def __init__(self, alns, ops, read_off, seconds):
        self.alns, self.ops, self.read_off, self.seconds = alns, ops, read_off, seconds


class Index:
    """Oracle Index (src/index.rs:40-44)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index creation failed: " + err())
        self.h = C.c_void_p(handle)

    @classmethod
    def create_from_files(cls, ref_path, annot_path, sa_rate=32, occ_rate=128):
        return cls(lib().orc_index_create_files(ref_path.encode(), annot_path.encode(), sa_rate, occ_rate))

    @classmethod
    def create(cls, fasta_text: bytes, gtf_text: bytes, sa_rate=32, occ_rate=128):
        return cls(lib().orc_index_create_mem(fasta_text, C.c_size_t(len(fasta_text)), gtf_text,
                                              C.c_size_t(len(gtf_text)), sa_rate, occ_rate))

    def __del__(self):
        try:
            lib().orc_index_free(self.h)
        except Exception:
            pass

    def text(self):
        n = lib().orc_index_text_len(self.h)
        out = np.zeros(n, np.uint8)
        lib().orc_index_text(self.h, _p(out))
        return out

    def sa(self):
        n = lib().orc_index_text_len(self.h)
        out = np.zeros(n, np.uint32)
        lib().orc_index_sa(self.h, _p(out))
        return out

    def refs(self):
        out = []
        for i in range(lib().orc_index_n_refs(self.h)):
            v = (C.c_uint64 * 4)()
            name = lib().orc_index_ref(self.h, i, v).decode()
            out.append(dict(name=name, start_idx=v[0], end_idx=v[1], len=v[2], strand=bool(v[3])))
        return out

    def txs(self):
        out = []
        for i in range(lib().orc_index_n_txs(self.h)):
            v = (C.c_uint64 * 4)()
            tid = lib().orc_index_tx(self.h, i, v).decode()
            seq = np.zeros(v[3], np.uint8)
            lib().orc_index_tx_seq(self.h, i, _p(seq))
            ex = np.zeros(2 * v[2], np.uint64)
            lib().orc_index_tx_exons(self.h, i, _p(ex))
            out.append(dict(id=tid, gene_idx=int(v[0]), strand=bool(v[1]), seq=seq.tobytes(),
                            exons=ex.reshape(-1, 2).tolist()))
        return out

    def genes(self):
        return [dict(id=lib().orc_index_gene_id(self.h, i).decode(), name=lib().orc_index_gene_name(self.h, i).decode())
                for i in range(lib().orc_index_n_genes(self.h))]

    def interval_find(self, tree, s, e):
        cap = 1 << 16
        out = np.zeros(cap, np.uint64)
        n = lib().orc_interval_find(self.h, tree, C.c_uint64(s), C.c_uint64(e), _p(out), C.c_uint64(cap))
        return out[:n].tolist()

    def all_smems(self, read: bytes, k: int, brute=False):
        cap = 1 << 20
        out = np.zeros(3 * cap, np.uint64)
        n = lib().orc_all_smems(self.h, read, C.c_uint64(len(read)), k, int(brute), _p(out), C.c_uint64(cap))
        if n < 0:
            raise RuntimeError(err())
        return [tuple(int(v) for v in out[3 * i: 3 * i + 3]) for i in range(min(n, cap))]

    def counters(self):
        v = (C.c_uint64 * 6)()
        lib().orc_counters(self.h, v)
        return dict(swg_cells=v[0], swg_calls=v[1], occ_lookups=v[2], fmd_ext=v[3], sa_locates=v[4], hits=v[5])

    def counters_reset(self):
        lib().orc_counters_reset(self.h)

    def align_batch(self, bases, offs, k=20, pct=0.66, min_score=30, score_range=1, intron_mode=False, n_threads=1):
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        r = lib().orc_align_batch(self.h, _p(bases), _p(offs), C.c_uint64(n), k, C.c_float(pct), min_score,
                                  score_range, int(intron_mode), n_threads)
        if not r:
            raise RuntimeError("oracle align failed: " + err())
        r = C.c_void_p(r)
        alns = np.zeros(lib().orc_result_n_alns(r), ALN_DTYPE)
        ops = np.zeros(lib().orc_result_n_ops(r), np.uint32)
        read_off = np.zeros(n + 1, np.uint64)
        lib().orc_result_copy(r, _p(alns), _p(ops), _p(read_off))
        secs = lib().orc_result_seconds(r)
        lib().orc_result_free(r)
        return AlignResult(alns, ops, read_off, secs)

    def align_fastq_text(self, fastq: bytes, k=20, pct=0.66, min_score=30, score_range=1, intron_mode=False, sam=False):
        n = C.c_uint64()
        p = lib().orc_align_fastq_text(self.h, fastq, C.c_uint64(len(fastq)), k, C.c_float(pct), min_score,
                                       score_range, int(intron_mode), int(sam), C.byref(n))
        if not p:
            raise RuntimeError("oracle align failed: " + err())
        s = C.string_at(p, n.value)
        lib().orc_free(C.c_void_p(p))
        return s


def suffix_array(text: bytes):
    t = np.frombuffer(text, np.uint8)
    out = np.zeros(len(t), np.uint32)
    assert lib().orc_suffix_array(_p(t), C.c_uint64(len(t)), _p(out)) == 0
    return out


def lift_mem_to_tx(exons, mem):
    ex = np.asarray(exons, np.uint64).reshape(-1)
    mi = np.asarray(mem, np.uint64)
    mo = np.zeros(3, np.uint64)
    assert lib().orc_lift_mem_to_tx(_p(ex), len(ex) // 2, _p(mi), _p(mo)) == 0, err()
    return tuple(int(v) for v in mo)


def lift_tx_to_gx(exons, ystart, yend, ops):
    """ops: list of (kind_name, n).  -> (ystart, yend, ops)"""
    ex = np.asarray(exons, np.uint64).reshape(-1)
    w = np.asarray([OP_NAMES.index(k) | (n << 3) for k, n in ops], np.uint32)
    out = np.zeros(4 * len(w) + 8, np.uint32)
    yy = np.zeros(2, np.uint64)
    n = C.c_uint32()
    st = lib().orc_lift_tx_to_gx(_p(ex), len(ex) // 2, C.c_uint64(ystart), C.c_uint64(yend), _p(w), len(w), _p(yy),
                                 _p(out), len(out), C.byref(n))
    assert st == 0, err()
    return int(yy[0]), int(yy[1]), [(OP_NAMES[v & 7], v >> 3) for v in out[: n.value].tolist()]


def filter_overlapping(rows):
    """rows: list of (name_id, strand, ystart, yend, score) -> kept row indices in output order."""
    r = np.asarray(rows, np.int64).reshape(-1)
    kept = np.zeros(len(rows), np.uint32)
    n = lib().orc_filter_overlapping(_p(r), len(rows), _p(kept))
    assert n >= 0, err()
    return kept[:n].tolist()


def extend_left_right(ref_seq: bytes, hit, read: bytes, bw, x_drop, max_bw=None):
    out = np.zeros(7, np.int64)
    cap = 2 * (len(ref_seq) + len(read)) + 8
    ops = np.zeros(cap, np.uint32)
    n = C.c_uint32()
    st = lib().orc_extend_left_right(ref_seq, C.c_uint64(len(ref_seq)), C.c_uint64(hit[0]), C.c_uint64(hit[1]),
                                     C.c_uint64(hit[2]), read, C.c_uint64(len(read)),
                                     C.c_uint64(bw if max_bw is None else max_bw), C.c_uint64(bw), C.c_int32(x_drop),
                                     _p(out), _p(ops), cap, C.byref(n))
    assert st == 0, err()
    keys = ["score", "ystart", "xstart", "yend", "xend", "ylen", "xlen"]
    d = {k: int(v) for k, v in zip(keys, out)}
    d["ops"] = expand_ops(ops[: n.value])
    return d


def multimapq(n):
    return lib().orc_multimapq(C.c_uint64(n))
