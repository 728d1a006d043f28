"""ctypes binding of thermite_b200/csrc/libtg_hosttest.so -- the TEST-ONLY host build of the kernel logic
(tg_core.h under an emulated warp).  Lets the CPU test-suite exercise the exact code the GPU runs."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.orc import ALN_DTYPE

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "thermite_b200", "csrc")
_LIB = None

SEED_DTYPE = np.dtype([("query_idx", "<u4"), ("len", "<u4"), ("sa_lo", "<u4"), ("count", "<u4"),
                       ("direct", "<u4"), ("pad", "<u4")])


class Opts(C.Structure):
    _fields_ = [("min_seed_len", C.c_uint32), ("min_aln_score_percent", C.c_float), ("min_aln_score", C.c_int32),
                ("multimap_score_range", C.c_uint32), ("intron_mode", C.c_uint32)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_CSRC, "libtg_hosttest.so")
        srcs = [os.path.join(_CSRC, f) for f in ("hosttest.cpp", "host_index.cpp", "host_batcher.cpp", "tg_core.h", "tg_rounds.h", "tg_dpt.h", "tg_internal.h", "tg_textfmt.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", _CSRC, "libtg_hosttest.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.tg_last_error.restype = C.c_char_p
        L.ht_ctx_create.restype = C.c_void_p
        L.ht_align_batch.restype = C.c_void_p
        L.ht_align_batch_mode.restype = C.c_void_p
        L.ht_seed_batch.restype = C.c_longlong
        L.ht_swg_extend_batch.restype = C.c_longlong
        L.tg_index_host_sa.restype = C.c_void_p
        L.tg_index_host_text_len.restype = C.c_uint64
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def batcher_selftest(n_threads, per_thread, max_batch, max_wait_us, fail_every=0):
    """csrc/host_batcher.cpp over a stand-in batch aligner: (mismatches, batches, largest batch, reads of failed batches)."""
    nb, lg, nf = C.c_uint64(), C.c_uint32(), C.c_uint64()
    bad = lib().ht_batcher_selftest(C.c_int(n_threads), C.c_int(per_thread), C.c_uint32(max_batch), C.c_uint32(max_wait_us),
                                    C.c_int(fail_every), C.byref(nb), C.byref(lg), C.byref(nf))
    return bad, nb.value, lg.value, nf.value


def format_sam(index, res, bases, offs, names, name_offs, quals, qual_offs):
    """tg_textfmt.h (the routine tg_paf.cu's SAM kernels run, one thread per read) on the host: count, scan, write."""
    out, ln = C.c_void_p(), C.c_size_t()
    st = lib().ht_format_sam(index.h, C.byref(res), _p(bases), _p(offs), _p(names), _p(name_offs), _p(quals), _p(qual_offs),
                             C.byref(out), C.byref(ln))
    assert st == 0
    text = C.string_at(out, ln.value)
    C.CDLL(None).free(out)
    return text


class HostIndex:
    def __init__(self, fasta: bytes, gtf: bytes):
        h = C.c_void_p()
        st = lib().tg_index_host_create_from_memory(fasta, C.c_size_t(len(fasta)), gtf, C.c_size_t(len(gtf)), C.byref(h))
        if st != 0:
            raise RuntimeError(lib().tg_last_error().decode())
        self.h = h

    def sa(self):
        n = lib().tg_index_host_text_len(self.h)
        p = lib().tg_index_host_sa(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy()

    def __del__(self):
        try:
            lib().tg_index_host_destroy(self.h)
        except Exception:
            pass


class HostCtx:
    def __init__(self, index: HostIndex, k=20, pct=0.66, min_score=30, score_range=1, intron_mode=False):
        self.index = index
        self.opts = Opts(k, pct, min_score, score_range, int(intron_mode))
        self.h = C.c_void_p(lib().ht_ctx_create(index.h, C.byref(self.opts)))

    def __del__(self):
        try:
            lib().ht_ctx_destroy(self.h)
        except Exception:
            pass

    def seed_batch(self, bases, offs, lanes=1):
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        cap = int(len(bases)) + 16
        pool = np.zeros(cap, SEED_DTYPE)
        first = np.zeros(n, np.uint64)
        count = np.zeros(n, np.uint32)
        tot = lib().ht_seed_batch(self.h, _p(bases), _p(offs), n, lanes, _p(pool), C.c_uint64(cap), _p(first), _p(count))
        assert tot >= 0
        return pool[:tot], first, count

    def align_batch(self, bases, offs, lanes=1, bound_stop=False, rounds=False, compact=None):
        """compact=(first_base, ops_base): write tg_aln_c records rebased like one shard of tg_multi_align_batch and expand
        them back with tg_aln_expand."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64)
        n = len(offs) - 1
        if compact is not None:
            lib().ht_set_compact(1, C.c_uint64(compact[0]), C.c_uint64(compact[1]))
        r = lib().ht_align_batch_mode(self.h, _p(bases), _p(offs), n, lanes, int(bound_stop), int(rounds))
        assert r
        r = C.c_void_p(r)
        info = (C.c_uint64 * 8)()
        lib().ht_result_info(r, info)
        first = np.zeros(n, np.uint64)
        count = np.zeros(n, np.uint32)
        alns = np.zeros(info[0], ALN_DTYPE)
        ops = np.zeros(info[1], np.uint32)
        lib().ht_result_copy(r, _p(first), _p(count), _p(alns), _p(ops))
        lib().ht_result_free(r)
        return dict(first=first, count=count, alns=alns, ops=ops, cells=info[2], n_ext=info[3], hits=info[4],
                    flags=info[5], items=info[6], rounds=info[7])


def expand_seeds(pool, first, count, sa, r):
    """tg_seed records of read r -> list of (ref_idx, query_idx, len) in Index::all_smems order."""
    out = []
    for s in pool[int(first[r]): int(first[r]) + int(count[r])]:
        if s["direct"]:
            out.append((int(s["sa_lo"]), int(s["query_idx"]), int(s["len"])))
        else:
            for rk in range(int(s["count"]) - 1, -1, -1):
                out.append((int(sa[int(s["sa_lo"]) + rk]), int(s["query_idx"]), int(s["len"])))
    return out


def swg_extend_batch(xs, xoff, ys, yoff, bw, x_drop, lanes=1, bound_stop=False):
    n = len(bw)
    xs = np.ascontiguousarray(xs, np.uint8); ys = np.ascontiguousarray(ys, np.uint8)
    xoff = np.ascontiguousarray(xoff, np.uint64); yoff = np.ascontiguousarray(yoff, np.uint64)
    bw = np.ascontiguousarray(bw, np.uint32); x_drop = np.ascontiguousarray(x_drop, np.int32)
    score = np.zeros(n, np.int32); xend = np.zeros(n, np.uint32); yend = np.zeros(n, np.uint32)
    ops_off = np.zeros(n + 1, np.uint64)
    cap = int(len(xs) + len(ys) + 4 * n + 16)
    ops = np.zeros(cap, np.uint32)
    cells = C.c_uint64()
    tot = lib().ht_swg_extend_batch(_p(xs), _p(xoff), _p(ys), _p(yoff), n, _p(bw), _p(x_drop), lanes, int(bound_stop), _p(score),
                                    _p(xend), _p(yend), _p(ops_off), _p(ops), C.c_uint64(cap), C.byref(cells))
    assert tot >= 0
    return dict(score=score, xend=xend, yend=yend, ops_off=ops_off, ops=ops[:tot].copy(), cells=cells.value)


def compare_alignments(res_a, res_b, n_reads):
    """Compare two align results (dicts or objects with first/count or read_off, alns, ops).  Returns list of diffs."""
    def norm(res):
        if isinstance(res, dict):
            return res["first"], res["count"], res["alns"], res["ops"]
        ro = res.read_off
        return ro[:-1], (ro[1:] - ro[:-1]).astype(np.uint32), res.alns, res.ops
    fa, ca, aa, oa = norm(res_a)
    fb, cb, ab, ob = norm(res_b)
    diffs = []
    fields = [f for f in ALN_DTYPE.names if f not in ("ops_off", "tx_ops_off", "pad")]
    for r in range(n_reads):
        if int(ca[r]) != int(cb[r]):
            diffs.append((r, "count", int(ca[r]), int(cb[r])))
            continue
        for i in range(int(ca[r])):
            x, y = aa[int(fa[r]) + i], ab[int(fb[r]) + i]
            for f in fields:
                if x[f] != y[f]:
                    diffs.append((r, i, f, int(x[f]), int(y[f])))
            ox = oa[int(x["ops_off"]): int(x["ops_off"]) + int(x["ops_len"])]
            oy = ob[int(y["ops_off"]): int(y["ops_off"]) + int(y["ops_len"])]
            if len(ox) != len(oy) or (ox != oy).any():
                diffs.append((r, i, "ops", ox.tolist(), oy.tolist()))
            tx = oa[int(x["tx_ops_off"]): int(x["tx_ops_off"]) + int(x["tx_ops_len"])]
            ty = ob[int(y["tx_ops_off"]): int(y["tx_ops_off"]) + int(y["tx_ops_len"])]
            if len(tx) != len(ty) or (tx != ty).any():
                diffs.append((r, i, "tx_ops", tx.tolist(), ty.tolist()))
    return diffs
