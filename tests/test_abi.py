"""The C-ABI library loads without a GPU, exports every symbol include/thermite_gpu.h declares, its host-side
entry points work, and device entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import thermite_b200 as tb
from common import bam_to_sam, golden
from oracle import orc
from thermite_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "thermite_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # drop comments
    declared = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", hdr))
    L = tb.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), sym
    assert declared == set(api.ABI_SYMBOLS), declared ^ set(api.ABI_SYMBOLS)


def test_record_layouts_match_between_abi_and_oracle():
    assert api.ALN_DTYPE == orc.ALN_DTYPE and api.ALN_DTYPE.itemsize == 104
    assert api.SEED_DTYPE.itemsize == 24


def test_host_index_metadata_and_blob_roundtrip(tmp_path):
    fa, gtf = golden("test_ref.fasta"), golden("test_ref.gtf")
    ix = tb.Index.create_from_memory(fa, gtf)
    oix = orc.Index.create(fa, gtf)
    assert [(r.name, r.strand, r.len, r.start_idx, r.end_idx) for r in ix.refs()] == \
           [(r["name"], r["strand"], r["len"], r["start_idx"], r["end_idx"]) for r in oix.refs()]
    assert [(t.id, t.strand, t.gene_idx, t.n_exons, t.seq_len) for t in ix.txome().txs] == \
           [(t["id"], t["strand"], t["gene_idx"], len(t["exons"]), len(t["seq"])) for t in oix.txs()]
    assert [(g.id, g.name) for g in ix.txome().genes] == [(g["id"], g["name"]) for g in oix.genes()]
    assert (ix.suffix_array() == oix.sa()).all()
    p = str(tmp_path / "t.tai")
    ix.save(p)
    ix2 = tb.Index.load(p)
    assert (ix2.blob() == ix.blob()).all() and ix2.refs() == ix.refs()
    ix3 = tb.Index.from_blob(ix.blob().copy())
    assert ix3.txome().txs == ix.txome().txs
    assert tb.sam_header(ix) == b"@SQ\tSN:some_ref\tLN:12\n@SQ\tSN:another_seq\tLN:12\n@SQ\tSN:introns_seq\tLN:26\n" \
                                b"@SQ\tSN:introns_revcomp\tLN:26\n@PG\tID:thermite\n"


def test_bad_inputs_are_errors_not_crashes():
    with pytest.raises(tb.ThermiteError):
        tb.Index.create_from_files("/nonexistent.fa", "/nonexistent.gtf")
    with pytest.raises(tb.ThermiteError):
        tb.Index.create_from_memory(b">a\nACGT1234\n", b"")
    # IUPAC ambiguity codes are indexed as N (documented deviation: the reference's FM alphabet has no rank for them)
    assert (tb.Index.create_from_memory(b">a\nACGTRYKMacgtn\n", b"").blob() ==
            tb.Index.create_from_memory(b">a\nACGTNNNNACGTN\n", b"").blob()).all()
    with pytest.raises(tb.ThermiteError):
        tb.Index.from_blob(np.zeros(1000, np.uint8))
    with pytest.raises(tb.ThermiteError):
        tb.Index.create_from_memory(b">a\nACGTACGT\n", b'b\t.\texon\t1\t4\t.\t+\t.\tgene_id "g"; transcript_id "t";\n')


def test_text_beyond_the_32_bit_limit_is_a_capacity_error():
    """The reference indexes with usize (src/index.rs:383-388); this build keeps 32-bit text positions and must say so:
    a reference whose both-strand text reaches 2^31 symbols is refused with TG_ERR_CAPACITY before anything is built."""
    n = (1 << 30) + 4096                       # 2 * (n + 1) symbols >= 2^31
    fa = b">big\n" + b"A" * n + b"\n"
    with pytest.raises(tb.ThermiteError) as e:
        tb.Index.create_from_memory(fa, b"")
    assert "2^31" in str(e.value)


def test_corrupt_index_blobs_are_refused(tmp_path):
    """A truncated, stale or foreign .tai must fail cleanly in tg_index_host_load / from_blob (the reference's bincode load
    does), never reach the string decoder or a kernel with out-of-range offsets."""
    ix = tb.Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"))
    good = ix.blob().copy()
    hdr = good[:512].view(np.uint64)
    assert tb.Index.from_blob(good).refs() == ix.refs()
    # field offsets in TgBlobHeader (u64 words): 0 magic, 1 nbytes, 2 text_len, 3 n_refs, 4 n_txs, 5 n_genes, 6/7 tree nodes,
    # 8 n_tx_exons, 9 txseq_len, 10/11 roots, 12 device_bytes, 13..30 section offsets, 33 format_version
    assert hdr[1] == len(good) and hdr[33] >= 4
    p = str(tmp_path / "t.tai")
    for what, mutate in (
            ("truncated", lambda b: b[: len(b) - 64]),
            ("truncated to the header", lambda b: b[:400]),
            ("other version", lambda b: _set(b, 33, 3)),
            ("text_len beyond the suffix array", lambda b: _set(b, 2, int(b[:512].view(np.uint64)[2]) + 4096)),
            ("more transcripts than there are", lambda b: _set(b, 4, 1 << 20)),
            ("section offset outside the blob", lambda b: _set(b, 14, len(b) * 2)),
            ("unaligned section offset", lambda b: _set(b, 15, int(b[:512].view(np.uint64)[15]) + 3)),
            ("device prefix larger than the blob", lambda b: _set(b, 12, len(b) + 256)),
            ("string table offsets decreasing", lambda b: _poke_u64(b, int(b[:512].view(np.uint64)[23]) + 8, 1 << 40)),
            ("suffix array entry outside the text", lambda b: _poke_u32(b, int(b[:512].view(np.uint64)[14]), 0xFFFFFFF0)),
            ("gene index out of range", lambda b: _poke_u32(b, int(b[:512].view(np.uint64)[27]), 9999))):
        bad = mutate(good.copy())
        with pytest.raises(tb.ThermiteError, match="index blob"):
            tb.Index.from_blob(bad)
        bad.tofile(p)
        with pytest.raises(tb.ThermiteError, match="index blob"):
            tb.Index.load(p)


def _set(b, word, value):
    b[:512].view(np.uint64)[word] = value
    return b


def _poke_u64(b, off, value):
    b[off: off + 8].view(np.uint64)[0] = value
    return b


def _poke_u32(b, off, value):
    b[off: off + 4].view(np.uint32)[0] = value
    return b


def test_fastq_parser_and_writers_match_oracle_text():
    """Host-side IO: records produced by the ORACLE, formatted by the PRODUCT writer, must equal the oracle's text."""
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    bases, offs, names, name_offs, quals, qual_offs = tb.parse_fastq(fq)
    assert len(offs) == 11 and bytes(names[: int(name_offs[1])]) == b"all_match"
    oix = orc.Index.create(fa, gtf)
    ores = oix.align_batch(bases, offs, k=3, min_score=0)
    ix = tb.Index.create_from_memory(fa, gtf)
    n = len(offs) - 1
    first = np.ascontiguousarray(ores.read_off[:-1])
    count = np.ascontiguousarray((ores.read_off[1:] - ores.read_off[:-1]).astype(np.uint32))
    res = api._Result(n, len(ores.alns), len(ores.ops), first.ctypes.data, count.ctypes.data, ores.alns.ctypes.data,
                      ores.ops.ctypes.data, 0, 0, 0, 0)
    for sam in (False, True):
        out, ln = C.c_void_p(), C.c_size_t()
        st = tb.lib().tg_format_batch(ix._h, C.byref(res), api._p(bases), api._p(offs), api._p(names), api._p(name_offs),
                                      api._p(quals), api._p(qual_offs), int(sam), C.byref(out), C.byref(ln))
        assert st == 0
        text = C.string_at(out, ln.value)
        tb.lib().tg_free(out)
        want = oix.align_fastq_text(fq, k=3, min_score=0, sam=sam)
        if sam:
            text = tb.sam_header(ix) + text
        assert text == want
    # the SAM routine the device formatter runs (csrc/tg_textfmt.h), built for the host
    import ht
    assert tb.sam_header(ix) + ht.format_sam(ht.HostIndex(fa, gtf), res, bases, offs, names, name_offs, quals, qual_offs) == want


def test_parallel_fastq_parser_and_writers_large_input():
    """The multi-threaded ingest / writers: a FASTQ text large enough to be cut into per-thread
    segments, with quality lines that start with '@' or '+', blank lines between records and CRLF line ends, must
    parse like a plain sequential parser; oracle records formatted by the (threaded) product writer must equal the
    oracle's own PAF / SAM text."""
    from common import small_world
    from thermite_b200 import synth
    contigs, gtf, txs, fa = small_world(5)
    n = 60_000
    rbases, roffs = synth.make_reads(3, contigs, txs, n, L=91, sub=0.01, ins=0.001, dele=0.001)
    rng = np.random.default_rng(9)
    qual_first = rng.choice(np.frombuffer(b"@+FI#", np.uint8), n)
    recs, want_names, want_quals = [], [], []
    for r in range(n):
        name = b"read%d extra field" % r
        seq = rbases[int(roffs[r]): int(roffs[r + 1])].tobytes()
        q = bytes([qual_first[r]]) + b"F" * 90
        eol = b"\r\n" if r % 1000 == 7 else b"\n"
        recs.append(b"@" + name + eol + seq + eol + b"+" + eol + q + eol + (b"\n" if r % 5000 == 11 else b""))
        want_names.append(name); want_quals.append(q)
    fq = b"".join(recs) + b"@truncated\nACGT\n"  # incomplete last record: dropped
    assert len(fq) > (8 << 20)
    bases, offs, names, name_offs, quals, qual_offs = tb.parse_fastq(fq)
    assert len(offs) - 1 == n and np.array_equal(bases, rbases) and np.array_equal(offs, roffs)
    assert names.tobytes() == b"".join(want_names) and quals.tobytes() == b"".join(want_quals)
    assert np.array_equal(name_offs, np.cumsum([0] + [len(x) for x in want_names]).astype(np.uint64))
    assert np.array_equal(qual_offs, (np.arange(n + 1) * 91).astype(np.uint64))
    # writers: oracle records -> product text (threaded: n >= 32768) == oracle text
    oix = orc.Index.create(fa, gtf)
    flags = dict(k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    ores = oix.align_batch(bases, offs, n_threads=8, **flags)
    ix = tb.Index.create_from_memory(fa, gtf)
    first = np.ascontiguousarray(ores.read_off[:-1])
    count = np.ascontiguousarray((ores.read_off[1:] - ores.read_off[:-1]).astype(np.uint32))
    res = api._Result(n, len(ores.alns), len(ores.ops), first.ctypes.data, count.ctypes.data, ores.alns.ctypes.data,
                      ores.ops.ctypes.data, 0, 0, 0, 0)
    plain = b"".join(b"@" + want_names[r] + b"\n" + rbases[int(roffs[r]): int(roffs[r + 1])].tobytes() + b"\n+\n" + want_quals[r] + b"\n"
                     for r in range(n))
    for sam in (False, True):
        out, ln = C.c_void_p(), C.c_size_t()
        assert tb.lib().tg_format_batch(ix._h, C.byref(res), api._p(bases), api._p(offs), api._p(names), api._p(name_offs),
                                        api._p(quals), api._p(qual_offs), int(sam), C.byref(out), C.byref(ln)) == 0
        text = C.string_at(out, ln.value)
        tb.lib().tg_free(out)
        want = oix.align_fastq_text(plain, sam=sam, **flags)
        if sam:
            text = tb.sam_header(ix) + text
        assert text == want
    import ht
    assert tb.sam_header(ix) + ht.format_sam(ht.HostIndex(fa, gtf), res, bases, offs, names, name_offs, quals, qual_offs) == want


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ix = tb.Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"))
    with pytest.raises(tb.ThermiteError, match="no CPU fallback|CUDA"):
        tb.Aligner(ix)
    # the GPU suffix-array builder has no host fallback either
    with pytest.raises(tb.ThermiteError, match="no CPU fallback"):
        tb.suffix_array_gpu(ix.text4(), ix.text_len(), 0)
    with pytest.raises(tb.ThermiteError, match="no CPU fallback"):
        tb.Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"), sa_device=0)
    with pytest.raises(tb.ThermiteError, match="no CPU fallback"):
        tb.MultiAligner(ix, devices=[0, 1])


def test_compact_record_layout_and_host_side_expansion():
    """tg_aln_c is 40 bytes and tg_aln_expand / tg_result_expand (host functions) rebuild the wide record bit for bit:
    oracle records -> compact (what the device writes) -> expanded == oracle records."""
    assert api.ALN_C_DTYPE.itemsize == 40
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    bases, offs, *_ = tb.parse_fastq(fq)
    ores = orc.Index.create(fa, gtf).align_batch(bases, offs, k=3, min_score=0, intron_mode=True)
    ix = tb.Index.create_from_memory(fa, gtf)
    n = len(offs) - 1
    w = ores.alns
    assert len(w) > 5 and (w["aln_type"] == 0).any() and (w["aln_type"] != 0).any()
    # the oracle lays transcript operations right behind the genome operations, like the device's output pools
    ex = w["aln_type"] == 0
    assert (w["tx_ops_off"][ex] == w["ops_off"][ex] + w["ops_len"][ex]).all()
    gap = 3  # a gap in front of the pool, as between two segments of a multi-GPU result
    c = np.zeros(len(w) + gap, api.ALN_C_DTYPE)
    for f in ("ystart", "yend", "tx_ystart", "tx_yend", "ref_id", "tx_or_gene_idx", "ops_off", "score", "xstart", "xend",
              "ops_len", "tx_ops_len", "aln_type", "primary"):
        c[f][gap:] = w[f]
    first = (ores.read_off[:-1] + gap).astype(np.uint32)
    count = (ores.read_off[1:] - ores.read_off[:-1]).astype(np.uint32)
    ops = np.ascontiguousarray(ores.ops)
    res = api._ResultC(n, 2, len(w), len(ops), len(c), len(ops), first.ctypes.data, count.ctypes.data, c.ctypes.data,
                       ops.ctypes.data, 0, 0, 0, 0)
    out = api.expand_result(ix, res, offs)
    assert (out.first == first).all() and (out.alns[:gap].view(np.uint8) == 0).all()
    got = out.alns[gap:]
    for f in api.ALN_DTYPE.names:
        if f == "tx_ops_off":
            assert (got[f][ex] == w[f][ex]).all() and (got[f][~ex] == 0).all()
        else:
            assert (got[f] == w[f]).all(), f
    one = np.zeros(1, api.ALN_DTYPE)
    assert tb.lib().tg_aln_expand(ix._h, c[gap:].ctypes.data_as(C.c_void_p), C.c_uint32(int(offs[1] - offs[0])),
                                  one.ctypes.data_as(C.c_void_p)) == 0
    assert one[0]["ylen"] == w[0]["ylen"] and one[0]["strand"] == w[0]["strand"] and one[0]["xlen"] == offs[1] - offs[0]
    c["ref_id"][gap] = 99  # a record of another index: refused, not read out of bounds
    with pytest.raises(tb.ThermiteError):
        api.expand_result(ix, res, offs)


def test_text4_is_the_packed_both_strand_text():
    """tg_index_host_text4 (input of tg_suffix_array_gpu): 4-bit codes $ACGNT = 0..5, first symbol in the top nibble,
    zero padded; the suffix array over it orders the suffixes byte-lexicographically, a proper prefix first."""
    ix = tb.Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"))
    n, t4 = ix.text_len(), ix.text4()
    codes = np.array([(int(t4[i >> 4]) >> (4 * (15 - (i & 15)))) & 15 for i in range(n)], np.uint8)
    assert codes.max() <= 5 and (codes == 0).sum() == len(ix.refs()) and codes[-1] == 0
    assert all(int(w) == 0 for w in t4[(n + 15) // 16:]) and len(t4) == n // 16 + 4
    b = bytes(codes + 1)
    assert list(ix.suffix_array()) == sorted(range(n), key=lambda i: b[i:])


def test_bam_output_decodes_to_the_sam_text():
    """SURVEY 8f N4 / OutputFormat::Bam (src/aligner.rs:41-47, 69-72, 98-101): header ++ records ++ EOF block is a valid
    BGZF / BAM stream whose header, reference list and records decode to exactly the SAM text of the same records
    (oracle records, product writer): mapped and unmapped reads, both strands, secondary records, every tag."""
    from common import small_world
    from thermite_b200 import synth
    cases = []
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    cases.append((fa, gtf, tb.parse_fastq(fq), dict(k=3, min_score=0)))
    contigs, gtf2, txs, fa2 = small_world(5)
    n = 3000
    rb, ro = synth.make_reads(3, contigs, txs, n, L=91, sub=0.01, ins=0.002, dele=0.002)
    rng = np.random.default_rng(1)
    fq2 = b"".join(b"@r%d some comment\n%s\n+\n%s\n" % (i, bytes(rb[int(ro[i]):int(ro[i + 1])]) if i % 50 else b"ACGTNNNNACGT" * 3,
                                                   bytes(rng.integers(35, 74, 36 if i % 50 == 0 else int(ro[i + 1] - ro[i]), dtype=np.uint8)))
                   for i in range(n))
    cases.append((fa2, gtf2, tb.parse_fastq(fq2), dict(k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)))
    for fa, gtf, (bases, offs, names, name_offs, quals, qual_offs), kw in cases:
        oix = orc.Index.create(fa, gtf)
        ores = oix.align_batch(bases, offs, **kw)
        ix = tb.Index.create_from_memory(fa, gtf)
        n = len(offs) - 1
        first = np.ascontiguousarray(ores.read_off[:-1])
        count = np.ascontiguousarray((ores.read_off[1:] - ores.read_off[:-1]).astype(np.uint32))
        res = api._Result(n, len(ores.alns), len(ores.ops), first.ctypes.data, count.ctypes.data, ores.alns.ctypes.data,
                          ores.ops.ctypes.data, 0, 0, 0, 0)
        out, ln = C.c_void_p(), C.c_size_t()
        args = (ix._h, C.byref(res), api._p(bases), api._p(offs), api._p(names), api._p(name_offs), api._p(quals), api._p(qual_offs))
        assert tb.lib().tg_format_batch(*args, 1, C.byref(out), C.byref(ln)) == 0
        sam = C.string_at(out, ln.value)
        tb.lib().tg_free(out)
        assert tb.lib().tg_format_batch_bam(*args, 1, C.byref(out), C.byref(ln)) == 0
        bam = tb.bam_header(ix) + C.string_at(out, ln.value)
        tb.lib().tg_free(out)
        assert bam.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        text, refs, lines, n_blocks = bam_to_sam(bam)
        assert text == tb.sam_header(ix)
        assert [b"@SQ\tSN:%s\tLN:%d" % r for r in refs] == [l for l in text.split(b"\n") if l.startswith(b"@SQ")]
        assert lines == sam and lines.count(b"\n") >= n
        assert n_blocks >= 3


def test_device_sam_routine_equals_host_writer_on_synthetic_records():
    """csrc/tg_textfmt.h (what tg_paf.cu's SAM kernels run) against the host writer (tg_format_batch) on records no aligner
    would produce together: empty reads and empty quality strings, lower-case and N bases on the reverse strand, negative and
    zero scores, secondary records, every alignment type, operation lists with adjacent clips of equal and unequal length,
    runs of Match / Subst to be merged, zero-length lists, up to 7 records per read, names with and without a space."""
    import ht
    from common import small_world
    from oracle.orc import ALN_DTYPE
    contigs, gtf, txs, fa = small_world(5)
    ix = tb.Index.create_from_memory(fa, gtf)
    L = tb.lib()
    n_refs, n_txs, n_genes = L.tg_index_host_n_refs(ix._h), L.tg_index_host_n_txs(ix._h), L.tg_index_host_n_genes(ix._h)
    assert n_refs >= 2 and n_txs >= 1 and n_genes >= 1
    rng = np.random.default_rng(77)
    n = 3000
    lens = rng.integers(0, 120, n)
    lens[rng.random(n) < 0.05] = 0
    qlens = np.where(rng.random(n) < 0.1, 0, lens)
    bases = rng.choice(np.frombuffer(b"ACGTNacgtn", np.uint8), int(lens.sum()))
    quals = rng.integers(33, 74, int(qlens.sum())).astype(np.uint8)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    qual_offs = np.concatenate([[0], np.cumsum(qlens)]).astype(np.uint64)
    name_list = [(b"r%d" % i) + (b" tail %d x" % i if i % 3 else b"") for i in range(n)]
    names = np.frombuffer(b"".join(name_list), np.uint8).copy()
    name_offs = np.concatenate([[0], np.cumsum([len(x) for x in name_list])]).astype(np.uint64)
    count = rng.integers(0, 8, n).astype(np.uint32)
    count[rng.random(n) < 0.2] = 0
    first = np.concatenate([[0], np.cumsum(count)]).astype(np.uint64)
    na = int(count.sum())
    alns = np.zeros(na, ALN_DTYPE)
    ops = []

    def op_list():
        k = int(rng.integers(0, 9))
        w = []
        for _ in range(k):
            kind = int(rng.integers(0, 6))
            run = int(rng.integers(1, 400))
            w.append(run << 3 | kind)
            if kind >= 4 and rng.random() < 0.5:      # the same clip again (collapses) or another length (does not)
                w.append((run if rng.random() < 0.6 else run + 1) << 3 | kind)
        return w

    for a in range(na):
        rec = alns[a]
        rec["ystart"] = rng.integers(0, 1 << 31); rec["tx_ystart"] = rng.integers(0, 100000)
        rec["score"] = rng.integers(-50, 200); rec["ref_id"] = rng.integers(0, n_refs)
        rec["aln_type"] = rng.integers(0, 3)
        rec["tx_or_gene_idx"] = rng.integers(0, n_txs) if rec["aln_type"] == 0 else rng.integers(0, n_genes) if rec["aln_type"] == 1 else 0xFFFFFFFF
        rec["primary"] = rng.integers(0, 2); rec["strand"] = rng.integers(0, 2)
        w = op_list(); rec["ops_off"] = len(ops); rec["ops_len"] = len(w); ops += w
        w = op_list(); rec["tx_ops_off"] = len(ops); rec["tx_ops_len"] = len(w); ops += w
    ops = np.array(ops + [0], np.uint32)
    firsts = np.ascontiguousarray(first[:-1])
    res = api._Result(n, na, len(ops), firsts.ctypes.data, count.ctypes.data, alns.ctypes.data, ops.ctypes.data, 0, 0, 0, 0)
    out, ln = C.c_void_p(), C.c_size_t()
    assert L.tg_format_batch(ix._h, C.byref(res), api._p(bases), api._p(offs), api._p(names), api._p(name_offs),
                             api._p(quals), api._p(qual_offs), 1, C.byref(out), C.byref(ln)) == 0
    host = C.string_at(out, ln.value)
    L.tg_free(out)
    dev = ht.format_sam(ht.HostIndex(fa, gtf), res, bases, offs, names, name_offs, quals, qual_offs)
    assert host.count(b"\n") == na + int((count == 0).sum())
    assert dev == host
