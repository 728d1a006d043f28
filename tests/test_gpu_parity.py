"""Parity of the CUDA path (through the C ABI of libthermite_gpu.so) against the CPU oracle.
Everything here needs a B200: run with `pytest -m gpu`."""
import numpy as np
import pytest

import ht
from common import bam_to_sam, golden, reads_to_batch, small_world, swg_pairs
from oracle import orc
from thermite_b200 import AlignOpts, Aligner, Index, ThermiteAligner, parse_fastq, sam_header, suffix_array_gpu, synth

pytestmark = pytest.mark.gpu


def _cmp(gpu_res, orc_res, n):
    d = ht.compare_alignments(dict(first=gpu_res.first, count=gpu_res.count, alns=gpu_res.alns, ops=gpu_res.ops),
                              orc_res, n)
    assert not d, d[:5]


def test_test_dataset_paf_sam_text():
    """config 1: data/test_query.fastq vs data/test_ref.fasta + .gtf, flags -k3 --min-aln-score=0 (data/Makefile:21)."""
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    oix = orc.Index.create(fa, gtf)
    ix = Index.create_from_memory(fa, gtf)
    al = Aligner(ix, AlignOpts(min_seed_len=3, min_aln_score=0))
    bases, offs, names, name_offs, quals, qual_offs = parse_fastq(fq)
    res = al.align_reads_raw(bases.ctypes.data, offs.ctypes.data, len(offs) - 1)
    paf = al.format_result_raw(res, bases, offs, names, name_offs, quals, qual_offs, sam=False)
    assert paf == oix.align_fastq_text(fq, k=3, min_score=0, sam=False)
    assert paf == golden("test_query.paf")
    sam = sam_header(ix) + al.format_result_raw(res, bases, offs, names, name_offs, quals, qual_offs, sam=True)
    assert sam == oix.align_fastq_text(fq, k=3, min_score=0, sam=True)
    assert sam == golden("test_query.sam")


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_worlds_records_and_seeds(seed):
    contigs, gtf, txs, fa = small_world(seed)
    rng = np.random.default_rng(seed + 100)
    L = int(rng.choice([40, 60, 91, 91, 120]))
    k = int(rng.choice([12, 16, 20, 20, 25]))
    pct = float(rng.choice([0.0, 0.5, 0.66]))
    mins = int(rng.choice([0, 20, 30]))
    intron = bool(rng.integers(0, 2))
    srange = int(rng.choice([0, 1, 1, 3]))
    n = 600
    bases, offs = synth.make_reads(seed + 5, contigs, txs, n, L=L, sub=0.02, ins=0.003, dele=0.003, polya_frac=0.15,
                                   polya_len=(10, 35))
    oix = orc.Index.create(fa, gtf)
    ix = Index.create_from_memory(fa, gtf)
    assert (oix.sa() == ix.suffix_array()).all()
    al = Aligner(ix, AlignOpts(k, pct, mins, srange, intron))
    # seeds == Index::all_smems
    seeds, first, count = al.seed_reads(bases, offs)
    sa = ix.suffix_array()
    for r in range(0, n, 7):
        rd = bases[int(offs[r]): int(offs[r + 1])].tobytes()
        assert ht.expand_seeds(seeds, first, count, sa, r) == oix.all_smems(rd, k), (seed, r)
    # records == align_read (default mode: bound-stopped extensions)
    res = al.align_reads(bases, offs)
    oix.counters_reset()
    ores = oix.align_batch(bases, offs, k=k, pct=pct, min_score=mins, score_range=srange, intron_mode=intron)
    _cmp(res, ores, n)
    assert res.counters["swg_cells"] <= oix.counters()["swg_cells"]
    assert res.counters["seed_hits"] == oix.counters()["hits"]
    # exact-cell mode: same records, and the DP cell count equals the reference's
    al.set_exact_cell_count(True)
    res = al.align_reads(bases, offs)
    _cmp(res, ores, n)
    assert res.counters["swg_cells"] == oix.counters()["swg_cells"]


def test_chrM_synthM_records():
    """config 2 stand-in at test size: reads simulated from the bundled chrM, flags -k20 -s0 --intron-mode."""
    fa, gtf = golden("GRCh38-2020-A-chrM.fasta"), golden("GRCh38-2020-A-chrM.gtf")
    oix = orc.Index.create(fa, gtf)
    ix = Index.create_from_memory(fa, gtf)
    g = np.frombuffer(b"".join(fa.split(b"\n")[1:]), np.uint8)
    txs = []
    for t in oix.txs():
        pass
    # transcripts for read simulation straight from the GTF (single-exon genes on chrM)
    for ln in gtf.decode().splitlines():
        f = ln.split("\t")
        if len(f) > 8 and f[2] == "exon":
            txs.append(dict(id="x", strand=f[6], exons=[(int(f[3]) - 1, int(f[4]))], chrom="chrM", gene="g"))
    n = 4000
    bases, offs = synth.make_reads(20211, [("chrM", g)], txs, n, L=91)
    al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
    res = al.align_reads(bases, offs)
    ores = oix.align_batch(bases, offs, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    _cmp(res, ores, n)
    assert res.count.sum() > 0.9 * n
    # the same batch in overlapped chunks (H2D / kernels / D2H pipelined inside tg_align_batch): one contiguous result,
    # identical records and counters
    al.set_chunk_reads(1024)
    res_c = al.align_reads(bases, offs)
    _cmp(res_c, ores, n)
    assert res_c.counters == res.counters
    assert np.array_equal(res_c.count, res.count) and len(res_c.alns) == len(res.alns) and len(res_c.ops) == len(res.ops)


@pytest.mark.parametrize("seed", [21, 22])
def test_repeat_rich_reads_many_hits_many_accepted(seed):
    """Repeat families with near-identical copies and a low k: hundreds of hits per read, batches that are cut and
    re-submitted under predicted states, reads with more than 16 accepted alignments (global finaliser scratch)."""
    g = synth.make_genome(seed, 20000, families=((60, 150, 0.0, 0.04), (20, 300, 0.0, 0.02)), polya_runs=10, polya_len=(20, 40))
    contigs = [("chrR", g)]
    gtf, txs = synth.make_annotation(seed + 1, "chrR", g, n_genes=8, tx_per_gene=(1, 4), exons_per_tx=(1, 5),
                                     exon_len=(20, 150), intron_len=(20, 300), lead=0, prefix="r")
    fa = synth.fasta_bytes(contigs)
    n = 3000
    bases, offs = synth.make_reads(seed + 2, contigs, txs, n, L=91, sub=0.03, ins=0.004, dele=0.004, polya_frac=0.2,
                                   polya_len=(15, 35))
    oix = orc.Index.create(fa, gtf)
    ix = Index.create_from_memory(fa, gtf)
    max_records = 0
    for opts in (AlignOpts(12, 0.0, 20, 1, True), AlignOpts(10, 0.5, 0, 3, True), AlignOpts(14, 0.66, 30, 0, False)):
        kw = dict(k=opts.min_seed_len, pct=opts.min_aln_score_percent, min_score=opts.min_aln_score,
                  score_range=opts.multimap_score_range, intron_mode=opts.intron_mode)
        oix.counters_reset()
        ores = oix.align_batch(bases, offs, n_threads=8, **kw)
        oc = oix.counters()
        al = Aligner(ix, opts)
        res = al.align_reads(bases, offs)
        _cmp(res, ores, n)
        assert res.counters["seed_hits"] == oc["hits"] and oc["hits"] > 5 * n
        al.set_exact_cell_count(True)
        res = al.align_reads(bases, offs)
        _cmp(res, ores, n)
        assert res.counters["swg_cells"] == oc["swg_cells"]
        max_records = max(max_records, int(res.count.max()))
    assert max_records > 16  # more records than the thread-local finaliser holds


def test_edge_reads():
    contigs, gtf, txs, fa = small_world(11)
    oix = orc.Index.create(fa, gtf)
    ix = Index.create_from_memory(fa, gtf)
    g = contigs[0][1]
    reads = [b"", b"A", b"ACGT", g[100:191].tobytes().lower(), b"N" * 91, g[200:260].tobytes() + b"N" * 31,
             b"ACGTRYKM" * 11, g[300:391].tobytes()[::-1], bytes(g[400:450]) + b"A" * 41, b"T" * 200,
             g[500:1000].tobytes()]
    bases, offs = reads_to_batch(reads)
    for opts in (AlignOpts(20, 0.66, 30, 1, False), AlignOpts(10, 0.0, 0, 2, True), AlignOpts(5, 0.5, 0, 0, True)):
        al = Aligner(ix, opts)
        res = al.align_reads(bases, offs)
        ores = oix.align_batch(bases, offs, k=opts.min_seed_len, pct=opts.min_aln_score_percent,
                               min_score=opts.min_aln_score, score_range=opts.multimap_score_range,
                               intron_mode=opts.intron_mode)
        _cmp(res, ores, len(reads))
    # empty batch
    e = al.align_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(e) == 0


def test_swg_batch_matches_oracle():
    """config 5 at test size: (score, xend, yend, ops) byte-identical to SwgExtend::extend, cells counted alike."""
    ix = Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"))
    al = Aligner(ix, AlignOpts(min_seed_len=3))
    for seed, kw in ((1, {}), (2, dict(alphabet=b"AC")), (3, dict(max_x=200, bw_choices=(3, 30, 100)))):
        xs, xo, ys, yo, bw, xd = swg_pairs(seed, 6000, **kw)
        b = orc.swg_extend_batch(xs, xo, ys, yo, bw, xd)
        for exact in (False, True):
            al.set_exact_cell_count(exact)
            a = al.swg_extend_batch(xs, xo, ys, yo, bw, xd)
            for key in ("score", "xend", "yend", "ops_off", "ops"):
                assert np.array_equal(a[key], b[key]), (seed, key, exact)
            assert a["cells"] == b["cells"] if exact else a["cells"] <= b["cells"]
    # reference KATs (src/swg.rs:249-317)
    kats = [(b"AAAAAAAA", b"AAAAAAAA", 1, 1), (b"AAAAATTT", b"AAAAAAAA", 1, 1), (b"AAATAAAA", b"AAAAAAAA", 1, 1),
            (b"AAATTTT", b"AAACCTTTT", 2, 3)]
    xs, xo = reads_to_batch([k[0] for k in kats])
    ys, yo = reads_to_batch([k[1] for k in kats])
    a = al.swg_extend_batch(xs, xo, ys, yo, [k[2] for k in kats], [k[3] for k in kats])
    assert a["score"].tolist() == [8, 5, 6, 4]
    assert a["xend"].tolist() == [8, 5, 8, 7] and a["yend"].tolist() == [8, 5, 8, 9]
    # x_drop < band width is refused (the reference panics / reads stale state there)
    from thermite_b200 import ThermiteError
    with pytest.raises(ThermiteError):
        al.swg_extend_batch(xs, xo, ys, yo, [4] * 4, [1] * 4)


def test_roundtrip_properties_full_size():
    """Size-independent properties on a larger batch: every perfect read maps back to where it came from with
    score L and an all-match CIGAR; alignment spans are consistent with their operations."""
    contigs, gtf, txs = synth.synth21(scale=0.02)
    fa = synth.fasta_bytes(contigs)
    ix = Index.create_from_memory(fa, gtf)
    al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
    n, L = 50_000, 91
    bases, offs = synth.make_reads(7, contigs, txs, n, L=L, sub=0.0, ins=0.0, dele=0.0, polya_frac=0.0, tso_frac=0.0)
    res = al.align_reads(bases, offs)
    assert (res.count >= 1).all()
    prim = res.alns[res.first.astype(np.int64)]
    assert (prim["score"] == L).all() and (prim["primary"] == 1).all()
    assert (prim["xstart"] == 0).all() and (prim["xend"] == L).all()
    # ops consistency for all records
    for a in res.alns[:: max(1, len(res.alns) // 2000)]:
        w = res.ops[int(a["ops_off"]): int(a["ops_off"]) + int(a["ops_len"])]
        kind, run = w & 7, w >> 3
        yspan = int(run[(kind <= 2) | (kind == 5)].sum())
        xspan = int(run[(kind <= 1) | (kind == 3)].sum())
        assert yspan == int(a["yend"] - a["ystart"]) and xspan == int(a["xend"] - a["xstart"])


def _pack4(codes):
    n = len(codes)
    c = np.zeros((n // 16 + 4) * 16, np.uint64)
    c[:n] = codes
    sh = np.uint64(4) * (np.uint64(15) - np.arange(16, dtype=np.uint64))
    return (c.reshape(-1, 16) << sh).sum(axis=1, dtype=np.uint64)


def test_suffix_array_gpu_matches_host_builder():
    """the suffix array built on the GPU (prefix doubling over unresolved suffixes, csrc/tg_sa.cu) is the
    array the host SA-IS and the oracle's independent sorter give (divsufsort64 order, src/index.rs:103-105), and the
    index blob built around it is byte-identical."""
    fa, gtf = golden("test_ref.fasta"), golden("test_ref.gtf")
    ix = Index.create_from_memory(fa, gtf)
    sa, _, _ = suffix_array_gpu(ix.text4(), ix.text_len())
    assert np.array_equal(sa, ix.suffix_array()) and np.array_equal(sa, orc.Index.create(fa, gtf).sa())
    assert np.array_equal(Index.create_from_memory(fa, gtf, sa_device=0).blob(), ix.blob())
    fa, gtf = golden("GRCh38-2020-A-chrM.fasta"), golden("GRCh38-2020-A-chrM.gtf")
    ix = Index.create_from_memory(fa, gtf)
    assert np.array_equal(Index.create_from_memory(fa, gtf, sa_device=0).blob(), ix.blob())
    for seed in (1, 2):
        contigs, gtf2, txs, fa2 = small_world(seed)
        ix = Index.create_from_memory(fa2, gtf2)
        sa, _, steps = suffix_array_gpu(ix.text4(), ix.text_len())
        assert np.array_equal(sa, ix.suffix_array()), seed


@pytest.mark.parametrize("kind", ["random", "n_runs", "periodic", "one_symbol", "tiny"])
def test_suffix_array_gpu_adversarial_texts(kind):
    """Texts that keep prefix doubling busy: long N runs on both strands (chr21 starts with 5 Mb of N), tandem repeats,
    a single repeated symbol (every step halves nothing but the last one), and texts shorter than one 16-symbol key.
    Checked by the suffix-array property itself: a permutation whose neighbours are in strictly ascending order
    (verified with rank-based comparison of the following suffixes) -- and against a naive sort when small."""
    rng = np.random.default_rng(7)
    if kind == "random":
        codes = rng.integers(1, 6, 300_000)
        codes[rng.integers(0, len(codes), 20)] = 0
    elif kind == "n_runs":
        body = rng.integers(1, 6, 100_000)
        body[body == 4] = 1
        fwd = np.concatenate([np.full(150_000, 4), body, np.full(20_000, 4), body[:30_000]])
        rev = np.array([0, 5, 3, 2, 4, 1])[fwd[::-1]]
        codes = np.concatenate([fwd, [0], rev, [0]])
    elif kind == "periodic":
        unit = rng.integers(1, 6, 37)
        codes = np.concatenate([np.tile(unit, 3000), [0], np.tile(unit[::-1], 2500), [0]])
    elif kind == "one_symbol":
        codes = np.full(70_001, 1)
    else:
        codes = np.array([1, 5, 1, 0, 5, 1, 5, 0])
    codes = codes.astype(np.uint64)
    n = len(codes)
    sa, ms, steps = suffix_array_gpu(_pack4(codes), n)
    assert np.array_equal(np.sort(sa), np.arange(n, dtype=np.uint32)), "not a permutation"
    # SA property: text[sa[p]:] < text[sa[p+1]:]  <=>  (c[a], rank[a+1]) < (c[b], rank[b+1]) with rank(n) = -1
    rank = np.empty(n + 1, np.int64)
    rank[sa] = np.arange(n)
    rank[n] = -1
    a, b = sa[:-1].astype(np.int64), sa[1:].astype(np.int64)
    ca, cb = codes[a].astype(np.int64), codes[b].astype(np.int64)
    ok = (ca < cb) | ((ca == cb) & (rank[a + 1] < rank[b + 1]))
    assert ok.all(), (kind, int(np.argmin(ok)))
    if n <= 1000:
        t = bytes((codes + 1).astype(np.uint8))
        assert list(sa) == sorted(range(n), key=lambda i: t[i:])


def test_thermite_aligner_per_read_calls_from_threads():
    """ThermiteAligner::align_read (src/wrapper.rs:72) -- one read per call, from several host threads, and
    pipelined submit / wait -- returns for every read exactly the records the batch call gives (and the oracle)."""
    import threading
    contigs, gtf2, txs, fa2 = small_world(4)
    n = 600
    b2, o2 = synth.make_reads(11, contigs, txs, n, L=91, sub=0.02, ins=0.003, dele=0.003)
    ix = Index.create_from_memory(fa2, gtf2)
    opts = AlignOpts(20, 0.0, 30, 1, True)
    ores = orc.Index.create(fa2, gtf2).align_batch(b2, o2, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    reads = [bytes(b2[int(o2[r]):int(o2[r + 1])]) for r in range(n)]
    ta = ThermiteAligner(ix, opts, max_batch_reads=128, max_wait_us=300)
    got = [None] * n

    def worker(t, nt):
        if t % 2 == 0:
            for r in range(t, n, nt):
                got[r] = ta.align_read_raw(reads[r])
        else:
            mine = list(range(t, n, nt))
            tickets = [ta.submit(reads[r]) for r in mine]
            for r, tk in zip(mine, tickets):
                got[r] = ta.wait_raw(tk)

    th = [threading.Thread(target=worker, args=(t, 6)) for t in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    st = ta.stats()
    assert st["reads"] == n and st["batches"] < n and 1 < st["largest_batch"] <= 128
    # stitch the per-read blocks back into one flat result and compare with the oracle
    first, count, alns, ops, na, no = [], [], [], [], 0, 0
    for r in range(n):
        g = got[r]
        a = g.alns.copy()
        a["ops_off"] += no
        a["tx_ops_off"] += no
        first.append(na); count.append(len(a)); alns.append(a); ops.append(g.ops)
        na += len(a); no += len(g.ops)
    flat = dict(first=np.array(first, np.uint64), count=np.array(count, np.uint32), alns=np.concatenate(alns),
                ops=np.concatenate(ops))
    d = ht.compare_alignments(flat, ores, n)
    assert not d, d[:5]
    assert ta.align_read(reads[0]) == Aligner(ix, opts).align_read(reads[0])
    assert ta.align_read(b"") == [] and ta.align_read(b"ACGT") == []
    with pytest.raises(Exception):
        ta.wait(10 ** 9)
    ta.close()


@pytest.mark.parametrize("small_batch", ["0", "1000000000"])
def test_large_and_small_batch_paths_give_the_same_records(small_batch, monkeypatch):
    """tg_align_batch treats batches below TG_SMALL_BATCH reads (default 32,768) as latency-bound: no early output of
    finished reads, no locus-ordered prep, an earlier "nothing left" check.  Both paths against the oracle on the same
    reads (the environment variable is read when the context is created)."""
    monkeypatch.setenv("TG_SMALL_BATCH", small_batch)
    contigs, gtf2, txs, fa2 = small_world(5)
    n = 3000
    b2, o2 = synth.make_reads(12, contigs, txs, n, L=91, sub=0.02, ins=0.003, dele=0.003)
    al = Aligner(Index.create_from_memory(fa2, gtf2), AlignOpts(20, 0.0, 30, 1, True))
    ores = orc.Index.create(fa2, gtf2).align_batch(b2, o2, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    _cmp(al.align_reads(b2, o2), ores, n)
    for m in (1, 2, 7):   # tiny batches: the host check after round 0 / 1
        _cmp(al.align_reads(b2[: int(o2[m])], o2[: m + 1]), ores, m)


def test_align_reads_from_file_paf_sam_bam(tmp_path):
    """align_reads_from_file (src/aligner.rs:22-120) end to end on the reference's own test files: PAF and SAM equal
    the golden text, the BAM file decodes to the SAM file (header, reference list, records)."""
    from thermite_b200 import OutputFormat, align_reads_from_file
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    qp = tmp_path / "q.fastq"
    qp.write_bytes(fq)
    ix = Index.create_from_memory(fa, gtf)
    opts = AlignOpts(min_seed_len=3, min_aln_score=0)
    outs = {}
    for fmt in (OutputFormat.Paf, OutputFormat.Sam, OutputFormat.Bam):
        op = tmp_path / ("out." + fmt)
        align_reads_from_file(ix, [str(qp)], str(op), fmt, opts, batch_reads=4)   # several batches per file
        outs[fmt] = op.read_bytes()
    assert outs[OutputFormat.Paf] == golden("test_query.paf")
    assert outs[OutputFormat.Sam] == golden("test_query.sam")
    text, refs, lines, n_blocks = bam_to_sam(outs[OutputFormat.Bam])
    assert text + lines == outs[OutputFormat.Sam]
    assert [r[0] for r in refs] == [b"some_ref", b"another_seq", b"introns_seq", b"introns_revcomp"]


def test_align_files_gz_inputs_many_batches(tmp_path):
    """The native file pipeline (tg_align_files: reader / aligner / writers overlapped) on a few thousand reads: gzip and
    BGZF input (the reference's data sets are .fastq.gz, data/Makefile:26,35), two query files, batches far smaller than a
    file so that every buffer set is reused many times; PAF and SAM text equal the oracle's text for the same reads."""
    import gzip
    from test_stream import bgzf
    from thermite_b200 import OutputFormat, align_reads_from_file
    contigs, gtf, txs, fa = small_world(11)
    bases, offs = synth.make_reads(5, contigs, txs, 3000, L=91, sub=0.02, ins=0.003, dele=0.003)
    lines = []
    for i in range(len(offs) - 1):
        seq = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        lines.append(b"@r%d extra words\n%s\n+\n%s\n" % (i, seq, b"F" * len(seq)))
    half = len(lines) // 2
    fq1, fq2 = b"".join(lines[:half]), b"".join(lines[half:])
    (tmp_path / "a.fastq.gz").write_bytes(gzip.compress(fq1))
    (tmp_path / "b.fastq.gz").write_bytes(bgzf(fq2, 20000))
    ix = Index.create_from_memory(fa, gtf)
    opts = AlignOpts(20, 0.0, 30, 1, True)
    oix = orc.Index.create(fa, gtf)
    for fmt, sam in ((OutputFormat.Paf, False), (OutputFormat.Sam, True)):
        op = tmp_path / ("out." + fmt)
        st = align_reads_from_file(ix, [str(tmp_path / "a.fastq.gz"), str(tmp_path / "b.fastq.gz")], str(op), fmt, opts, batch_reads=257)
        assert st["n_reads"] == 3000 and st["n_batches"] >= 12
        want = oix.align_fastq_text(fq1 + fq2, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True, sam=sam)
        assert op.read_bytes() == want, fmt   # (the oracle's SAM text starts with the header, like the file)


def test_contexts_share_one_kmer_table():
    """Two contexts of one index with the same min_seed_len share the k-mer table (reference counted); a third with another
    k builds its own.  Results do not depend on which context built the table or on the order they are closed in."""
    import torch
    contigs, gtf, txs, fa = small_world(4)
    bases, offs = synth.make_reads(9, contigs, txs, 400, L=91, sub=0.02, ins=0.003, dele=0.003)
    ix = Index.create_from_memory(fa, gtf)
    oix = orc.Index.create(fa, gtf)
    ores = oix.align_batch(bases, offs, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    a1 = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
    free1 = torch.cuda.mem_get_info()[0]
    a2 = Aligner(ix, AlignOpts(20, 0.66, 30, 1, False))   # same k, other options: same table
    free2 = torch.cuda.mem_get_info()[0]
    assert a1.kmer_table_bytes() == a2.kmer_table_bytes() > 0
    assert free1 - free2 < a1.kmer_table_bytes() + (64 << 20)  # a context's own allocations, but no second table
    a3 = Aligner(ix, AlignOpts(12, 0.0, 30, 1, True))
    _cmp(a1.align_reads(bases, offs), ores, 400)
    a1.close()                                              # the table lives on for a2
    r2 = a2.align_reads(bases, offs)
    _cmp(r2, oix.align_batch(bases, offs, k=20, pct=0.66, min_score=30, score_range=1, intron_mode=False), 400)
    a4 = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
    a2.close()
    _cmp(a4.align_reads(bases, offs), ores, 400)
    _cmp(a3.align_reads(bases, offs), oix.align_batch(bases, offs, k=12, pct=0.0, min_score=30, score_range=1, intron_mode=True), 400)


def test_paf_and_sam_written_on_the_gpu_equal_the_host_writer(tmp_path, monkeypatch):
    """tg_paf_* / tg_sam_create (csrc/tg_paf.cu, csrc/tg_textfmt.h): PAF lines and SAM records formatted on the device are
    byte-identical to the host writer's and to the oracle's text -- test data (config 1), a world with both strands /
    multi-mappers / unmapped reads, long read names with spaces -- and tg_align_files gives the same file with either
    writer (TG_PAF_HOST=1 selects the host one)."""
    import ctypes as C
    from thermite_b200 import OutputFormat, align_reads_from_file, lib, sam_header
    from thermite_b200.api import _ReadBatch, _Result, _check
    cases = []
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    cases.append((fa, gtf, fq, AlignOpts(min_seed_len=3, min_aln_score=0), dict(k=3, min_score=0)))
    contigs, gtf2, txs, fa2 = small_world(7)
    b2, o2 = synth.make_reads(3, contigs, txs, 4000, L=91, sub=0.03, ins=0.004, dele=0.004, polya_frac=0.2, polya_len=(10, 40))
    fq2 = b"".join(b"@read_%d/1 a rather long header with spaces %d\n%s\n+\n%s\n" %
                   (i, i * 7919, b2[int(o2[i]):int(o2[i + 1])].tobytes(),
                    bytes(33 + (i + 3 * k) % 40 for k in range(int(o2[i + 1] - o2[i])))) for i in range(len(o2) - 1))
    cases.append((fa2, gtf2, fq2, AlignOpts(20, 0.0, 30, 3, True), dict(k=20, pct=0.0, min_score=30, score_range=3, intron_mode=True)))
    for fa_, gtf_, fq_, opts, okw in cases:
        ix = Index.create_from_memory(fa_, gtf_)
        al = Aligner(ix, opts)
        oix = orc.Index.create(fa_, gtf_)
        bases, offs, names, name_offs, quals, qual_offs = parse_fastq(fq_)
        batch = _ReadBatch(len(offs) - 1, 0, bases.ctypes.data, offs.ctypes.data, names.ctypes.data, name_offs.ctypes.data,
                           quals.ctypes.data, qual_offs.ctypes.data)
        qp = tmp_path / "q.fastq"
        qp.write_bytes(fq_)
        for sam in (False, True):
            want = oix.align_fastq_text(fq_, sam=sam, **okw)
            head = sam_header(ix) if sam else b""
            f = C.c_void_p()
            _check((lib().tg_sam_create if sam else lib().tg_paf_create)(ix._h, al._h, 0, C.byref(f)))
            for _ in range(3):  # the two text buffers alternate
                text, n = C.c_void_p(), C.c_size_t()
                res = _Result()
                _check(lib().tg_paf_align_batch(f, C.byref(batch), C.byref(text), C.byref(n), C.byref(res)))
                assert head + C.string_at(text, n.value) == want
            lib().tg_paf_destroy(f)
            for host in ("0", "1"):
                if host == "1":
                    monkeypatch.setenv("TG_PAF_HOST", "1")
                else:
                    monkeypatch.delenv("TG_PAF_HOST", raising=False)
                op = tmp_path / ("o%s.%s" % (host, "sam" if sam else "paf"))
                st = align_reads_from_file(ix, [str(qp)], str(op), OutputFormat.Sam if sam else OutputFormat.Paf, opts, batch_reads=777)
                assert op.read_bytes() == want and st["n_reads"] == len(offs) - 1
        monkeypatch.delenv("TG_PAF_HOST", raising=False)


def test_file_pipeline_names_a_read_that_is_too_long(tmp_path):
    """Limits (include/thermite_gpu.h): reads of more than TG_MAX_READ_LEN bases are not aligned by the device path.  The file
    pipeline says which read it was, the batches before it are in the output and every stage comes to rest."""
    from thermite_b200 import OutputFormat, ThermiteError, align_reads_from_file
    fa, gtf, fq = golden("test_ref.fasta"), golden("test_ref.gtf"), golden("test_query.fastq")
    ix = Index.create_from_memory(fa, gtf)
    opts = AlignOpts(min_seed_len=3, min_aln_score=0)
    qp = tmp_path / "q.fastq"
    qp.write_bytes(fq + b"@the_long_one extra\n" + b"ACGT" * 150 + b"\n+\n" + b"F" * 600 + b"\n" + fq)
    for fmt in (OutputFormat.Paf, OutputFormat.Sam, OutputFormat.Bam):
        with pytest.raises(ThermiteError, match="the_long_one extra.* has 600 bases"):
            align_reads_from_file(ix, [str(qp)], str(tmp_path / "o.out"), fmt, opts, batch_reads=4)


def test_file_pipeline_bam_from_all_cores(tmp_path):
    """BAM through tg_align_files with a batch large enough for the threaded writers (>= 32768 reads): every thread formats its
    reads, encodes them as BAM records and compresses its own BGZF blocks; the file decodes to exactly the SAM file."""
    from thermite_b200 import OutputFormat, align_reads_from_file
    contigs, gtf, txs, fa = small_world(12)
    n = 40000
    bases, offs = synth.make_reads(6, contigs, txs, n, L=91, sub=0.02, ins=0.003, dele=0.003)
    rows = bases.reshape(n, 91)
    fq = b"".join(b"@q%d some words\n%s\n+\n%s\n" % (i, rows[i].tobytes(), bytes(33 + (i + k) % 41 for k in range(91))) for i in range(n))
    qp = tmp_path / "q.fastq"
    qp.write_bytes(fq)
    ix = Index.create_from_memory(fa, gtf)
    opts = AlignOpts(20, 0.0, 30, 2, True)
    outs = {}
    for fmt in (OutputFormat.Sam, OutputFormat.Bam):
        op = tmp_path / ("out." + fmt)
        st = align_reads_from_file(ix, [str(qp)], str(op), fmt, opts, batch_reads=n)
        assert st["n_reads"] == n and st["n_batches"] == 1
        outs[fmt] = op.read_bytes()
    text, refs, lines, n_blocks = bam_to_sam(outs[OutputFormat.Bam])
    assert text + lines == outs[OutputFormat.Sam] and n_blocks > 16
