"""Compact results (tg_align_batch_compact) and the one-process multi-GPU entry (tg_multi_*) against the CPU oracle.
The multi tests run on ONE GPU too (two or three contexts on device 0: the shard / segment / rebase logic is the same) and
additionally over every visible GPU when the box has more than one.  Run with `pytest -m gpu`."""
import numpy as np
import pytest

import ht
from common import small_world
from oracle import orc
from thermite_b200 import AlignOpts, Aligner, Index, MultiAligner, compact_arrays, synth

pytestmark = pytest.mark.gpu


def _cmp(gpu_res, orc_res, n):
    d = ht.compare_alignments(dict(first=gpu_res.first, count=gpu_res.count, alns=gpu_res.alns, ops=gpu_res.ops), orc_res, n)
    assert not d, d[:5]


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _world(seed, n, L=91, **kw):
    contigs, gtf, txs, fa = small_world(seed)
    bases, offs = synth.make_reads(seed + 5, contigs, txs, n, L=L, sub=0.02, ins=0.003, dele=0.003, polya_frac=0.15,
                                   polya_len=(10, 35))
    return fa, gtf, bases, offs


@pytest.mark.parametrize("seed", [1, 4])
def test_compact_records_expand_to_the_wide_records(seed):
    fa, gtf, bases, offs = _world(seed, 3000)
    opts = AlignOpts(20, 0.0, 30, 1, True)
    ix = Index.create_from_memory(fa, gtf)
    al = Aligner(ix, opts)
    wide = al.align_reads(bases, offs)
    comp = al.align_reads_compact(bases, offs)
    n = len(offs) - 1
    assert np.array_equal(wide.count, comp.count) and comp.counters["n_segments"] == 1
    for f in wide.alns.dtype.names:
        if f not in ("ops_off", "tx_ops_off"):
            # same pool order is not promised; compare per read below.  Here: multiset of field values must agree
            assert np.array_equal(np.sort(wide.alns[f]), np.sort(comp.alns[f])), f
    ores = orc.Index.create(fa, gtf).align_batch(bases, offs, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    _cmp(comp, ores, n)
    # the small-batch path (no early output, no item sort) writes compact records as well
    comp_small = al.align_reads_compact(bases[: int(offs[50])], offs[:51])
    _cmp(comp_small, ores, 50)


@pytest.mark.parametrize("slots", [2, 3])
def test_multi_on_one_gpu_equals_oracle_in_read_order(slots, monkeypatch):
    """Shards, pool segments and on-device rebasing with several contexts on device 0; TG_SMALL_BATCH=0 forces the
    large-batch path (early output after rounds 0 and 1, fix-up list) even at test size."""
    monkeypatch.setenv("TG_SMALL_BATCH", "0")
    fa, gtf, bases, offs = _world(2, 5000)
    n = len(offs) - 1
    ix = Index.create_from_memory(fa, gtf)
    oix = orc.Index.create(fa, gtf)
    kw = dict(k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    ores = oix.align_batch(bases, offs, **kw)
    m = MultiAligner(ix, AlignOpts(20, 0.0, 30, 1, True), devices=[0] * slots)
    assert m.replication()[0] == "single"
    for rep in range(2):  # second call: segment sizes come from the first call's record rates
        res = m.align_reads(bases, offs)
        assert res.counters["n_segments"] == slots
        _cmp(res, ores, n)
    # raw view: every read's records lie inside its shard's segment, in shard order
    raw = m.align_reads_raw(bases.ctypes.data, offs.ctypes.data, n)
    first, count, alns, ops = compact_arrays(raw)
    assert raw.n_alns == int(count.sum()) and raw.alns_extent >= raw.n_alns
    bounds = [n * g // slots for g in range(slots + 1)]
    prev_hi = 0
    for g in range(slots):
        f, c = first[bounds[g]: bounds[g + 1]].astype(np.int64), count[bounds[g]: bounds[g + 1]].astype(np.int64)
        used = f[c > 0]
        if len(used):
            assert used.min() >= prev_hi
            prev_hi = int((f + c)[c > 0].max())
    # a batch whose segments are far too small for it: the call re-runs with the sizes the shards report
    g = synth.make_genome(31, 20000, families=((60, 150, 0.0, 0.04), (20, 300, 0.0, 0.02)), polya_runs=10, polya_len=(20, 40))
    contigs = [("chrR", g)]
    gtf2, txs2 = synth.make_annotation(32, "chrR", g, n_genes=8, tx_per_gene=(1, 4), exons_per_tx=(1, 5), exon_len=(20, 150),
                                       intron_len=(20, 300), lead=0, prefix="r")
    fa2 = synth.fasta_bytes(contigs)
    n2 = 60000
    b2, o2 = synth.make_reads(33, contigs, txs2, n2, L=91, sub=0.03, ins=0.004, dele=0.004)
    ix2 = Index.create_from_memory(fa2, gtf2)
    m2 = MultiAligner(ix2, AlignOpts(12, 0.0, 20, 1, True), devices=[0] * slots)
    res2 = m2.align_reads(b2, o2)
    ores2 = orc.Index.create(fa2, gtf2).align_batch(b2, o2, n_threads=8, k=12, pct=0.0, min_score=20, score_range=1, intron_mode=True)
    assert res2.counters["n_ops"] > 12.5 * n2  # ~16 operation words per read: the default segments (8 per read + 12 %) cannot hold them
    _cmp(res2, ores2, n2)
    m.close(); m2.close()


def test_multi_over_all_gpus_equals_oracle_in_read_order(monkeypatch):
    """The real thing: one replica per GPU made by ONE NCCL broadcast, one shard per GPU, one result in read order."""
    g = _n_gpus()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    monkeypatch.setenv("TG_SMALL_BATCH", "0")
    fa, gtf, bases, offs = _world(3, 8000)
    n = len(offs) - 1
    ix = Index.create_from_memory(fa, gtf)
    ores = orc.Index.create(fa, gtf).align_batch(bases, offs, n_threads=8, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)
    m = MultiAligner(ix, AlignOpts(20, 0.0, 30, 1, True), devices=list(range(g)))
    how, ms = m.replication()
    assert how in ("nccl", "peer-copy")
    res = m.align_reads(bases, offs)
    assert res.counters["n_segments"] == g
    _cmp(res, ores, n)
    one = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True), device=0).align_reads_compact(bases, offs)
    assert np.array_equal(one.count, res.count)
    for k in ("swg_extensions", "seed_hits", "n_smems", "n_alns", "n_ops"):
        assert one.counters[k] == res.counters[k], k
    m.close()
    # the same without NCCL (device-to-device copies)
    monkeypatch.setenv("TG_MULTI_NO_NCCL", "1")
    m = MultiAligner(ix, AlignOpts(20, 0.0, 30, 1, True), devices=list(range(g)))
    assert m.replication()[0] == "peer-copy"
    _cmp(m.align_reads(bases, offs), ores, n)
    m.close()


def test_file_pipeline_over_tg_multi_equals_oracle_text(tmp_path):
    """tg_align_files with a tg_multi (two contexts here; every visible GPU when there are several): the reads of every
    batch are sharded, the text is written in input order and equals the oracle's PAF text."""
    import gzip
    from thermite_b200 import OutputFormat, align_reads_from_file
    fa, gtf, bases, offs = _world(6, 2500)
    lines = []
    for i in range(len(offs) - 1):
        seq = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        lines.append(b"@q%d\n%s\n+\n%s\n" % (i, seq, b"I" * len(seq)))
    fq = b"".join(lines)
    (tmp_path / "q.fastq.gz").write_bytes(gzip.compress(fq))
    ix = Index.create_from_memory(fa, gtf)
    want = orc.Index.create(fa, gtf).align_fastq_text(fq, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True, sam=False)
    devs = list(range(_n_gpus())) if _n_gpus() > 1 else [0, 0]
    st = align_reads_from_file(ix, [str(tmp_path / "q.fastq.gz")], str(tmp_path / "o.paf"), OutputFormat.Paf,
                               AlignOpts(20, 0.0, 30, 1, True), batch_reads=333, devices=devs)
    assert st["n_reads"] == 2500 and st["n_batches"] == 8
    assert (tmp_path / "o.paf").read_bytes() == want
