"""The kernel LOGIC (thermite_b200/csrc/tg_core.h, the code the GPU runs) compiled for the host under an emulated
warp (1 lane, and 32 lanes on threads with barrier-backed shuffles) against the CPU oracle.  The GPU itself is
covered by tests/test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

import ht
from common import golden, reads_to_batch, small_world, swg_pairs
from oracle import orc
from thermite_b200 import synth


def test_product_suffix_array_matches_oracle_and_is_sorted():
    for seed in (1, 2):
        contigs, gtf, txs, fa = small_world(seed)
        hix, oix = ht.HostIndex(fa, gtf), orc.Index.create(fa, gtf)
        sa = hix.sa()
        assert (sa == oix.sa()).all()
        text = oix.text().tobytes()
        assert sorted(sa.tolist()) == list(range(len(text)))
        for i in range(0, len(sa) - 1, 37):
            assert text[sa[i]:] < text[sa[i + 1]:]


def test_seeds_equal_all_smems_test_dataset():
    fa, gtf = golden("test_ref.fasta"), golden("test_ref.gtf")
    hix, oix = ht.HostIndex(fa, gtf), orc.Index.create(fa, gtf)
    reads = [b"ATT", b"CGAT", b"A" * 24, b"AATCGGCTTTT", b"AATTTTT", b"AATCGGCTCTT", b"AACTTTTT", b"AACCCCTT",
             b"AAAGCCGATT", b"AATGCCGATT", b"", b"N", b"ccccc"]
    bases, offs = reads_to_batch(reads)
    for k in (1, 2, 3, 5):
        ctx = ht.HostCtx(hix, k=k, min_score=0)
        pool, first, count = ctx.seed_batch(bases, offs)
        sa = hix.sa()
        for r, rd in enumerate(reads):
            assert ht.expand_seeds(pool, first, count, sa, r) == oix.all_smems(rd, k), (k, rd)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_align_records_equal_oracle_random_worlds(seed):
    contigs, gtf, txs, fa = small_world(seed)
    rng = np.random.default_rng(seed + 100)
    L = int(rng.choice([40, 60, 91, 91, 120]))
    k = int(rng.choice([12, 16, 20, 20, 25]))
    pct = float(rng.choice([0.0, 0.5, 0.66]))
    mins = int(rng.choice([0, 20, 30]))
    intron = bool(rng.integers(0, 2))
    srange = int(rng.choice([0, 1, 1, 3]))
    n = 200
    bases, offs = synth.make_reads(seed + 5, contigs, txs, n, L=L, sub=0.02, ins=0.003, dele=0.003, polya_frac=0.15,
                                   polya_len=(10, 35))
    hix, oix = ht.HostIndex(fa, gtf), orc.Index.create(fa, gtf)
    ctx = ht.HostCtx(hix, k=k, pct=pct, min_score=mins, score_range=srange, intron_mode=intron)
    pool, first, count = ctx.seed_batch(bases, offs)
    pool0, first0, count0 = ctx.seed_batch(bases, offs, lanes=0)  # the device's pack / probe / select pipeline
    sa = hix.sa()
    for r in range(0, n, 5):
        rd = bases[int(offs[r]): int(offs[r + 1])].tobytes()
        want = oix.all_smems(rd, k)
        assert ht.expand_seeds(pool, first, count, sa, r) == want
        assert ht.expand_seeds(pool0, first0, count0, sa, r) == want
    res = ctx.align_batch(bases, offs, lanes=1)
    oix.counters_reset()
    ores = oix.align_batch(bases, offs, k=k, pct=pct, min_score=mins, score_range=srange, intron_mode=intron)
    d = ht.compare_alignments(res, ores, n)
    assert not d, d[:3]
    assert res["flags"] == 0 and res["cells"] == oix.counters()["swg_cells"] and res["hits"] == oix.counters()["hits"]
    # bound-stopped extensions (the product default): identical records, never more cells
    resb = ctx.align_batch(bases, offs, lanes=1, bound_stop=True)
    assert not ht.compare_alignments(resb, ores, n) and resb["cells"] <= res["cells"]
    # the speculative round pipeline (what the GPU runs by default): identical records AND identical work counters
    # rounds=2 additionally runs every eligible extension through the thread-per-extension DP (tg_dpt.h)
    for bs in (False, True):
        for mode in (1, 2):
            resr = ctx.align_batch(bases, offs, lanes=1, bound_stop=bs, rounds=mode)
            d = ht.compare_alignments(resr, ores, n)
            assert not d, (mode, d[:3])
            assert resr["flags"] == 0 and resr["hits"] == res["hits"] and resr["n_ext"] == res["n_ext"]
            assert resr["cells"] == (resb["cells"] if bs else res["cells"]), mode
    # compact records (tg_aln_c) rebased like a shard of tg_multi_align_batch, expanded back by tg_aln_expand: both writers
    for mode in (0, 1):
        resc = ctx.align_batch(bases, offs, lanes=1, bound_stop=True, rounds=mode, compact=(1000 * seed + 7, 5000 * seed + 3))
        assert resc["flags"] == 0 and not ht.compare_alignments(resc, ores, n), mode
    # the 32-lane wavefront (what the GPU executes) on a slice
    m = 12
    for bs in (False, True):
        res32 = ctx.align_batch(bases[: int(offs[m])], offs[: m + 1], lanes=32, bound_stop=bs)
        d = ht.compare_alignments(res32, ores, m)
        assert not d, d[:3]


def test_swg_wavefront_equals_oracle():
    # lanes = 0: the thread-per-extension DP of tg_dpt.h (every band class) where eligible, else the 1-lane wavefront
    for seed, lanes, n, kw in ((1, 1, 3000, {}), (2, 32, 150, {}), (3, 32, 60, dict(max_x=200, bw_choices=(3, 30, 100))),
                               (4, 1, 1500, dict(alphabet=b"AC")), (5, 0, 6000, {}), (6, 0, 3000, dict(alphabet=b"AC")),
                               (7, 0, 3000, dict(max_x=128, bw_choices=(0, 1, 3, 5, 11, 19, 27, 35, 39, 50, 64))),
                               (8, 0, 2000, dict(alphabet=b"ACGTN", bw_choices=(2, 7, 15, 16, 23, 40)))):
        xs, xo, ys, yo, bw, xd = swg_pairs(seed, n, **kw)
        b = orc.swg_extend_batch(xs, xo, ys, yo, bw, xd)
        for bs in (False, True):
            a = ht.swg_extend_batch(xs, xo, ys, yo, bw, xd, lanes=lanes, bound_stop=bs)
            for key in ("score", "xend", "yend", "ops_off", "ops"):
                assert np.array_equal(a[key], b[key]), (seed, key, bs)
            assert a["cells"] <= b["cells"] if bs else a["cells"] == b["cells"]


def test_chrM_reads_records():
    fa, gtf = golden("GRCh38-2020-A-chrM.fasta"), golden("GRCh38-2020-A-chrM.gtf")
    hix, oix = ht.HostIndex(fa, gtf), orc.Index.create(fa, gtf)
    g = np.frombuffer(b"".join(fa.split(b"\n")[1:]), np.uint8)
    txs = [dict(id="x", strand=f[6], exons=[(int(f[3]) - 1, int(f[4]))], chrom="chrM", gene="g")
           for f in (ln.split("\t") for ln in gtf.decode().splitlines()) if len(f) > 8 and f[2] == "exon"]
    n = 400
    bases, offs = synth.make_reads(20211, [("chrM", g)], txs, n, L=91)
    for flags in (dict(k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True),
                  dict(k=20, pct=0.66, min_score=30, score_range=1, intron_mode=False)):
        ctx = ht.HostCtx(hix, **flags)
        res = ctx.align_batch(bases, offs)
        ores = oix.align_batch(bases, offs, **flags)
        assert not ht.compare_alignments(res, ores, n)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_round_pipeline_repeat_rich_reads(seed):
    """Many hits per read (repeat families with near-identical copies, low k): batches of hits are evaluated
    speculatively and re-planned when an accepted hit narrows the band; records and counters must not change."""
    rng = np.random.default_rng(seed)
    g = synth.make_genome(seed, 6000, families=((25, 150, 0.0, 0.06), (8, 300, 0.0, 0.03)), polya_runs=6, polya_len=(20, 40))
    contigs = [("chrR", g)]
    gtf, txs = synth.make_annotation(seed + 1, "chrR", g, n_genes=5, tx_per_gene=(1, 4), exons_per_tx=(1, 5),
                                     exon_len=(20, 150), intron_len=(20, 300), lead=0, prefix="r")
    fa = synth.fasta_bytes(contigs)
    n = 150
    bases, offs = synth.make_reads(seed + 2, contigs, txs, n, L=int(rng.choice([60, 91])), sub=0.03, ins=0.004, dele=0.004,
                                   polya_frac=0.2, polya_len=(15, 35))
    hix, oix = ht.HostIndex(fa, gtf), orc.Index.create(fa, gtf)
    for flags in (dict(k=12, pct=0.0, min_score=20, score_range=1, intron_mode=True),
                  dict(k=10, pct=0.5, min_score=0, score_range=3, intron_mode=True),
                  dict(k=14, pct=0.66, min_score=30, score_range=0, intron_mode=False)):
        ctx = ht.HostCtx(hix, **flags)
        oix.counters_reset()
        ores = oix.align_batch(bases, offs, **flags)
        oc = oix.counters()
        assert oc["hits"] > 3 * n  # the point of this world
        res2 = ctx.align_batch(bases, offs, lanes=1, rounds=2)
        assert not ht.compare_alignments(res2, ores, n) and res2["cells"] == oc["swg_cells"]
        res = ctx.align_batch(bases, offs, lanes=1, rounds=True)
        d = ht.compare_alignments(res, ores, n)
        assert not d, d[:3]
        assert res["flags"] == 0 and res["cells"] == oc["swg_cells"] and res["hits"] == oc["hits"]
        # speculation really happened: several rounds, and some evaluated hits were discarded and redone
        assert res["rounds"] >= 3 and res["items"] > res["hits"], (res["rounds"], res["items"], res["hits"])


def test_micro_batcher_hands_every_caller_its_own_records():
    """SURVEY 8f N4 (ThermiteAligner::align_read from many threads, src/wrapper.rs:20-27): tg_batcher's queueing,
    batch cutting, per-read result blocks (operations rebased, transcript and genome words in either order), ticket
    life cycle and error propagation, driven by 1-16 threads over a stand-in batch aligner whose records are a
    function of the read bytes."""
    bad, batches, largest, failed = ht.batcher_selftest(1, 200, 64, 0)
    assert (bad, batches, largest, failed) == (0, 200, 1, 0)      # one blocking caller: every read is its own batch
    bad, batches, largest, failed = ht.batcher_selftest(16, 1500, 256, 200)
    assert bad == 0 and failed == 0 and 1 < largest <= 256 and batches < 16 * 1500
    bad, batches, largest, failed = ht.batcher_selftest(8, 1000, 7, 50)
    assert bad == 0 and failed == 0 and largest == 7              # the batch limit cuts the queue
    bad, batches, largest, failed = ht.batcher_selftest(8, 600, 64, 100, fail_every=5)
    assert bad == 0 and failed > 0                                # a failed batch fails exactly its reads, nothing leaks through
