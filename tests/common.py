"""Shared fixtures/builders for the parity tests (no reference tree access: /root/reference does not exist
on the GPU box -- the reference's tiny fixtures are committed under tests/golden/)."""
import os

import numpy as np

from thermite_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name, mode="rb"):
    with open(os.path.join(GOLDEN, name), mode) as f:
        return f.read()


def reads_to_batch(reads):
    bases = np.frombuffer(b"".join(reads), np.uint8)
    offs = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    return bases, offs


def small_world(seed, n_contigs=3, n_genes=4, min_len=3000, max_len=9000):
    """Multi-contig genome with repeats / N / poly-A and a multi-exon annotation on both strands."""
    rng = np.random.default_rng(seed)
    contigs, gtfs, txs = [], [], []
    for c in range(n_contigs):
        g = synth.make_genome(seed * 10 + c, int(rng.integers(min_len, max_len)), lead_n=int(rng.integers(0, 50)),
                              n_blocks=1, n_block_len=30, families=((6, 120, 0.0, 0.1), (3, 400, 0.0, 0.05)),
                              polya_runs=4, polya_len=(20, 40))
        contigs.append((f"chr{c}", g))
        gt, tx = synth.make_annotation(seed * 7 + c, f"chr{c}", g, n_genes=n_genes, tx_per_gene=(1, 3),
                                       exons_per_tx=(1, 6), exon_len=(15, 120), intron_len=(20, 400), lead=60,
                                       prefix=f"c{c}")
        gtfs.append(gt)
        txs += tx
    return contigs, b"".join(gtfs), txs, synth.fasta_bytes(contigs)


def swg_pairs(seed, n, bw_choices=(1, 2, 4, 8, 16, 31, 61), max_x=71, alphabet=b"ACGT"):
    """Config-5 style pairs: y = mutated copy of x padded with random bases, plus unrelated pairs and empties."""
    rng = np.random.default_rng(seed)
    al = np.frombuffer(alphabet, np.uint8)
    xs, ys, xo, yo, bws = [], [], [0], [0], []
    for t in range(n):
        bw = int(rng.choice(bw_choices))
        kind = rng.random()
        xl = int(rng.integers(1, max_x + 1))
        x = al[rng.integers(0, len(al), xl)]
        if kind < 0.02:
            x = x[:0]
            y = al[rng.integers(0, len(al), int(rng.integers(0, 20)))]
        elif kind < 0.04:
            y = x[:0]
        elif kind < 0.12:
            y = al[rng.integers(0, len(al), int(rng.integers(1, xl + bw + 25)))]
        else:
            yl = []
            for b in x:
                u = rng.random()
                if u < 0.02:
                    yl.append(al[rng.integers(0, len(al))])
                elif u < 0.03:
                    continue
                elif u < 0.04:
                    yl.append(b); yl.append(al[rng.integers(0, len(al))])
                else:
                    yl.append(b)
            pad = int(rng.integers(0, bw + 21))
            y = np.concatenate([np.array(yl, np.uint8), al[rng.integers(0, len(al), pad)]]) if (yl or pad) else x[:0]
            if rng.random() < 0.1:
                y = y[: int(rng.integers(0, len(y) + 1))]
        xs.append(x); ys.append(y); xo.append(xo[-1] + len(x)); yo.append(yo[-1] + len(y)); bws.append(bw)
    cat = lambda v: np.concatenate(v).astype(np.uint8) if len(v) else np.zeros(0, np.uint8)
    bw = np.array(bws, np.uint32)
    return cat(xs), np.array(xo, np.uint64), cat(ys), np.array(yo, np.uint64), bw, bw.astype(np.int32)
