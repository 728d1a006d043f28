"""Synthetic stand-ins for the inputs BASELINE.json names but the reference tree does not ship
(SURVEY.md section 8d): a chr21-sized genome with N blocks / repeat families / poly-A runs, a GTF with
multi-exon genes on both strands, and 10x-R2-shaped reads (~91 bp) with substitutions, indels, poly-A
tails and TSO-like leaders.  Deterministic for a given seed and numpy version.  Pure data generation:
used by tests/ and bench.py, never by the alignment path itself.
"""
import numpy as np

_COMP = np.zeros(256, np.uint8)
_COMP[:] = np.arange(256)
for a, b in (("A", "T"), ("C", "G"), ("G", "C"), ("T", "A"), ("N", "N")):
    _COMP[ord(a)] = ord(b)
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def random_bases(rng, n, gc=0.41):
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return _ACGT[rng.choice(4, size=n, p=p)]


def mutate_copy(rng, seq, div):
    out = seq.copy()
    m = rng.random(len(seq)) < div
    k = int(m.sum())
    if k:
        out[m] = _ACGT[(np.searchsorted(_ACGT, out[m]) + rng.integers(1, 4, k)) % 4]
    return out


def make_genome(seed, length, lead_n=0, n_blocks=0, n_block_len=50_000, families=(), polya_runs=0,
                polya_len=(20, 45), gc=0.41):
    """One contig.  families: iterable of (copies, element_len, div_lo, div_hi)."""
    rng = np.random.default_rng(seed)
    g = random_bases(rng, length, gc)
    for copies, elen, dlo, dhi in families:
        elem = random_bases(rng, elen, gc)
        starts = rng.integers(lead_n, length - elen, copies)
        for s in starts:
            c = mutate_copy(rng, elem, rng.uniform(dlo, dhi))
            if rng.random() < 0.5:
                c = revcomp(c)
            g[s:s + elen] = c
    for _ in range(polya_runs):
        n = int(rng.integers(polya_len[0], polya_len[1] + 1))
        s = int(rng.integers(lead_n, length - n))
        g[s:s + n] = ord("A") if rng.random() < 0.5 else ord("T")
    if lead_n:
        g[:lead_n] = ord("N")
    for _ in range(n_blocks):
        s = int(rng.integers(lead_n, max(lead_n + 1, length - n_block_len)))
        g[s:s + n_block_len] = ord("N")
    return g


def make_annotation(seed, chrom, genome, n_genes, tx_per_gene=(1, 6), exons_per_tx=(2, 20), exon_len=(50, 400),
                    intron_len=(100, 50_000), lead=0, prefix=""):
    """-> (gtf_text bytes, transcripts) with transcripts = list of dict(id, strand, exons=[(s,e) 0-based], gene).
    Genes alternate strand, do not contain N, and are laid out left to right without overlap."""
    rng = np.random.default_rng(seed)
    length = len(genome)
    lines, txs = [], []
    is_n = genome == ord("N")
    n_prefix = np.concatenate(([0], np.cumsum(is_n)))
    pos = lead
    span_budget = max(1, (length - lead) // max(1, n_genes))
    for g in range(n_genes):
        strand = "+" if g % 2 == 0 else "-"
        gid = f"{prefix}G{g:05d}"
        # master exon chain for the gene
        n_ex_max = int(rng.integers(exons_per_tx[0], exons_per_tx[1] + 1))
        placed = None
        for _attempt in range(20):
            start = pos + int(rng.integers(0, max(1, span_budget // 4)))
            chain, p = [], start
            for e in range(n_ex_max):
                el = int(rng.integers(exon_len[0], exon_len[1] + 1))
                chain.append((p, p + el))
                lo, hi = np.log(intron_len[0]), np.log(intron_len[1])
                p = p + el + int(np.exp(rng.uniform(lo, hi)))
                if p - start > span_budget * 3 // 4:
                    break
            end = chain[-1][1]
            if end >= length:
                break
            if n_prefix[end] - n_prefix[start] == 0:
                placed = chain
                break
        if placed is None:
            pos += span_budget
            continue
        chain = placed
        gs, ge = chain[0][0], chain[-1][1]
        lines.append(f'{chrom}\tsynth\tgene\t{gs + 1}\t{ge}\t.\t{strand}\t.\tgene_id "{gid}"; gene_name "SYN{g}";')
        n_tx = int(rng.integers(tx_per_gene[0], tx_per_gene[1] + 1))
        for t in range(n_tx):
            if t == 0 or len(chain) <= 2:
                ex = list(chain)
            else:  # skip a random subset of internal exons
                keep = rng.random(len(chain)) < 0.7
                keep[0] = keep[-1] = True
                ex = [c for c, k in zip(chain, keep) if k]
            tid = f"{gid}.T{t}"
            lines.append(f'{chrom}\tsynth\ttranscript\t{ex[0][0] + 1}\t{ex[-1][1]}\t.\t{strand}\t.\t'
                         f'gene_id "{gid}"; transcript_id "{tid}"; gene_name "SYN{g}";')
            for (s, e) in ex:
                lines.append(f'{chrom}\tsynth\texon\t{s + 1}\t{e}\t.\t{strand}\t.\t'
                             f'gene_id "{gid}"; transcript_id "{tid}"; gene_name "SYN{g}";')
            txs.append(dict(id=tid, strand=strand, exons=ex, gene=gid, chrom=chrom))
        pos = max(pos + span_budget, ge + 1000) if ge + 1000 > pos + span_budget else pos + span_budget
    return ("\n".join(lines) + "\n").encode(), txs


def fasta_bytes(contigs):
    """contigs: list of (name, uint8 array)"""
    parts = []
    for name, seq in contigs:
        parts.append(b">" + name.encode() + b"\n")
        parts.append(seq.tobytes())
        parts.append(b"\n")
    return b"".join(parts)


def transcript_seq(genome, tx):
    s = np.concatenate([genome[a:b] for a, b in tx["exons"]])
    return s if tx["strand"] == "+" else revcomp(s)


def make_reads(seed, contigs, txs, n, L=91, frac_tx=0.7, sub=0.005, ins=0.0005, dele=0.0005, polya_frac=0.10,
               polya_len=(10, 40), tso_frac=0.02, tso_len=(8, 20)):
    """-> (bases uint8[n*L], offs uint64[n+1]).  Reads never contain N (windows with N are re-drawn)."""
    rng = np.random.default_rng(seed)
    genomes = {name: seq for name, seq in contigs}
    # source pool: transcripts (sense) and both genome strands
    tx_seqs = [transcript_seq(genomes[t["chrom"]], t) for t in txs]
    tx_seqs = [s for s in tx_seqs if len(s) >= L + 8]
    pad = L + 8
    n_tx = int(n * frac_tx) if tx_seqs else 0
    out = np.empty((n, L), np.uint8)
    win = np.empty((n, pad), np.uint8)
    if n_tx:
        lens = np.array([len(s) for s in tx_seqs])
        cat = np.concatenate(tx_seqs)
        starts = np.concatenate(([0], np.cumsum(lens)))[:-1]
        pick = rng.integers(0, len(tx_seqs), n_tx)
        offs_in = (rng.random(n_tx) * (lens[pick] - pad + 1)).astype(np.int64)
        idx = (starts[pick] + offs_in)[:, None] + np.arange(pad)[None, :]
        win[:n_tx] = cat[idx]
    names = list(genomes)
    glens = np.array([len(genomes[k]) for k in names])
    remaining = np.arange(n_tx, n)
    for _round in range(50):
        if len(remaining) == 0:
            break
        m = len(remaining)
        which = rng.choice(len(names), m, p=glens / glens.sum())
        for ci, name in enumerate(names):
            sel = np.nonzero(which == ci)[0]
            if len(sel) == 0:
                continue
            g = genomes[name]
            if len(g) < pad:
                continue
            st = rng.integers(0, len(g) - pad + 1, len(sel))
            w = g[st[:, None] + np.arange(pad)[None, :]]
            rc = rng.random(len(sel)) < 0.5
            w[rc] = _COMP[w[rc][:, ::-1]]
            win[remaining[sel]] = w
        bad = (win[remaining] == ord("N")).any(axis=1)
        remaining = remaining[bad]
    if len(remaining):
        win[remaining] = random_bases(rng, len(remaining) * pad).reshape(-1, pad)
    # substitutions (vectorised)
    m = rng.random((n, pad)) < sub
    k = int(m.sum())
    if k:
        win[m] = _ACGT[(np.searchsorted(_ACGT, win[m]) + rng.integers(1, 4, k)) % 4]
    out[:] = win[:, :L]
    # indels (per read; rare)
    n_ins = rng.binomial(L, ins, n)
    n_del = rng.binomial(L, dele, n)
    for r in np.nonzero((n_ins + n_del) > 0)[0]:
        seq = list(win[r])
        for _ in range(n_del[r]):
            p = int(rng.integers(1, L - 1))
            del seq[p]
        for _ in range(n_ins[r]):
            p = int(rng.integers(1, L - 1))
            seq.insert(p, int(_ACGT[rng.integers(0, 4)]))
        out[r] = np.array(seq[:L], np.uint8)
    # poly-A tails and TSO-like leaders
    pa = np.nonzero(rng.random(n) < polya_frac)[0]
    for r, t in zip(pa, rng.integers(polya_len[0], polya_len[1] + 1, len(pa))):
        out[r, L - int(t):] = ord("A")
    ts = np.nonzero(rng.random(n) < tso_frac)[0]
    for r, t in zip(ts, rng.integers(tso_len[0], tso_len[1] + 1, len(ts))):
        out[r, :int(t)] = random_bases(rng, int(t))
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    return out.reshape(-1), offs


# ---- named stand-in configurations ------------------------------------------------------------
CHR21_LEN = 46_709_983  # data/GRCh38-2020-A-chr21.fasta.fai:1


def synth21(scale=1.0, seed=20212):
    """chr21 stand-in.  scale < 1 shrinks the contig (tests); scale = 1 is the bench configuration."""
    length = int(CHR21_LEN * scale)
    lead = int(5_010_000 * scale)
    fam = ((int(30_000 * scale), 300, 0.05, 0.15), (max(1, int(1_000 * scale)), 6000, 0.02, 0.10))
    g = make_genome(seed, length, lead_n=lead, n_blocks=3, n_block_len=max(100, int(50_000 * scale)), families=fam,
                    polya_runs=max(1, int(300 * scale)))
    gtf, txs = make_annotation(seed + 1, "chr21", g, n_genes=max(2, int(800 * scale)), lead=lead)
    return [("chr21", g)], gtf, txs


def swg_pairs(seed, n, bw, max_x=71):
    """SURVEY 8d config 5 (SWG-only microbench): len(x) ~ U[1,71]; y = mutated copy of x (2 % subst, 0.5 % indel: a few
    copies start one symbol late) padded with random bases to len(x)+bw+20; 5 % unrelated pairs; 0.5 % empty x, 0.5 % empty y.
    Vectorised (2^20 pairs in about a second).  Returns xs, xoff, ys, yoff, band_width, x_drop (= band_width)."""
    rng = np.random.default_rng(seed)
    al = np.frombuffer(b"ACGT", np.uint8)
    xl = rng.integers(1, max_x + 1, n)
    kind = rng.random(n)
    xl[kind < 0.005] = 0
    xo = np.concatenate(([0], np.cumsum(xl))).astype(np.uint64)
    xs = al[rng.integers(0, 4, int(xo[-1]), dtype=np.uint8)]
    yl = xl + bw + 20
    yl[(kind >= 0.005) & (kind < 0.01)] = 0
    yo = np.concatenate(([0], np.cumsum(yl))).astype(np.uint64)
    ys = al[rng.integers(0, 4, int(yo[-1]), dtype=np.uint8)]
    rel = (kind >= 0.06) & (xl > 0)
    t_of = np.repeat(np.arange(n), xl)                      # task of every x symbol
    keep = rel[t_of]
    pos = np.arange(len(xs)) - xo[:-1].astype(np.int64)[t_of]  # position inside its x
    xm = xs.copy()
    sub = rng.random(len(xs), dtype=np.float32) < 0.02
    xm[sub] = al[rng.integers(0, 4, int(sub.sum()))]
    shift = (rng.random(n) < 0.005 * xl).astype(np.int64)
    dst = yo[:-1].astype(np.int64)[t_of] + shift[t_of] + pos
    ys[dst[keep]] = xm[keep]
    return xs, xo, ys, yo, np.full(n, bw, np.uint32), np.full(n, bw, np.int32)
