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


def bam_to_sam(data: bytes) -> bytes:
    """Minimal BAM reader for the tests (SAM spec v1 4.1 / 4.2): BGZF members -> header text + one SAM line per record."""
    import gzip
    import struct
    # every BGZF block is a gzip member with a BC extra field whose BSIZE is the block size - 1
    p, n_blocks = 0, 0
    while p < len(data):
        assert data[p:p + 4] == b"\x1f\x8b\x08\x04" and data[p + 12:p + 16] == b"BC\x02\x00", "not a BGZF block"
        p += struct.unpack_from("<H", data, p + 16)[0] + 1
        n_blocks += 1
    assert p == len(data)
    raw = gzip.decompress(data)
    assert raw[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", raw, 4)
    text = raw[8:8 + l_text]
    q = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, q)
    q += 4
    refs = []
    for _ in range(n_ref):
        ln, = struct.unpack_from("<i", raw, q)
        refs.append((raw[q + 4:q + 4 + ln - 1], struct.unpack_from("<i", raw, q + 4 + ln)[0]))
        q += 8 + ln
    lines = []
    while q < len(raw):
        bs, = struct.unpack_from("<i", raw, q)
        r = raw[q + 4:q + 4 + bs]
        q += 4 + bs
        ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", r, 0)
        assert (nref, npos, tlen) == (-1, -1, 0)
        o = 32
        name = r[o:o + l_name - 1]
        o += l_name
        cig = b"".join(b"%d%c" % (c >> 4, b"MIDNSHP=X"[c & 15]) for c in struct.unpack_from("<%dI" % n_cig, r, o))
        o += 4 * n_cig
        sq = bytes(b"=ACMGRSVTWYHKDBN"[(r[o + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        o += (l_seq + 1) // 2
        ql = r[o:o + l_seq]
        o += l_seq
        qual = b"*" if l_seq == 0 or ql == b"\xff" * l_seq else bytes(c + 33 for c in ql)
        tags = []
        while o < len(r):
            tag, ty = r[o:o + 2], r[o + 2:o + 3]
            o += 3
            if ty == b"A":
                tags.append(tag + b":A:" + r[o:o + 1]); o += 1
            elif ty == b"Z":
                e = r.index(b"\0", o)
                tags.append(tag + b":Z:" + r[o:e]); o = e + 1
            else:
                fmt = {b"c": "<b", b"C": "<B", b"s": "<h", b"S": "<H", b"i": "<i", b"I": "<I"}[ty]
                v, = struct.unpack_from(fmt, r, o)
                o += struct.calcsize(fmt)
                tags.append(tag + b":i:%d" % v)
        f = [name, b"%d" % flag, refs[ref_id][0] if ref_id >= 0 else b"*", b"%d" % (pos + 1), b"%d" % mapq, cig or b"*", b"*", b"0", b"0",
             sq or b"*", qual] + tags
        lines.append(b"\t".join(f))
        # the bin the spec asks for
        end = pos + sum(c >> 4 for c in struct.unpack_from("<%dI" % n_cig, r, 32 + l_name) if (c & 15) in (0, 2, 3, 7, 8))
        assert _bin == _reg2bin(pos, end if end > pos else pos + 1), (_bin, pos, end)
    return text, refs, b"".join(ln + b"\n" for ln in lines), n_blocks


def _reg2bin(beg, end):
    end -= 1
    for sh, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> sh == end >> sh:
            return base + (beg >> sh)
    return 0
