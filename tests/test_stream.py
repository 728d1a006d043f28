"""Streaming FASTQ ingest (csrc/host_stream.cpp): plain / gzip / BGZF input, batches cut at any size, equal to
the whole-text parser and to a plain Python restatement of needletail's FASTQ rules (src/aligner.rs:51-55).  CPU only."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from thermite_b200 import FastqReader, ThermiteError, parse_fastq


def make_fastq(seed, n, crlf=False, blank_every=0, truncated=False, no_final_newline=False):
    rng = np.random.default_rng(seed)
    out = []
    nl = b"\r\n" if crlf else b"\n"
    for i in range(n):
        L = int(rng.integers(0, 160))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), L))
        # quality lines may start with '@' or '+': the record-start test must not be fooled
        q = bytes(rng.choice(np.frombuffer(b"@+IIFF#5<", np.uint8), L))
        name = b"read%d sample=%d/%d" % (i, seed, int(rng.integers(0, 1 << 30)))
        out.append(b"@" + name + nl + seq + nl + b"+" + (name if i % 7 == 0 else b"") + nl + q + nl)
        if blank_every and i % blank_every == blank_every - 1:
            out.append(nl)
    text = b"".join(out)
    if truncated:
        text += b"@cut short\nACGT\n+\n"
    elif no_final_newline and text.endswith(b"\n"):
        text = text[:-1]
    return text


def python_parse(text):
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    lines = [l[:-1] if l.endswith(b"\r") else l for l in lines]
    recs, i = [], 0
    while i < len(lines):
        if lines[i] == b"":
            i += 1
            continue
        if i + 3 >= len(lines):
            break
        assert lines[i][:1] == b"@"
        recs.append((lines[i][1:], lines[i + 1], lines[i + 3]))
        i += 4
    return recs


def bgzf(text, block=60000):
    out = []
    for p in list(range(0, len(text), block)) + [None]:
        chunk = b"" if p is None else text[p:p + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        data = c.compress(chunk) + c.flush()
        bsize = len(data) + 25
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + data +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def read_all(path, max_reads):
    r = FastqReader(path, max_reads)
    fmt = r.format
    recs, sizes = [], []
    for bases, offs, names, name_offs, quals, qual_offs in r:
        n = len(offs) - 1
        sizes.append(n)
        b, nm, q = bases.tobytes(), names.tobytes(), quals.tobytes()
        for i in range(n):
            recs.append((nm[int(name_offs[i]):int(name_offs[i + 1])], b[int(offs[i]):int(offs[i + 1])], q[int(qual_offs[i]):int(qual_offs[i + 1])]))
    r.close()
    return fmt, recs, sizes


@pytest.mark.parametrize("variant", ["plain", "crlf_blank", "truncated", "no_final_newline"])
def test_reader_equals_python_parse_plain_gzip_bgzf(tmp_path, variant):
    kw = dict(plain={}, crlf_blank=dict(crlf=True, blank_every=5), truncated=dict(truncated=True, blank_every=11),
              no_final_newline=dict(no_final_newline=True))[variant]
    text = make_fastq(3, 4000, **kw)
    want = python_parse(text)
    assert len(want) == 4000
    # the whole-text parser agrees with the restatement
    bases, offs, names, name_offs, quals, qual_offs = parse_fastq(text)
    assert len(offs) - 1 == len(want)
    files = {"plain": text, "gzip": gzip.compress(text[: len(text) // 3]) + gzip.compress(text[len(text) // 3:]), "bgzf": bgzf(text, 7000)}
    for fmt, data in files.items():
        p = os.path.join(tmp_path, "q." + fmt)
        open(p, "wb").write(data)
        for max_reads in (1, 7, 1000, 4000, 1 << 20):
            if max_reads == 1 and fmt != "plain":
                continue
            got_fmt, recs, sizes = read_all(p, max_reads)
            assert got_fmt == fmt
            assert recs == want, (fmt, max_reads)
            assert all(0 < s <= max_reads for s in sizes)
            assert sum(sizes) == len(want)


def test_reader_large_multithreaded_segments(tmp_path):
    text = make_fastq(5, 60000, blank_every=97)  # > 4 MB: the parallel cut / count / fill path
    assert len(text) > (4 << 20)
    want = python_parse(text)
    for fmt, data in (("plain", text), ("bgzf", bgzf(text)), ("gzip", gzip.compress(text, 1))):
        p = os.path.join(tmp_path, "big." + fmt)
        open(p, "wb").write(data)
        for max_reads in (25000, 1 << 20):
            got_fmt, recs, sizes = read_all(p, max_reads)
            assert got_fmt == fmt and recs == want
            assert max(sizes) <= max_reads


def test_reader_errors(tmp_path):
    with pytest.raises(ThermiteError):
        FastqReader(os.path.join(tmp_path, "missing.fastq"))
    p = os.path.join(tmp_path, "bad.fastq")
    open(p, "wb").write(b"not a fastq\nACGT\n+\nIIII\n")
    with pytest.raises(ThermiteError):
        list(FastqReader(p))
    p = os.path.join(tmp_path, "bad.gz")
    open(p, "wb").write(gzip.compress(make_fastq(1, 50))[:-20])  # member cut short
    with pytest.raises(ThermiteError):
        list(FastqReader(p))
    p = os.path.join(tmp_path, "empty.fastq")
    open(p, "wb").write(b"")
    assert list(FastqReader(p)) == []


def test_align_files_rejects_bad_arguments_without_touching_the_gpu(tmp_path):
    """tg_align_files needs exactly one of ctx / multi and a known format; nothing is created or run otherwise."""
    import ctypes as C
    from thermite_b200 import Index, lib
    ix = Index.create_from_memory(b">a\nACGTACGTACGTACGTACGTACGT\n", b"")
    paths = (C.c_char_p * 1)(b"/nonexistent.fastq")
    out = os.path.join(tmp_path, "never.paf").encode()
    assert lib().tg_align_files(ix._h, None, None, paths, 1, out, 0, 0, None) == -1          # TG_ERR_INVALID: no context
    assert lib().tg_align_files(ix._h, C.c_void_p(1), C.c_void_p(1), paths, 1, out, 0, 0, None) == -1   # both given
    assert lib().tg_align_files(ix._h, C.c_void_p(1), None, paths, 1, out, 7, 0, None) == -1  # unknown format
    assert not os.path.exists(out.decode())


def test_reader_reads_a_pipe_through_stdin(tmp_path):
    """"-" reads stdin (the reference's CLI takes query paths only, but needletail can read any reader): plain and gzip text
    through a pipe take the streaming path (no positional reads)."""
    import subprocess
    import sys
    text = make_fastq(11, 3000, blank_every=13)
    want = python_parse(text)
    prog = ("import sys; sys.path.insert(0, %r); from thermite_b200 import FastqReader; r = FastqReader('-', 700); "
            "n = sum(len(o) - 1 for _, o, *rest in r); print(r.format if False else '', n)") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for data in (text, gzip.compress(text)):
        out = subprocess.run([sys.executable, "-c", prog], input=data, capture_output=True, timeout=120)
        assert out.returncode == 0, out.stderr.decode()[-500:]
        assert int(out.stdout.split()[-1]) == len(want)
