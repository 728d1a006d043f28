"""SURVEY 8f rows N1 / N2 end to end: FASTQ file (plain, BGZF, gzip) -> tg_align_files -> PAF / SAM / BAM file, bench workload.
usage: python tools/file_pipeline.py [reads] [batch_reads] [n_gpus]"""
import gzip
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench  # noqa: E402
from test_stream import bgzf  # noqa: E402
from thermite_b200 import AlignOpts, Index, OutputFormat, align_reads_from_file  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
opts = AlignOpts(bench.FLAGS["k"], bench.FLAGS["pct"], bench.FLAGS["min_score"], bench.FLAGS["score_range"], bench.FLAGS["intron_mode"])
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
L = bench.READ_LEN
rows = bases.reshape(n, L)
q = b"F" * L
fq = b"".join(b"@r%d\n" % r + rows[r].tobytes() + b"\n+\n" + q + b"\n" for r in range(n))
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
kinds = os.environ.get("TG_INPUTS", "plain,bgzf,gzip").split(",")
files = {k: (fq if k == "plain" else bgzf(fq) if k == "bgzf" else gzip.compress(fq, 4)) for k in kinds}
for k, v in files.items():
    open(os.path.join(tmp, "q." + k), "wb").write(v)
print(f"{n} reads, FASTQ text {len(fq) / 1e6:.0f} MB (" + ", ".join(f"{k} {len(v) / 1e6:.0f} MB" for k, v in files.items()) + "), "
      f"batches of {batch}, {gpus} GPU(s), host threads {os.cpu_count()}, files in {tmp}", flush=True)
devices = list(range(gpus)) if gpus > 1 else None
for inp in kinds:
    for fmt in os.environ.get("TG_FORMATS", "paf,sam,bam").split(","):
        if fmt == OutputFormat.Bam and inp != "plain":
            continue
        best = None
        for rep in range(2):
            out = os.path.join(tmp, "out." + fmt)
            t0 = time.perf_counter()
            st = align_reads_from_file(ix, [os.path.join(tmp, "q." + inp)], out, fmt, opts, batch_reads=batch, devices=devices)
            dt = time.perf_counter() - t0
            if best is None or st["wall_ms"] < best["wall_ms"]:
                best = dict(st, total_s=dt)
        print(f"{inp:5s} -> {fmt:3s}: {n / best['wall_ms'] / 1e3:6.2f} M reads/s whole run, {best.get('steady_reads_per_s', 0) / 1e6:6.2f} M reads/s after the first two batches (pipeline {best['wall_ms']:.0f} ms; stage busy ms: read {best['read_ms']:.0f}, "
              f"align {best['align_ms']:.0f}, write {best['write_ms']:.0f} (format {best['format_ms']:.0f}); {best['bytes_out'] / 1e6:.0f} MB out; call incl. context {best['total_s']:.2f} s)", flush=True)
for f in os.listdir(tmp):
    os.remove(os.path.join(tmp, f))
os.rmdir(tmp)
