"""FASTQ text -> reads (tg_parse_fastq, threaded) -> records (tg_align_batch) -> PAF / SAM text
(tg_format_batch, threaded), all in memory, bench workload.  usage: python tools/fastq_to_paf.py [reads]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index, parse_fastq  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
al = Aligner(ix, AlignOpts(bench.FLAGS["k"], bench.FLAGS["pct"], bench.FLAGS["min_score"], bench.FLAGS["score_range"], bench.FLAGS["intron_mode"]))
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
L = bench.READ_LEN
rows = bases.reshape(n, L)
q = b"F" * L
fq = b"".join(b"@r%d\n" % r + rows[r].tobytes() + b"\n+\n" + q + b"\n" for r in range(n))
print(f"FASTQ text {len(fq) / 1e6:.0f} MB, {n} reads, host threads {os.cpu_count()}", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    b, o, nm, no, ql, qo = parse_fastq(fq)
    t1 = time.perf_counter()
    res = al.align_reads_raw(b.ctypes.data, o.ctypes.data, len(o) - 1)
    t2 = time.perf_counter()
    paf = al.format_result_raw(res, b, o, nm, no, ql, qo, sam=False)
    t3 = time.perf_counter()
    sam = al.format_result_raw(res, b, o, nm, no, ql, qo, sam=True)
    t4 = time.perf_counter()
    bam = al.format_result_bam_raw(res, b, o, nm, no, ql, qo, eof=True)
    t5 = time.perf_counter()
    print(f"rep {rep}: BAM {1e3 * (t5 - t4):.1f} ms ({len(bam) / 1e6:.0f} MB BGZF; SAM text -> records -> deflate on all host threads)", flush=True)
    print(f"rep {rep}: parse {1e3 * (t1 - t0):.1f} ms (incl. numpy copies), align {1e3 * (t2 - t1):.1f} ms, PAF {1e3 * (t3 - t2):.1f} ms "
          f"({len(paf) / 1e6:.0f} MB), SAM {1e3 * (t4 - t3):.1f} ms ({len(sam) / 1e6:.0f} MB); FASTQ->PAF {n / (t3 - t0) / 1e6:.2f} M reads/s", flush=True)
assert np.array_equal(b, bases)
