"""Latency of tg_align_batch for small batches (what a blocking per-read caller sees through tg_batcher): round pipeline
against the single-warp path.  usage: python tools/small_batch_latency.py [scale=0.25]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
contigs, gtf, txs, fa = bench.make_world(scale)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
bases, offs = bench.make_reads(contigs, txs, 8192, 20213)
for n in (1, 16, 64, 256, 1024, 4096, 8192):
    b, o = bases[: int(offs[n])], offs[: n + 1]
    row = []
    for rounds in (True, False):
        al.set_round_pipeline(rounds)
        for _ in range(5):
            al.align_reads_raw(b.ctypes.data, o.ctypes.data, n)
        t0 = time.perf_counter()
        reps = 30
        for _ in range(reps):
            al.align_reads_raw(b.ctypes.data, o.ctypes.data, n)
        row.append((time.perf_counter() - t0) / reps * 1e6)
    print(f"n={n:5d}: round pipeline {row[0]:8.0f} us, single-warp path {row[1]:8.0f} us; launches {al.last_kernel_launches()}", flush=True)
