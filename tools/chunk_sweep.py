"""e2e (host buffers through tg_align_batch) vs input chunk size.  usage: python tools/chunk_sweep.py [reads]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
al = Aligner(ix, AlignOpts(bench.FLAGS["k"], bench.FLAGS["pct"], bench.FLAGS["min_score"], bench.FLAGS["score_range"], bench.FLAGS["intron_mode"]))
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
hb = torch.from_numpy(bases).pin_memory()
ho = torch.from_numpy(offs.view(np.int64)).pin_memory()
for chunk in (131072, 262144, 400000, 524288, 1 << 20):
    al.set_chunk_reads(chunk)
    for _ in range(3):
        al.align_reads_compact_raw(hb.data_ptr(), ho.data_ptr(), n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        al.align_reads_compact_raw(hb.data_ptr(), ho.data_ptr(), n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk:8d}: {1e3 * dt:7.2f} ms per {n} reads  ({n / dt / 1e6:.1f} M reads/s), kernels (seed, extend) ms {al.last_kernel_ms()}", flush=True)
