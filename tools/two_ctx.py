"""Experiment: two contexts on two host threads sharing one GPU (the reference's embedding model: one ThermiteAligner
clone per worker thread) vs one context, device-resident inputs.  usage: python tools/two_ctx.py [reads] [steps]"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf)
opts = AlignOpts(bench.FLAGS["k"], bench.FLAGS["pct"], bench.FLAGS["min_score"], bench.FLAGS["score_range"], bench.FLAGS["intron_mode"])
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
dev = torch.device("cuda", 0)


def run(n_ctx, host):
    als = [Aligner(ix, opts, device=0) for _ in range(n_ctx)]
    per = n // n_ctx
    parts = []
    for i in range(n_ctx):
        r0, r1 = i * per, (i + 1) * per if i + 1 < n_ctx else n
        b = bases[int(offs[r0]): int(offs[r1])]
        o = (offs[r0: r1 + 1] - offs[r0]).astype(np.uint64)
        if host:
            parts.append((torch.from_numpy(b.copy()).pin_memory(), torch.from_numpy(o.view(np.int64).copy()).pin_memory(), r1 - r0, int(o[-1])))
        else:
            parts.append((torch.from_numpy(b.copy()).to(dev), torch.from_numpy(o.view(np.int64).copy()).to(dev), r1 - r0, int(o[-1])))
    torch.cuda.synchronize()

    def work(i, k):
        b, o, m, tot = parts[i]
        for _ in range(k):
            if host:
                als[i].align_reads_raw(b.data_ptr(), o.data_ptr(), m)
            else:
                als[i].align_reads_device_raw(b.data_ptr(), o.data_ptr(), m, tot, bench.READ_LEN)

    for phase, k in (("warm", 3), ("timed", steps)):
        th = [threading.Thread(target=work, args=(i, k)) for i in range(n_ctx)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"contexts {n_ctx} host_buffers {host}: {n * steps / dt / 1e6:.2f} M reads/s ({1e3 * dt / steps:.2f} ms per {n} reads)", flush=True)
    del als


for host in (False, True):
    for n_ctx in (1, 2, 3):
        run(n_ctx, host)
