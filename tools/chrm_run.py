"""Config 2 stand-in: synthM-500k -- 500,000 simulated 91-bp reads vs the bundled GRCh38-2020-A chrM FASTA+GTF,
flags -k20 -s0 --intron-mode (data/Makefile:30).  Reports reads/s (host buffers through the C ABI and device-resident)
and checks every record against the oracle.  usage: python tools/chrm_run.py [reads]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from common import golden  # noqa: E402
from oracle import orc  # noqa: E402  (checker only)
from thermite_b200 import AlignOpts, Aligner, Index, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
fa, gtf = golden("GRCh38-2020-A-chrM.fasta"), golden("GRCh38-2020-A-chrM.gtf")
g = np.frombuffer(b"".join(fa.split(b"\n")[1:]), np.uint8)
txs = [dict(id="x", strand=f[6], exons=[(int(f[3]) - 1, int(f[4]))], chrom="chrM", gene="g")
       for f in (ln.split("\t") for ln in gtf.decode().splitlines()) if len(f) > 8 and f[2] == "exon"]
bases, offs = synth.make_reads(20211, [("chrM", g)], txs, n, L=91)
ix = Index.create_from_memory(fa, gtf)
al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
hb, ho = torch.from_numpy(bases).pin_memory(), torch.from_numpy(offs.view(np.int64)).pin_memory()
db, do = torch.from_numpy(bases).cuda(), torch.from_numpy(offs.view(np.int64)).cuda()
for name, fn in (("host buffers (e2e)", lambda: al.align_reads_raw(hb.data_ptr(), ho.data_ptr(), n)),
                 ("device-resident", lambda: al.align_reads_device_raw(db.data_ptr(), do.data_ptr(), n, int(offs[n]), 91))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chrM synthM-{n}: {name}: {n / dt / 1e6:.1f} M reads/s ({1e3 * dt:.2f} ms), kernels (seed, extend, dp) ms "
          f"{al.last_kernel_ms()} {al.last_dp_ms():.2f}; hits/read {r.seed_hits / n:.2f} cells/read {r.swg_cells / n:.0f}", flush=True)
res = al.align_reads(bases, offs)
t0 = time.time()
ores = orc.Index.create(fa, gtf).align_batch(bases, offs, k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True,
                                             n_threads=os.cpu_count() or 1)
print(f"oracle: {n / ores.seconds:.0f} reads/s on {os.cpu_count()} threads")
print("parity:", bench.parity_vs_oracle(res, ores, n))
