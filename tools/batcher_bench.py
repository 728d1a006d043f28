"""throughput of the per-read API (ThermiteAligner::align_read, src/wrapper.rs:72) through tg_batcher, native
worker threads (tools/batcher_bench.cpp), synth21 world.  usage: python tools/batcher_bench.py [scale=1.0] [reads=1000000]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from thermite_b200 import Index  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
csrc = os.path.join(ROOT, "thermite_b200", "csrc")
exe = "/tmp/batcher_bench"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tools", "batcher_bench.cpp"), "-o", exe,
                       "-L" + csrc, "-lthermite_gpu", "-Wl,-rpath," + csrc])
contigs, gtf, txs, fa = bench.make_world(scale)
Index.create_from_memory(fa, gtf, sa_device=0).save("/tmp/bb_index.tai")
bases, offs = bench.make_reads(contigs, txs, n, 20213)
with open("/tmp/bb_reads.bin", "wb") as f:
    f.write(np.uint64(n).tobytes())
    f.write(np.ascontiguousarray(offs, np.uint64).tobytes())
    f.write(np.ascontiguousarray(bases, np.uint8).tobytes())
print(f"host threads available: {os.cpu_count()}", flush=True)
# blocking calls keep only `threads` reads in flight (a batch per handful of reads): run them on a small sample
for threads, window, max_batch, wait, use in ((16, 1, 65536, 200, 20_000), (16, 1024, 65536, 2000, n), (16, 16384, 262144, 5000, n),
                                              (4, 65536, 262144, 5000, n), (1, 262144, 262144, 5000, n)):
    subprocess.check_call([exe, "/tmp/bb_index.tai", "/tmp/bb_reads.bin", str(threads), str(window), str(max_batch), str(wait),
                           str(use)])
