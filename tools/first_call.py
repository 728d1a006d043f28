"""Cost of the FIRST tg_align_batch call on a fresh context (device buffers grow there) against the following ones, 1 M reads.
usage: TG_DEBUG_TIMING=1 python tools/first_call.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from thermite_b200 import AlignOpts, Aligner, Index
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
t0=time.perf_counter()
al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
print("ctx create %.1f ms" % ((time.perf_counter()-t0)*1e3), flush=True)
bases, offs = bench.make_reads(contigs, txs, 1 << 20, 20213)
for i in range(4):
    t0=time.perf_counter()
    al.align_reads_compact_raw(bases.ctypes.data, offs.ctypes.data, len(offs)-1)
    print("call %d: %.1f ms" % (i, (time.perf_counter()-t0)*1e3), flush=True)
