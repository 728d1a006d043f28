"""suffix array of the index on the GPU (csrc/tg_sa.cu) against the host SA-IS, on the bench's synth21
world (93.4 M symbols of both-strand text) or a scaled one.  Checks that the two index blobs are byte-identical and
prints the times.  usage: python tools/sa_build.py [genome scale, default 1.0] [reps, default 3]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from thermite_b200 import Index, suffix_array_gpu  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
contigs, gtf, txs, fa = bench.make_world(scale)
t0 = time.perf_counter()
ix_host = Index.create_from_memory(fa, gtf)
t_host = time.perf_counter() - t0
n = ix_host.text_len()
print(f"text: {n:,} symbols (both strands + separators), genome scale {scale}", flush=True)
print(f"index creation, suffix array by SA-IS on the host (1 thread): {t_host:.2f} s", flush=True)
suffix_array_gpu(ix_host.text4()[:1024].copy(), 1024 * 16 - 64)  # CUDA context + module load, untimed
for r in range(reps):
    t0 = time.perf_counter()
    sa, ms, steps = suffix_array_gpu(ix_host.text4(), n)
    wall = time.perf_counter() - t0
    same = bool(np.array_equal(sa, ix_host.suffix_array()))
    print(f"tg_suffix_array_gpu run {r}: device {ms:.1f} ms in {steps} sort steps = {n / ms / 1e3:.1f} M suffixes/s; "
          f"wall incl. allocations, H2D text and D2H array {1e3 * wall:.1f} ms; identical to the host array: {same}", flush=True)
    assert same
t0 = time.perf_counter()
ix_gpu = Index.create_from_memory(fa, gtf, sa_device=0)
t_gpu = time.perf_counter() - t0
same = bool(np.array_equal(ix_gpu.blob(), ix_host.blob()))
print(f"index creation, suffix array on the GPU: {t_gpu:.2f} s (host {t_host:.2f} s: {t_host / t_gpu:.1f}x); "
      f"blob ({ix_gpu.blob().nbytes / 1e6:.0f} MB) byte-identical: {same}", flush=True)
assert same
