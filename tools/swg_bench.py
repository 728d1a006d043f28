"""Config 5: SWG-only microbench -- synthetic read/ref-window pairs, band widths 8..64 (x_drop = bw), GCUPS.
usage: python tools/swg_bench.py [n_pairs] [exact(0/1)]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import golden  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index  # noqa: E402


from thermite_b200.synth import swg_pairs as pairs  # noqa: E402


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    exact = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    ix = Index.create_from_memory(golden("test_ref.fasta"), golden("test_ref.gtf"))
    al = Aligner(ix, AlignOpts(min_seed_len=3))
    al.set_exact_cell_count(exact)
    for bw in (8, 16, 24, 32, 48, 61, 64):
        xs, xo, ys, yo, b, xd = pairs(20215 + bw, n, bw)
        r = al.swg_extend_batch(xs, xo, ys, yo, b, xd)
        r = al.swg_extend_batch(xs, xo, ys, yo, b, xd)
        print(f"bw {bw:3d} pairs {n} cells {r['cells']:>13d} kernel_ms {r['kernel_ms']:8.3f} GCUPS {r['cells'] / r['kernel_ms'] / 1e6:8.2f} "
              f"cells/pair {r['cells'] / n:7.0f} exact={exact}", flush=True)
