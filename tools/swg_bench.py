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


def pairs(seed, n, bw, max_x=71):
    """len(x) ~ U[1,71]; y = mutated copy of x (2 % subst, 0.5 % indel) padded to len(x)+bw+20; 5 % unrelated; 1 % empty."""
    rng = np.random.default_rng(seed)
    al = np.frombuffer(b"ACGT", np.uint8)
    xl = rng.integers(1, max_x + 1, n)
    kind = rng.random(n)
    xl[kind < 0.005] = 0
    xo = np.concatenate(([0], np.cumsum(xl))).astype(np.uint64)
    xs = al[rng.integers(0, 4, int(xo[-1]))]
    yl = xl + bw + 20
    yl[(kind >= 0.005) & (kind < 0.01)] = 0
    yo = np.concatenate(([0], np.cumsum(yl))).astype(np.uint64)
    ys = al[rng.integers(0, 4, int(yo[-1]))]
    # copy x into the head of y with substitutions (indels are approximated by shifting a few copies by one)
    for t in np.nonzero((kind >= 0.06) & (xl > 0))[0]:
        x = xs[int(xo[t]): int(xo[t + 1])].copy()
        m = rng.random(len(x)) < 0.02
        x[m] = al[rng.integers(0, 4, int(m.sum()))]
        sh = 1 if rng.random() < 0.005 * len(x) else 0
        ys[int(yo[t]) + sh: int(yo[t]) + sh + len(x)] = x
    return xs, xo, ys, yo, np.full(n, bw, np.uint32), np.full(n, bw, np.int32)


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
