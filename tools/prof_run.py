"""Small fixed invocation of the hot path for ncu captures (see profiles/README.md).
usage: python tools/prof_run.py <genome scale> <reads> [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index  # noqa: E402

scale, n = float(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
contigs, gtf, txs, fa = bench.make_world(scale)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
al = Aligner(ix, AlignOpts(bench.FLAGS["k"], bench.FLAGS["pct"], bench.FLAGS["min_score"], bench.FLAGS["score_range"],
                           bench.FLAGS["intron_mode"]))
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
for _ in range(steps):
    r = al.align_reads(bases, offs)
print("reads", n, "alns", len(r.alns), "kernel ms (seed, extend)", al.last_kernel_ms(), r.counters)
import ctypes as C
from thermite_b200 import lib
out = (C.c_uint64 * 16)()
lib().tg_ctx_debug_phases(al._h, out)
print("reads leaving the round path: too many hits", out[12], "table overflow", out[13], "finaliser scratch", out[14], "still active", out[15])
d = (C.c_uint64 * 52)()
lib().tg_ctx_debug_rounds(al._h, d)
print("items", d[0], "hops words", d[1], "complex reads", d[2], "fin words", d[3], "launches", al.last_kernel_launches())
print("round_active", [int(d[4 + r]) for r in range(16)])
print("round_tasks", [int(d[20 + r]) for r in range(16)])
print("round_ops  ", [int(d[36 + r]) for r in range(16)])
cl = (C.c_uint32 * (16 * 12))()
lib().tg_ctx_debug_classes(al._h, cl)
for r in range(6):
    print("round", r, "tasks per band class (0 = warp kernel, 1 -> WB 4, c -> WB 8(c-1)):", [int(cl[r * 12 + c]) for c in range(12)])
dpms = (C.c_float * 16)()
nr = lib().tg_ctx_debug_round_dp_ms(al._h, dpms)
print("rounds", nr, "DP ms per round", [round(float(dpms[r]), 3) for r in range(nr)], "total", round(al.last_dp_ms(), 3))
