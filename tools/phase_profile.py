"""Per-phase cycle breakdown of k_extend (needs the profiling build: make -C thermite_b200/csrc libthermite_gpu_prof.so;
run with THERMITE_GPU_LIB=thermite_b200/csrc/libthermite_gpu_prof.so python tools/phase_profile.py <scale> <reads>)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from thermite_b200 import AlignOpts, Aligner, Index, lib  # noqa: E402

NAMES = ["ELR stage x/y", "ELR swg fill+trace", "ELR stitch", "hit: ref+window", "exon tree next", "tx lift+seed match",
         "tx same-problem/ELR/keep-best", "decide+lift_tx_to_gx+gene find+concat", "read setup", "accept/store cand",
         "finalize+output", "-"]
scale, n = float(sys.argv[1]), int(sys.argv[2])
contigs, gtf, txs, fa = bench.make_world(scale)
ix = Index.create_from_memory(fa, gtf)
al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
bases, offs = bench.make_reads(contigs, txs, n, bench.SEEDS["reads"])
for _ in range(2):
    r = al.align_reads(bases, offs)
out = (C.c_uint64 * 16)()
lib().tg_ctx_debug_phases(al._h, out)
tot = sum(out[:11])
print("kernel ms", al.last_kernel_ms(), "total lane-0 cycles", tot, "per read", tot / n)
for k in range(11):
    print(f"{NAMES[k]:42s} {out[k] / n:10.0f} cyc/read  {100.0 * out[k] / max(tot, 1):5.1f}%")
