"""Device-resident 4 M-read passes: CUDA-event times of the seeding stage, the extension stage and the DP section alone
(what the A/B runs of profiles/r2_dp_pair_experiment.txt compare).  usage: python tools/seed_time.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from thermite_b200 import AlignOpts, Aligner, Index
contigs, gtf, txs, fa = bench.make_world(1.0)
ix = Index.create_from_memory(fa, gtf, sa_device=0)
al = Aligner(ix, AlignOpts(20, 0.0, 30, 1, True))
bases, offs = bench.make_reads(contigs, txs, 4000000, 20213)
import torch, numpy as np
d_b = torch.from_numpy(bases).cuda(); d_o = torch.from_numpy(offs.view(np.int64)).cuda()
for i in range(4):
    r = al.align_reads_device_raw(d_b.data_ptr(), d_o.data_ptr(), len(offs)-1, int(offs[-1]), 91)
    torch.cuda.synchronize()
    print("seed, extend ms", al.last_kernel_ms(), "dp", al.last_dp_ms())
