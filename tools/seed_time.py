import sys
sys.path.insert(0,'/root/repo')
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
