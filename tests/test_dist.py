"""N>1 host logic on CPU: world_size-2 gloo.  The index blob is broadcast once, reads shard contiguously, per-rank
records are merged back in read order.  (No GPU here: the per-rank compute is the CPU oracle standing in for the
device; the GPU path of the same plumbing is exercised by bench.py under torchrun.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import small_world
from oracle import orc
from thermite_b200 import Index, api, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs, gtf, txs, fa = small_world(5)
    index = Index.create_from_memory(fa, gtf) if rank == 0 else None
    index, _ = api.broadcast_index(index, rank, device=None)
    # every rank sees the same batch and takes its contiguous shard
    bases, offs = synth.make_reads(9, contigs, txs, 101, L=60, sub=0.02)
    lo, hi = api.shard_range(len(offs) - 1, rank, world)
    sb = bases[int(offs[lo]): int(offs[hi])]
    so = offs[lo: hi + 1] - offs[lo]
    oix = orc.Index.create(fa, gtf)
    r = oix.align_batch(sb, so, k=16, pct=0.0, min_score=20, intron_mode=True)
    first = r.read_off[:-1].copy()
    count = (r.read_off[1:] - r.read_off[:-1]).astype(np.uint32)
    q.put((rank, [x.name for x in index.refs()], len(index.txome().txs), index.blob().nbytes, lo, hi,
           first, count, r.alns, r.ops))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_broadcast_shard_merge():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # identical index metadata on both ranks after the broadcast
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2] and got[0][3] == got[1][3]
    assert (got[0][4], got[0][5], got[1][4], got[1][5]) == (0, 50, 50, 101)
    first, count, alns, ops = api.merge_shards([(g[6], g[7], g[8], g[9]) for g in got])
    # equals the single-process result on the whole batch
    contigs, gtf, txs, fa = small_world(5)
    bases, offs = synth.make_reads(9, contigs, txs, 101, L=60, sub=0.02)
    whole = orc.Index.create(fa, gtf).align_batch(bases, offs, k=16, pct=0.0, min_score=20, intron_mode=True)
    import ht
    assert not ht.compare_alignments(dict(first=first, count=count, alns=alns, ops=ops), whole, 101)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 3, 8):
            r = [api.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
