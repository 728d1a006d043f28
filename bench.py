#!/usr/bin/env python
"""bench.py -- reads/s of the B200 alignment hot path on the chr21 stand-in workload (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (seeding -> hit loop with banded SWG -> lifting -> filters) over one batch of
synthetic 10x-R2-shaped reads per GPU.  Reads shard across ranks with no data-path collective (weak scaling);
the only collective is the NCCL broadcast of the flat index at start-up.

Rank 0 prints ONE JSON line.  `value` = reads/s with inputs resident in HBM; `e2e` = the same through
tg_align_batch with pinned HOST buffers (H2D + kernels + D2H inside the timed region); `roofline` describes the
dominant kernel; `cpu_baseline` is the CPU oracle (a C++ restatement of thermite's CPU path -- the Rust reference
cannot be built in this image) timed on the box's host cores on a bounded sample of the same workload.
`--impl reference` times that CPU restatement as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAGS = dict(k=20, pct=0.0, min_score=30, score_range=1, intron_mode=True)  # data/Makefile:39  -k20 -s0 --intron-mode
READ_LEN = 91
SEEDS = dict(genome=20212, reads=20213)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale (1.0 = chr21-sized stand-in)")
    ap.add_argument("--reads", type=int, default=4_000_000, help="reads per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline work per step (all threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--swg-pairs", type=int, default=1 << 20, help="config 5 microbench: pairs per band width per GPU")
    ap.add_argument("--no-swg-microbench", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true", help="skip the default-flags / pageable / file-pipeline blocks")
    ap.add_argument("--file-reads", type=int, default=4_000_000, help="reads of the FASTQ -> PAF / SAM file-pipeline block")
    return ap.parse_args()


def workload_name(scale):
    base = "synth21-91bp: synthetic chr21 stand-in (46,709,983 bp, N blocks, repeat families, poly-A runs; GTF 800 genes) " \
           "vs 10x-R2-shaped reads, flags -k20 -s0 --intron-mode"
    return base if scale == 1.0 else base + f" [genome scale {scale}]"


def make_world(scale):
    from thermite_b200 import synth
    contigs, gtf, txs = synth.synth21(scale, SEEDS["genome"])
    return contigs, gtf, txs, synth.fasta_bytes(contigs)


def make_reads(contigs, txs, n, seed):
    from thermite_b200 import synth
    return synth.make_reads(seed, contigs, txs, n, L=READ_LEN, frac_tx=0.8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            load = [v for v in sm if v >= 0.5 * max(sm)] or sm
            out = dict(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, streaming copy)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def seed_algorithmic_bytes(n_reads, L, k, hits):
    # sector-granular  32*(ceil(L/128) + (L-k+1) + 2H) + 8H  per read, H = Mems handed to align_seed_hit
    return 32 * (n_reads * (-(-L // 128) + (L - k + 1)) + 2 * hits) + 8 * hits


def cpu_baseline(fa, gtf, bases, offs, cpu_seconds, threads):
    """Oracle (C++ port of thermite's CPU path) on a bounded sample of the same reads."""
    from oracle import orc
    t0 = time.time()
    oix = orc.Index.create(fa, gtf)
    build_s = time.time() - t0
    probe = 2000
    r = oix.align_batch(bases[: int(offs[probe])], offs[: probe + 1], n_threads=1, **_orc_flags())
    rate1 = probe / max(r.seconds, 1e-9)
    n = int(min(len(offs) - 1, max(probe, rate1 * cpu_seconds * max(1, threads) * 0.6)))
    r = oix.align_batch(bases[: int(offs[n])], offs[: n + 1], n_threads=threads, **_orc_flags())
    return dict(value=n / r.seconds, unit="reads/s", cores=threads, kind="port",
                sample=f"first {n} reads of the step batch, {threads} host threads over contiguous shards "
                       f"(single thread: {rate1:.0f} reads/s on {probe} reads); oracle index build {build_s:.0f}s not timed",
                single_thread_reads_per_s=rate1), (n, r)


def _flat(first, count, alns, ops):
    """Records and operations of a result in read order (vectorised)."""
    count = count.astype(np.int64)
    first = first.astype(np.int64)
    rec = np.repeat(first - np.concatenate(([0], np.cumsum(count)[:-1])), count) + np.arange(int(count.sum()))
    a = alns[rec]

    def gather(off, ln):
        ln = ln.astype(np.int64)
        start = np.concatenate(([0], np.cumsum(ln)[:-1]))
        idx = np.repeat(off.astype(np.int64) - start, ln) + np.arange(int(ln.sum()))
        return ops[idx], ln

    gx, gl = gather(a["ops_off"], a["ops_len"])
    tx, tl = gather(a["tx_ops_off"], a["tx_ops_len"])
    return a, gx, gl, tx, tl


def parity_vs_oracle(gpu_res, orc_res, n):
    """Bit-exact comparison of the GPU records of the first n reads with the oracle's (the checker, not the product)."""
    from oracle.orc import ALN_DTYPE
    ro = orc_res.read_off.astype(np.int64)
    ocount = (ro[1:] - ro[:-1])[:n]
    gcount = gpu_res.count[:n].astype(np.int64)
    bad_reads = int((ocount != gcount).sum())
    if bad_reads:
        return dict(reads=n, reads_with_different_record_count=bad_reads, identical=False)
    ga, ggx, ggl, gtx, gtl = _flat(gpu_res.first[:n], gpu_res.count[:n], gpu_res.alns, gpu_res.ops)
    oa, ogx, ogl, otx, otl = _flat(ro[:n], ocount, orc_res.alns, orc_res.ops)
    fields = [f for f in ALN_DTYPE.names if f not in ("ops_off", "tx_ops_off", "pad")]
    diff_fields = [f for f in fields if not np.array_equal(ga[f], oa[f])]
    ops_ok = np.array_equal(ggl, ogl) and np.array_equal(ggx, ogx) and np.array_equal(gtl, otl) and np.array_equal(gtx, otx)
    return dict(reads=n, records=int(len(ga)), op_words=int(len(ggx) + len(gtx)), fields_compared=len(fields),
                fields_with_differences=diff_fields, operations_identical=bool(ops_ok),
                identical=bool(not diff_fields and ops_ok))


SWG_BWS = (8, 16, 24, 32, 48, 61, 64)


def swg_microbench(aligner, n_pairs, int_roof, seed_shift=0, parity=True, threads=1):
    """BASELINE.json config 5 / SURVEY 8d: SwgExtend::extend alone on synthetic read / ref-window pairs, band widths 8..64,
    x_drop = bw, through tg_swg_extend_batch (the thread-per-extension DP kernels).  Per band width: GCUPS on the cells the
    reference's loops visit (credited) and on the cells the bound-stopped kernels really visit (executed), both as a
    fraction of the INT-pipe roofline, and -- parity -- score / xend / yend / operations of ALL pairs compared with the
    oracle's (blake2b digest of the concatenated arrays on both sides)."""
    import hashlib
    from thermite_b200 import synth
    out = []
    for bw in SWG_BWS:
        xs, xo, ys, yo, b, xd = synth.swg_pairs(20215 + bw + seed_shift, n_pairs, bw)
        aligner.set_exact_cell_count(True)
        aligner.swg_extend_batch(xs, xo, ys, yo, b, xd)
        ex = aligner.swg_extend_batch(xs, xo, ys, yo, b, xd)
        aligner.set_exact_cell_count(False)
        aligner.swg_extend_batch(xs, xo, ys, yo, b, xd)
        r = aligner.swg_extend_batch(xs, xo, ys, yo, b, xd)
        row = dict(bw=bw, pairs=n_pairs, cells_reference=int(ex["cells"]), cells_executed=int(r["cells"]),
                   dp_ms=r["dp_ms"], dp_ms_exact_mode=ex["dp_ms"], all_kernels_ms=r["kernel_ms"],
                   gcups_credited=ex["cells"] / r["dp_ms"] / 1e6, gcups_executed=r["cells"] / r["dp_ms"] / 1e6,
                   gcups_exact_mode=ex["cells"] / ex["dp_ms"] / 1e6,
                   gcups_executed_all_kernels=r["cells"] / r["kernel_ms"] / 1e6)
        row["frac_credited"] = row["gcups_credited"] / int_roof
        row["frac_executed"] = row["gcups_executed"] / int_roof
        row["frac_exact_mode"] = row["gcups_exact_mode"] / int_roof

        def digest(d):
            h = hashlib.blake2b(digest_size=16)
            for k in ("score", "xend", "yend", "ops_off", "ops"):
                h.update(np.ascontiguousarray(d[k]).tobytes())
            return h.hexdigest()
        row["digest"] = digest(r)
        same_modes = digest(ex) == row["digest"]
        if parity:
            from oracle import orc
            t0 = time.time()
            o = orc.swg_extend_batch(xs, xo, ys, yo, b, xd, n_threads=threads)
            row["oracle_digest"] = digest(o)
            row["oracle_cells"] = int(o["cells"])
            row["oracle_gcups"] = o["cells"] / (time.time() - t0) / 1e9
            row["identical"] = bool(row["oracle_digest"] == row["digest"] and same_modes and o["cells"] == ex["cells"] and
                                    all(np.array_equal(r[k], o[k]) for k in ("score", "xend", "yend", "ops_off", "ops")))
        else:
            row["identical"] = None if same_modes else False
        out.append(row)
    return out


def _orc_flags():
    return dict(k=FLAGS["k"], pct=FLAGS["pct"], min_score=FLAGS["min_score"], score_range=FLAGS["score_range"],
                intron_mode=FLAGS["intron_mode"])


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores, same workload/metric."""
    if rank != 0:
        return
    from oracle import orc
    threads = os.cpu_count() or 1
    contigs, gtf, txs, fa = make_world(args.scale)
    bases, offs = make_reads(contigs, txs, min(args.reads, 400_000), SEEDS["reads"])
    oix = orc.Index.create(fa, gtf)
    probe = 2000
    r = oix.align_batch(bases[: int(offs[probe])], offs[: probe + 1], n_threads=1, **_orc_flags())
    rate1 = probe / max(r.seconds, 1e-9)
    n = int(min(len(offs) - 1, max(probe, rate1 * args.cpu_seconds * threads * 0.6)))
    times = []
    for s in range(args.warmup + args.steps):
        r = oix.align_batch(bases[: int(offs[n])], offs[: n + 1], n_threads=threads, **_orc_flags())
        if s >= args.warmup:
            times.append(r.seconds)
    total = sum(times)
    value = n * len(times) / total
    line = dict(metric="reads/sec", value=value, unit="reads/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.scale), reads_per_step=n, read_len=READ_LEN,
                            note="CPU arm: C++ restatement of thermite's CPU algorithm (oracle), not the Rust binary "
                                 "(no cargo/rustc in the image); bounded sample per step"),
                cpu_baseline=dict(value=value, unit="reads/s", cores=threads, kind="port",
                                  sample=f"{n} reads per step, {threads} host threads (single thread {rate1:.0f} reads/s)"),
                e2e=dict(value=value, unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _RESULT_LINE.append(json.dumps(line))


def run_multi(args):
    """`--gpus N` in ONE process (no torchrun): the product's multi-GPU entry point.  `value` = all N contexts running
    device-resident shards concurrently; `e2e` = tg_multi_align_batch on one read-ordered batch of N x reads in pinned host
    memory, result merged in read order (records land in one pinned result, nothing is copied afterwards)."""
    import threading
    import torch
    from thermite_b200 import AlignOpts, Index, MultiAligner

    G = args.gpus
    if not torch.cuda.is_available() or torch.cuda.device_count() < G:
        raise SystemExit(f"bench.py --gpus {G} needs {G} CUDA devices: the product path has no CPU fallback")
    t_setup = time.time()
    contigs, gtf, txs, fa = make_world(args.scale)
    t_index = time.time()
    index = Index.create_from_memory(fa, gtf, sa_device=0)
    index_s = time.time() - t_index
    opts = AlignOpts(FLAGS["k"], FLAGS["pct"], FLAGS["min_score"], FLAGS["score_range"], FLAGS["intron_mode"])
    m = MultiAligner(index, opts, devices=list(range(G)))
    how, bcast_ms = m.replication()
    per = args.reads
    shards = [make_reads(contigs, txs, per, SEEDS["reads"] + 1000 * g) for g in range(G)]
    n = per * G
    bases = np.concatenate([b for b, _ in shards])
    offs = np.zeros(n + 1, np.uint64)
    pos = 0
    for g, (b, o) in enumerate(shards):
        offs[g * per: (g + 1) * per + 1] = o + np.uint64(pos)
        pos += len(b)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    ctxs = [m.context(g) for g in range(G)]
    d_in = []
    for g, (b, o) in enumerate(shards):
        dev = torch.device("cuda", g)
        d_in.append((torch.from_numpy(b).to(dev), torch.from_numpy(o.view(np.int64)).to(dev), int(o[-1])))
    setup_s = time.time() - t_setup

    def sync_all():
        for g in range(G):
            torch.cuda.synchronize(g)

    acc = [dict(seed=0.0, ext=0.0, dp=0.0, launches=0, ms=0.0, res=None) for _ in range(G)]

    def dev_steps(g, k, timed):
        torch.cuda.set_device(g)
        st = torch.cuda.ExternalStream(ctxs[g].stream_ptr(), device=torch.device("cuda", g))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(k):
            r = ctxs[g].align_reads_device_raw(d_in[g][0].data_ptr(), d_in[g][1].data_ptr(), per, d_in[g][2], READ_LEN)
            if timed:
                a, b = ctxs[g].last_kernel_ms()
                acc[g]["seed"] += a; acc[g]["ext"] += b; acc[g]["dp"] += ctxs[g].last_dp_ms()
                acc[g]["launches"] += ctxs[g].last_kernel_launches()
        e1.record(st)
        torch.cuda.synchronize(g)
        acc[g]["ms"] = e0.elapsed_time(e1)
        acc[g]["res"] = r

    def all_devices(k, timed):
        th = [threading.Thread(target=dev_steps, args=(g, k, timed)) for g in range(G)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    all_devices(args.warmup, False)
    m.set_exact_cell_count(True)
    all_devices(1, False)
    ref_cells = float(sum(a["res"].swg_cells for a in acc))
    m.set_exact_cell_count(False)
    all_devices(1, False)
    sync_all()
    sampler = ClockSampler(0)
    all_devices(args.steps, True)
    sync_all()
    dev_ms = max(a["ms"] for a in acc)
    cells = float(sum(a["res"].swg_cells for a in acc))
    launches = sum(a["launches"] for a in acc)

    for _ in range(max(1, args.warmup // 2)):
        hres = m.align_reads_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hres = m.align_reads_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    timing = m.last_timing()
    clocks = sampler.stop()
    h2d = int(bases.nbytes + offs.nbytes)
    d2h = int(n * 8 + hres.n_alns * 40 + hres.n_ops * 4)
    hbm_peak, peak_src, sm_max = peaks()
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    dp_ms = max(a["dp"] for a in acc) / args.steps
    int_roof = G * 148 * 4 * 16 * (sm_mhz * 1e6) / 9.0 / 1e9
    line = dict(
        metric="reads/sec", value=n * args.steps / (dev_ms / 1e3), unit="reads/s", n_gpus=G, steps=args.steps, warmup=args.warmup,
        ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
        config=dict(workload=workload_name(args.scale), reads_per_step_per_gpu=per, read_len=READ_LEN, flags="-k20 -s0 --intron-mode",
                    parallelism=f"ONE process, tg_multi_*: {G} GPUs, one host thread + context each, contiguous read shards, "
                                f"index replicated by {how} in {bcast_ms:.2f} ms",
                    l2=f"inputs larger than L2: per step per GPU {per * READ_LEN / 1e6:.0f} MB of reads against several GB of index",
                    index_bytes=int(index.blob().nbytes), index_replication=how, index_broadcast_ms=bcast_ms, setup_s=setup_s,
                    index_create_s=round(index_s, 2)),
        e2e=dict(value=n * args.steps / (e2e_ms / 1e3), unit="reads/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=e2e_ms / args.steps,
                 call="tg_multi_align_batch: one read-ordered batch in pinned host memory -> one result in read order "
                      "(40-byte records, segmented pools; every GPU's copies land at their final place)",
                 per_device_last_call=timing, segments=int(hres.n_segments)),
        gpu_launches=int(launches),
        roofline=dict(bound="int-pipe", kernel="k_round_dpt<1..11> on every GPU", achieved=ref_cells / (dp_ms / 1e3) / 1e9, peak=int_roof,
                      unit="GCUPS", frac=ref_cells / (dp_ms / 1e3) / 1e9 / int_roof, traffic=None,
                      achieved_computed_cells=cells / (dp_ms / 1e3) / 1e9, ms_per_step=dp_ms, sm_mhz=sm_mhz,
                      peak_source=f"{G} x 148 SM x 4 SMSP x 16 lanes/clk x sm_mhz / 9 ALU-pipe instructions per cell",
                      note="aggregate over the GPUs; DP time = the slowest GPU's summed DP sections per step"),
        kernel_share=dict(seeding=max(a["seed"] for a in acc) / dev_ms, extension_total=max(a["ext"] for a in acc) / dev_ms,
                          swg_dp=max(a["dp"] for a in acc) / dev_ms),
        clocks=clocks,
    )
    if not args.no_cpu_baseline:
        cb, (n_cpu, orc_res) = cpu_baseline(fa, gtf, bases, offs, args.cpu_seconds, os.cpu_count() or 1)
        line["cpu_baseline"] = cb
        # parity of the MERGED multi-GPU result, in read order, on a batch that is itself cut into N shards
        gres = m.align_reads(bases[: int(offs[n_cpu])], offs[: n_cpu + 1])
        line["parity_vs_oracle"] = dict(parity_vs_oracle(gres, orc_res, n_cpu), shards=G,
                                        what="tg_multi_align_batch over the first reads of the batch, cut into one shard per GPU")
    else:
        line["cpu_baseline"] = None
    _RESULT_LINE.append(json.dumps(line))
    m.close()


def extra_workloads(args, index, aligner, bases, offs, n, dev, stream):
    """Single-GPU side measurements next to the headline (not the `value`): config 4's DEFAULT-flags variant
    (src/main.rs:115-132: -k20 -s0.66, intron mode off), the host-buffer call with PAGEABLE inputs, and the file-to-file
    pipeline (FASTQ file -> tg_align_files -> PAF / SAM file)."""
    import tempfile as _tf
    import torch
    from thermite_b200 import AlignOpts, Aligner, OutputFormat, align_reads_from_file
    out = {}
    steps = max(2, args.steps // 2)

    # (1) pageable inputs through the same call as `e2e` (a caller that did not page-lock its buffers)
    for _ in range(2):
        aligner.align_reads_compact_raw(bases.ctypes.data, offs.ctypes.data, n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        aligner.align_reads_compact_raw(bases.ctypes.data, offs.ctypes.data, n)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    out["e2e_pageable_inputs"] = dict(value=n / (ms / 1e3), unit="reads/s", ms_per_step=ms,
                                      call="tg_align_batch_compact with ordinary (pageable) host arrays")

    # (2) default flags: fewer hits survive the 0.66 x L score floor, exonic records only
    d_bases = torch.from_numpy(bases).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    al2 = Aligner(index, AlignOpts(), device=dev.index)
    st2 = torch.cuda.ExternalStream(al2.stream_ptr(), device=dev)
    for _ in range(2):
        r2 = al2.align_reads_device_raw(d_bases.data_ptr(), d_offs.data_ptr(), n, int(offs[n]), READ_LEN)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st2)
    for _ in range(steps):
        r2 = al2.align_reads_device_raw(d_bases.data_ptr(), d_offs.data_ptr(), n, int(offs[n]), READ_LEN)
    e1.record(st2)
    torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / steps
    al2.align_reads_compact_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        h2 = al2.align_reads_compact_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)
    torch.cuda.synchronize()
    hms = (time.perf_counter() - t0) * 1e3 / steps
    out["default_flags"] = dict(flags="-k20 -s0.66 (intron mode off): src/main.rs:115-132", reads_per_step=n,
                                value=n / (dms / 1e3), unit="reads/s", ms_per_step=dms,
                                e2e=dict(value=n / (hms / 1e3), ms_per_step=hms),
                                alns_per_read=r2.n_alns / n, swg_ext_per_read=r2.swg_extensions / n)
    del al2, d_bases, d_offs

    # (3) FASTQ file -> PAF / SAM file (tg_align_files: reader, aligner, writers overlapped; files on tmpfs)
    nf = min(args.file_reads, n)
    rows = bases[: nf * READ_LEN].reshape(nf, READ_LEN)
    q = b"F" * READ_LEN
    tmp = _tf.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        qp = os.path.join(tmp, "q.fastq")
        with open(qp, "wb") as f:
            for lo in range(0, nf, 65536):
                f.write(b"".join(b"@r%d\n" % r + rows[r].tobytes() + b"\n+\n" + q + b"\n" for r in range(lo, min(nf, lo + 65536))))
        opts = AlignOpts(FLAGS["k"], FLAGS["pct"], FLAGS["min_score"], FLAGS["score_range"], FLAGS["intron_mode"])
        fp = {}
        for fmt in (OutputFormat.Paf, OutputFormat.Sam):
            best = None
            for _ in range(2):
                st = align_reads_from_file(index, [qp], os.path.join(tmp, "out." + fmt), fmt, opts, device=dev.index, batch_reads=1 << 19)
                if best is None or st["wall_ms"] < best["wall_ms"]:
                    best = st
            fp[fmt] = dict(value=nf / (best["wall_ms"] / 1e3), unit="reads/s", steady_value=best.get("steady_reads_per_s"), wall_ms=best["wall_ms"], read_ms=best["read_ms"],
                           align_ms=best["align_ms"], write_ms=best["write_ms"], format_ms=best["format_ms"],
                           bytes_out=int(best["bytes_out"]), batches=int(best["n_batches"]))
        out["file_pipeline"] = dict(reads=nf, input="plain FASTQ on tmpfs, %d MB" % (os.path.getsize(qp) >> 20), host_threads=os.cpu_count(),
                                    batch_reads=1 << 19,
                                    note="best of 2; PAF lines and SAM records are written on the GPU; a fresh context per run, so the first batches grow the device and page-locked buffers", **fp)
    finally:
        for f in os.listdir(tmp):
            os.remove(os.path.join(tmp, f))
        os.rmdir(tmp)
    return out


def main():
    args = parse_args()
    # exactly ONE line on stdout (the JSON): libraries that print there (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


_RESULT_LINE = []


def _main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        run_multi(args)
        return

    import torch
    import torch.distributed as dist
    from thermite_b200 import AlignOpts, Aligner, Index

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload + index: built on rank 0, replicated with ONE NCCL broadcast -----------------------------------
    t_setup = time.time()
    contigs, gtf, txs, fa = make_world(args.scale)
    t_index = time.time()
    if rank == 0:
        # suffix array on this GPU (csrc/tg_sa.cu); the blob is byte-identical to the host SA-IS build
        index = Index.create_from_memory(fa, gtf, sa_device=local_rank)
        blob = torch.from_numpy(index.blob())
        nbytes = torch.tensor([blob.numel()], dtype=torch.int64, device=dev)
    else:
        index, blob, nbytes = None, None, torch.zeros(1, dtype=torch.int64, device=dev)
    index_s = time.time() - t_index
    bcast_ms = 0.0
    if world > 1:
        dist.broadcast(nbytes, 0)
        d_blob = blob.to(dev) if rank == 0 else torch.empty(int(nbytes.item()), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(d_blob, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            index = Index.from_blob(d_blob.cpu().numpy())
        index.adopt_device_blob(d_blob.data_ptr(), d_blob.numel(), local_rank, keepalive=d_blob)
    opts = AlignOpts(FLAGS["k"], FLAGS["pct"], FLAGS["min_score"], FLAGS["score_range"], FLAGS["intron_mode"])
    aligner = Aligner(index, opts, device=local_rank)
    bases, offs = make_reads(contigs, txs, args.reads, SEEDS["reads"] + 1000 * rank)
    n = len(offs) - 1
    setup_s = time.time() - t_setup

    # HBM random-access yardstick for the seeding roofline, measured live on this GPU over the k-mer table itself
    gather_gbs = aligner.random_gather_gbs(1 << 26, 5) if rank == 0 else 0.0
    # ALU-pipe yardstick for the SWG roofline: dependency-free VIADDMNMX / VIMNMX3 streams, lane-operations per second
    int_peak = aligner.int_peak_lane_ops(5)

    # HBM-resident inputs for `value`; pinned host inputs for `e2e`
    d_bases = torch.from_numpy(bases).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_offs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(aligner.stream_ptr(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return aligner.align_reads_device_raw(d_bases.data_ptr(), d_offs.data_ptr(), n, int(offs[n]), READ_LEN)

    def step_host():  # the call the Rust shim makes: compact records (tg_aln_c), host buffers in and out
        return aligner.align_reads_compact_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)

    def step_host_wide():  # the same with 104-byte records (tg_align_batch)
        return aligner.align_reads_raw(h_bases.data_ptr(), h_offs.data_ptr(), n)

    # ---- value: device-resident ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        res = step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the reference's DP cell count for this batch (what GCUPS is credited on): one untimed pass that runs every column
    # the reference runs; the timed passes stop extensions early where that provably cannot change a record
    aligner.set_exact_cell_count(True)
    ref_cells = float(step_device().swg_cells)
    aligner.set_exact_cell_count(False)
    step_device()
    barrier()
    seed_ms = ext_ms = dp_ms = 0.0
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        res = step_device()
        a, b = aligner.last_kernel_ms()
        seed_ms += a
        ext_ms += b
        dp_ms += aligner.last_dp_ms()
        launches += aligner.last_kernel_launches()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    counters = dict(swg_cells=res.swg_cells, swg_extensions=res.swg_extensions, seed_hits=res.seed_hits, n_smems=res.n_smems,
                    n_alns=res.n_alns, n_ops=res.n_ops)

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        hres = step_host()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev2.record(stream)
    for _ in range(args.steps):
        hres = step_host()
    ev3.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(ev2.elapsed_time(ev3), e2e_wall_ms)  # the call is host-synchronous: wall clock bounds it from above
    clocks = sampler.stop() if sampler else None
    h2d = int(bases.nbytes + offs.nbytes)
    d2h = int(n * 8 + hres.n_alns * 40 + hres.n_ops * 4)
    # the wide-record call (104-byte tg_aln, u64 firsts) for comparison: a few steps, not the headline
    wide_steps = max(2, args.steps // 4)
    step_host_wide()
    barrier()
    t0 = time.perf_counter()
    for _ in range(wide_steps):
        wres = step_host_wide()
    barrier()
    wide_ms = (time.perf_counter() - t0) * 1e3 / wide_steps
    d2h_wide = int(n * 12 + wres.n_alns * 104 + wres.n_ops * 4)

    # ---- config 5: SWG-only microbench on every rank (parity against the oracle on rank 0 of a single-GPU run) ----------
    INSTR_PER_CELL = 9.0
    int_roof_measured = min(int_peak) / INSTR_PER_CELL / 1e9
    swg_rows = None
    if not args.no_swg_microbench:
        swg_rows = swg_microbench(aligner, args.swg_pairs, int_roof_measured, seed_shift=1000 * rank,
                                  parity=(world == 1 and not args.no_cpu_baseline), threads=os.cpu_count() or 1)
    swg_gcups = torch.tensor([[r["gcups_credited"], r["gcups_executed"], r["gcups_exact_mode"]] for r in swg_rows] if swg_rows
                             else [[0.0, 0.0, 0.0]] * len(SWG_BWS), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(swg_gcups, op=dist.ReduceOp.SUM)

    t = torch.tensor([dev_ms, e2e_ms, seed_ms, ext_ms, dp_ms, wide_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([counters["swg_cells"], counters["seed_hits"], counters["n_alns"], counters["n_ops"], counters["n_smems"],
                        counters["swg_extensions"], ref_cells, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, seed_ms, ext_ms, dp_ms, wide_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_reads = n * world * args.steps
    value = total_reads / (dev_ms / 1e3)
    e2e_value = total_reads / (e2e_ms / 1e3)
    hbm_peak, peak_src, sm_max = peaks()
    cells, hits, n_alns, n_ops, n_smems, n_ext, ref_cells, launches = [float(x) / world for x in cnt.tolist()]  # per GPU
    seed_launch_ms = seed_ms / args.steps   # pack + probe waves + select
    ext_launch_ms = ext_ms / args.steps     # all rounds: control kernels + DP
    dp_launch_ms = dp_ms / args.steps       # the banded-SWG kernels of all rounds (the per-class kernels run concurrently)
    seed_bytes = seed_algorithmic_bytes(n, READ_LEN, FLAGS["k"], hits)
    seed_gbs = seed_bytes / (seed_launch_ms / 1e3) / 1e9
    gcups_ref = ref_cells / (dp_launch_ms / 1e3) / 1e9
    gcups_computed = cells / (dp_launch_ms / 1e3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    # INT-pipe roofline: 148 SMs x 4 SMSP x 16 lanes/clk x f x P / I.  P = 1 cell per lane-op (32-bit scores),
    # I = 9 ALU-pipe instructions per cell in the shipped inner loop of k_round_dpt (cuobjdump -sass: VIADD, @P VIADD,
    # 5 x VIADDMNMX, VIMNMX3, VIMNMX; 4 more IMADs issue on the FMA pipe) -- see DESIGN.md section 4.
    int_roof_theory = 148 * 4 * 16 * (sm_mhz * 1e6) * 1.0 / INSTR_PER_CELL / 1e9
    int_roof = int_roof_measured  # measured ALU-pipe rate / 9 instructions per cell (theory kept beside it)
    traffic = {}
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except OSError:
        pass
    line = dict(
        metric="reads/sec", value=value, unit="reads/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
        data="synthetic",
        config=dict(workload=workload_name(args.scale), reads_per_step_per_gpu=n, read_len=READ_LEN,
                    flags="-k20 -s0 --intron-mode", parallelism=f"reads sharded over {world} GPU(s), index replicated",
                    l2=f"inputs larger than L2: per step {n * READ_LEN / 1e6:.0f} MB of reads against a k-mer table + suffix array + text of several GB",
                    index_bytes=int(index.blob().nbytes), kmer_table_bytes=int(aligner.kmer_table_bytes()),
                    index_broadcast_ms=bcast_ms, setup_s=setup_s, index_create_s=round(index_s, 2),
                    index_suffix_array="gpu (tg_index_host_create_from_memory_gpu)"),
        e2e=dict(value=e2e_value, unit="reads/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / args.steps,
                 call="tg_align_batch_compact: pinned host buffers in, 40-byte records (tg_aln_c) + operations + per-read first/count out",
                 wide_records=dict(value=n * world / (wide_ms / 1e3), ms_per_step=wide_ms, d2h_bytes_per_step=d2h_wide,
                                   call="tg_align_batch: 104-byte records (tg_aln)")),
        gpu_launches=int(launches),
        roofline=dict(bound="int-pipe", kernel="k_round_dpt<1..11> (banded SWG, thread per extension, one kernel per band class)", achieved=gcups_ref,
                      peak=int_roof, unit="GCUPS", frac=gcups_ref / int_roof, traffic=(traffic.get("k_round_dpt") or {}).get("bytes"),
                      traffic_note=(traffic.get("k_round_dpt") or {}).get("what"),
                      achieved_computed_cells=gcups_computed, ms_per_step=dp_launch_ms,
                      cells_per_read_reference=ref_cells / n, cells_per_read_computed=cells / n, sm_mhz=sm_mhz,
                      alu_instr_per_cell=INSTR_PER_CELL, cells_per_lane_op=1.0, frac_computed_cells=gcups_computed / int_roof,
                      int_peak_measured=dict(viaddmnmx_lane_ops_per_s=int_peak[0], vimnmx3_lane_ops_per_s=int_peak[1],
                                             how="tg_bench_int_peak: 16 independent chains per thread, 8 CTAs of 256 threads per SM, "
                                                 "best of 5, CUDA events, measured in this run"),
                      peak_theory=int_roof_theory,
                      hbm_view=dict(peak=hbm_peak, unit="GB/s",
                                    note="not HBM bound: the DP launches of the heaviest round move 0.87 GB of DRAM traffic "
                                         "in 3.3 ms (ncu, profiles/r2_ncu_dpt.csv) = 0.04 of the streaming peak; the contract's "
                                         "bound enum (hbm | tensor) has no entry for an integer max-plus recurrence"),
                      peak_source="measured ALU-pipe rate (min of the VIADDMNMX and VIMNMX3 streams, int_peak_measured) / 9 ALU-pipe "
                                  "instructions per cell; peak_theory = 148 SM x 4 SMSP x 16 lanes/clk x sm_mhz / 9 (no tensor "
                                  "cores: max-plus integer DP); MEASURED_PEAKS.json has no integer figure",
                      note="achieved = DP cells as the reference's loops visit them (credited count, one exact-count pass) / "
                           "summed CUDA-event time of the DP sections of all rounds; achieved_computed_cells counts only "
                           "the cells the early-stopped extensions really visit"),
        roofline_seed=dict(bound="hbm", kernel="k_seed_probe x3 + k_pack_reads + k_seed_select", achieved=seed_gbs, peak=gather_gbs,
                           unit="GB/s", frac=seed_gbs / gather_gbs if gather_gbs else None, traffic=(traffic.get("k_seed_probe") or {}).get("bytes"),
                           traffic_note=(traffic.get("k_seed_probe") or {}).get("what"),
                           peak_source="random 16-B loads (32-B sectors) over the 4.3 GB k-mer table, measured live in this run "
                                       "(tg_bench_random_gather, best of 5, CUDA events); streaming copy peak "
                                       f"{hbm_peak:.0f} GB/s ({peak_src})",
                           frac_of_streaming_peak=seed_gbs / hbm_peak,
                           ms_per_step=seed_launch_ms, algorithmic_bytes_per_read=seed_bytes / n,
                           note="algorithmic sector bytes (SURVEY 8d: one probe per read offset) / CUDA-event time of the "
                                "seeding stage; the shipped kernels skip most probes (E(q) is monotone), so fewer bytes move"),
        kernel_share=dict(seeding=seed_ms / dev_ms, extension_total=ext_ms / dev_ms, swg_dp=dp_ms / dev_ms,
                          extension_control=(ext_ms - dp_ms) / dev_ms),
        counters=dict(hits_per_read=hits / n, alns_per_read=n_alns / n, smems_per_read=n_smems / n, swg_ext_per_read=n_ext / n),
        clocks=clocks,
    )
    if swg_rows:
        agg = swg_gcups.tolist()
        line["swg_microbench"] = dict(
            what="config 5: SwgExtend::extend alone (tg_swg_extend_batch -> k_round_dpt<1..11>), synthetic pairs, x_drop = bw; "
                 "GCUPS = cells / CUDA-event time of the DP section (task sort + the per-class DP kernels, as in the align path; "
                 "all_kernels adds the byte -> 4-bit packing and the result collection of this entry point); credited = cells "
                 "the reference's loops visit, executed = cells the bound-stopped kernels visit; exact_mode = every reference "
                 "column computed",
            pairs_per_bw_per_gpu=args.swg_pairs, n_gpus=world, int_roofline_gcups_per_gpu=int_roof,
            rows=swg_rows,
            aggregate_gcups=[dict(bw=bw, credited=a[0], executed=a[1], exact_mode=a[2], frac_executed=a[1] / (int_roof * world))
                             for bw, a in zip(SWG_BWS, agg)],
            identical_to_oracle=(all(r["identical"] for r in swg_rows) if swg_rows[0].get("oracle_digest") else None),
            min_frac_executed=min(r["frac_executed"] for r in swg_rows))
    if world == 1 and not args.no_extra_workloads:
        line["other_workloads"] = extra_workloads(args, index, aligner, bases, offs, n, dev, stream)
    if world == 1 and not args.no_cpu_baseline:
        cb, (n_cpu, orc_res) = cpu_baseline(fa, gtf, bases, offs, args.cpu_seconds, os.cpu_count() or 1)
        line["cpu_baseline"] = cb
        # the oracle output of that sample doubles as a full-scale parity check of the records the GPU produced
        gres = aligner.align_reads(bases[: int(offs[n_cpu])], offs[: n_cpu + 1])
        line["parity_vs_oracle"] = parity_vs_oracle(gres, orc_res, n_cpu)
    else:
        line["cpu_baseline"] = None
    _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
