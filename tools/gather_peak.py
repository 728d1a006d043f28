"""HBM random-access peak for the seeding roofline: random 16-B loads (one 32-B sector each) from a table far
larger than L2, measured with CUDA events.  Plain torch gather: this is a hardware yardstick, not product code.
usage: python tools/gather_peak.py [table_GiB] [loads_M]"""
import json
import sys

import torch


def measure(table_gib=4.0, loads_m=64, reps=5, dev="cuda:0"):
    rows = int(table_gib * (1 << 30)) // 16
    table = torch.empty((rows, 4), dtype=torch.int32, device=dev)
    table.random_(0, 1 << 30)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    idx = torch.randint(0, rows, (loads_m << 20,), device=dev, generator=g)
    out = torch.empty((loads_m << 20, 4), dtype=torch.int32, device=dev)
    torch.index_select(table, 0, idx, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.index_select(table, 0, idx, out=out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = loads_m << 20
    return dict(table_gib=table_gib, loads=n, ms=best, loads_per_s=n / (best / 1e3),
                sector_gbs=32 * n / (best / 1e3) / 1e9, payload_gbs=16 * n / (best / 1e3) / 1e9,
                how="torch.index_select of 16-B rows at uniformly random indices, best of %d, CUDA events" % reps)


if __name__ == "__main__":
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    lm = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    print(json.dumps(measure(gib, lm)))
