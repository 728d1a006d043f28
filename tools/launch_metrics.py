"""Per-kernel summary of ONE alignment pass from an ncu launch list taken with a few metrics (see profiles/README.md):
    python tools/launch_metrics.py <launches.csv> [pass index]"""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
idx = {h: i for i, h in enumerate(rows[hi])}
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(idx):
        continue
    name = re.sub(r"\(.*", "", re.sub(r"<unnamed>::|void ", "", r[idx["Kernel Name"]]))
    d = L.setdefault(r[idx["ID"]], {"name": name})
    try:
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1.0)
    except ValueError:
        pass
launches = list(L.values())
inits = [i for i, l in enumerate(launches) if l["name"] == "k_round_init"]
packs = [i for i, l in enumerate(launches) if l["name"] == "k_pack_reads"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
start = max(p for p in packs if p < inits[k])
while start - 1 in packs or (start > 0 and launches[start - 1]["name"].startswith("k_probe")):
    start -= 1
end = min([p for p in packs if p > inits[k]] + [len(launches)])
agg = collections.OrderedDict()
for l in launches[start:end]:
    a = agg.setdefault(l["name"], dict(n=0, t=0.0, inst=0.0, tinst=0.0, rd=0.0, wr=0.0, occ=0.0, iss=0.0))
    t = l.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["t"] += t
    a["inst"] += l.get("smsp__inst_executed.sum", 0.0)
    a["tinst"] += l.get("smsp__inst_executed.sum", 0.0) * l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    a["rd"] += l.get("dram__bytes_read.sum", 0.0); a["wr"] += l.get("dram__bytes_write.sum", 0.0)
    a["occ"] += t * l.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    a["iss"] += t * l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(a["t"] for a in agg.values())
print(f"launches {len(launches)} in the capture; pass {k}: launches [{start}, {end}) = {end - start}, serialised kernel time {tot:.2f} ms")
print(f"{'kernel':22s} {'n':>3s} {'ms':>8s} {'share':>6s} {'lanes':>6s} {'warp-instr':>11s} {'warps%':>7s} {'issue%':>7s} {'DRAM rd MB':>11s} {'wr MB':>8s}")
for n, a in agg.items():
    print(f"{n:22s} {a['n']:3d} {a['t']:8.3f} {100 * a['t'] / tot:5.1f}% {a['tinst'] / max(a['inst'], 1):6.1f} {a['inst'] / 1e6:10.1f}M "
          f"{a['occ'] / max(a['t'], 1e-9):6.1f}% {a['iss'] / max(a['t'], 1e-9):6.1f}% {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:8.1f}")
