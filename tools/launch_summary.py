"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and first durations."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "")
    tot.setdefault(name, []).append(float(r[-1]) / 1e6)
all_ms = sum(sum(v) for v in tot.values())
for k, v in tot.items():
    print(f"{k:28s} n={len(v):3d} total={sum(v):8.3f} ms ({100 * sum(v) / all_ms:5.1f}%)  first={', '.join(f'{x:.3f}' for x in v[:6])}")
