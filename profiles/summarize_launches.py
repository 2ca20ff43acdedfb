"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (n[:90], c, ns / 1e6, 100 * ns / tot, ns / c / 1e3))
print("\ntotal: %d launches, %.3f ms (cold-cache, serialised under ncu: compare shares, not absolutes)" % (len(rows), tot / 1e6))
