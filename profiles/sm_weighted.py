"""CUPTI timeline (tools/prof_timeline.py csv) -> per-kernel time weighted by the share of the 148 SMs its grid can
occupy: the graph keeps the device saturated, so this share -- not the summed duration -- is what a kernel costs.
usage: python profiles/sm_weighted.py gpurun_out/timeline.csv"""
import csv, re, sys
from collections import defaultdict
fam = defaultdict(lambda: [0, 0.0, 0.0])
tot = 0
rows = list(csv.DictReader(open(sys.argv[1])))
t0 = min(float(r['start_us']) for r in rows); t1 = max(float(r['start_us']) + float(r['dur_us']) for r in rows)
for r in rows:
    n = r['name']; n = re.sub(r"^void ", "", n); n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", n); n = re.sub(r"\(.*", "", n)[:50]
    c = int(r['ctas']); th = int(r['threads']); sm = int(r['smem']); rg = int(r['regs']); d = float(r['dur_us'])
    per = 32
    if th: per = min(per, 2048 // th)
    if sm: per = min(per, max(1, (227 * 1024) // (sm + 1024)))
    if rg and th: per = min(per, max(1, 65536 // (rg * th)))
    frac = min(1.0, c / (148.0 * per)) if c else 0.02
    fam[n][0] += 1; fam[n][1] += d; fam[n][2] += d * frac
    tot += d * frac
print("span %.3f ms, sum of durations %.3f ms, SM-weighted sum %.3f ms" % ((t1 - t0) / 1e3, sum(v[1] for v in fam.values()) / 1e3, tot / 1e3))
print("| kernel | launches | total ms | SM-weighted ms | share |"); print("|---|---:|---:|---:|---:|")
for n, (c, d, w) in sorted(fam.items(), key=lambda kv: -kv[1][2])[:40]:
    print("| `%s` | %d | %.3f | %.3f | %.1f%% |" % (n, c, d / 1e3, w / 1e3, 100 * w / tot))
