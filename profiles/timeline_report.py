"""Summarises tools/prof_timeline.py output: step span, per-stream busy time, concurrency histogram,
time by kernel family on the whole device, and the idle gaps.
usage: python profiles/timeline_report.py gpurun_out/timeline.csv"""
import csv, re, sys
from collections import defaultdict

rows = []
full = []
for r in csv.DictReader(open(sys.argv[1])):
    rows.append((r["name"], r["stream"], float(r["start_us"]), float(r["dur_us"])))
    if "ctas" in r:
        full.append((float(r["start_us"]), float(r["dur_us"]), int(r["ctas"]), int(r["threads"]), int(r["smem"]), int(r["regs"])))
rows.sort(key=lambda r: r[2])
t0 = rows[0][2]
t1 = max(r[2] + r[3] for r in rows)
print("events %d, span %.3f ms, sum of durations %.3f ms" % (len(rows), (t1 - t0) / 1e3, sum(r[3] for r in rows) / 1e3))
by_stream = defaultdict(float)
for n, s, a, d in rows:
    by_stream[s] += d
print("busy ms per stream:", {k: round(v / 1e3, 2) for k, v in sorted(by_stream.items(), key=lambda kv: -kv[1])})
# concurrency histogram (time-weighted number of kernels in flight)
ev = []
for n, s, a, d in rows:
    ev.append((a, 1)); ev.append((a + d, -1))
ev.sort()
hist = defaultdict(float); cur = 0; last = ev[0][0]
for t, k in ev:
    hist[cur] += t - last; last = t; cur += k
tot = sum(hist.values())
print("kernels in flight (share of span):", {k: "%.1f%%" % (100 * v / tot) for k, v in sorted(hist.items())})
fam = defaultdict(lambda: [0, 0.0])
for n, s, a, d in rows:
    n = re.sub(r"^void ", "", n); n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", n); n = re.sub(r"\(.*", "", n)
    fam[n[:60]][0] += 1; fam[n[:60]][1] += d
print("| kernel | launches | total ms | avg us |"); print("|---|---:|---:|---:|")
for n, (c, d) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:45]:
    print("| `%s` | %d | %.3f | %.1f |" % (n, c, d / 1e3, d / c))

if full:
    # crude SM-fill estimate: a kernel occupies min(1, CTAs / (148 * resident CTAs per SM)) of the device
    def share(ctas, threads, smem, regs):
        if ctas == 0:
            return 0.05
        per_sm = 32
        if threads: per_sm = min(per_sm, 2048 // max(threads, 1))
        if smem: per_sm = min(per_sm, max(1, (227 * 1024) // (smem + 1024)))
        if regs and threads: per_sm = min(per_sm, max(1, 65536 // (regs * threads)))
        return min(1.0, ctas / (148.0 * max(per_sm, 1)))
    ev = []
    for a, d, c, th, sm, rg in full:
        sh = share(c, th, sm, rg)
        ev.append((a, sh)); ev.append((a + d, -sh))
    ev.sort()
    cur = 0.0; last = ev[0][0]; busy = 0.0; hist = defaultdict(float)
    for tt, k in ev:
        f = min(cur, 1.0)
        busy += f * (tt - last); hist[min(int(f * 4), 4)] += tt - last
        last = tt; cur += k
    span = t1 - t0
    print("estimated SM fill over the span: %.1f%%; time share by fill quartile [0-25,25-50,50-75,75-100,100]: %s"
          % (100 * busy / span, ["%.0f%%" % (100 * hist[i] / span) for i in range(5)]))
