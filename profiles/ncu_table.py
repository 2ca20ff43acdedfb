"""Condenses an `ncu --page raw --csv` export into one markdown row per profiled launch.
usage: python profiles/ncu_table.py gpurun_out/xxx_raw.csv > profiles/rNN_xxx.md"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]


def find(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            return i
    return None


idx = [(find(n), lab) for n, lab in COLS]
print("| " + " | ".join(lab + ("" if i is None or not units[i] or lab in ("kernel", "grid", "block") else " [" + units[i] + "]") for i, lab in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    cells = []
    for i, lab in idx:
        v = "n/a" if i is None else r[i]
        if lab == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48] + "`"
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
