"""Raw `ncu --page raw --csv` export -> the columns the roofline discussion uses.
usage: python profiles/ncu_conv_table.py profiles/r02_conv_full_raw.csv [more.csv ...]"""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of active cycles"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of elapsed"),
        ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"), ("launch__registers_per_thread", "regs"),
        ("sm__cycles_active.avg", "SM active cycles (avg)"), ("sm__cycles_elapsed.max", "cycles elapsed"),
        ("smsp__inst_executed.sum", "warp instructions")]
print("| " + " | ".join(l for _, l in WANT) + " |")
print("|" + "---|" * len(WANT))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(n) if n in hdr else None for n, _ in WANT]
    for r in rows[2:]:
        cells = []
        for (n, lab), i in zip(WANT, idx):
            v = "n/a" if i is None else r[i]
            if lab == "kernel":
                v = "`" + v.split("(")[0].replace("void ", "").replace("<unnamed>::", "") + "`"
            elif i is not None and lab.endswith("MB"):
                f = float(v.replace(",", ""))
                u = units[i].lower()
                f = f * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
                v = "%.2f" % f
            elif i is not None and lab not in ("grid",):
                try:
                    v = "%.1f" % float(v.replace(",", ""))
                except ValueError:
                    pass
            cells.append(v)
        print("| " + " | ".join(cells) + " |")
