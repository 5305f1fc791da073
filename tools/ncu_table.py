"""Per-kernel table (launches, time, DRAM bytes, GB/s, fraction of the measured copy peak) from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log. usage: ncu_table.py log.csv [peak_gbs]"""
import collections, csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ui, gi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6545.3
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    key = (r[0], name, r[gi])
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per.setdefault(key, {})[r[mi]] = v * scale
agg = collections.OrderedDict()
for (_id, name, grid), m in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    t = m.get("gpu__time_duration.sum", 0.0)
    b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a[0] += 1; a[1] += t; a[2] += b
    if t > a[3]:
        a[3], a[4] = t, b
print("| kernel | launches | total us | avg us | DRAM MB / launch | DRAM GB/s (all launches) | of %.0f GB/s | largest launch: us, GB/s |" % peak)
print("|---|---|---|---|---|---|---|---|")
for name, (n, t, b, tm, bm) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = b / t / 1e3 if t else 0
    print("| %s | %d | %.1f | %.1f | %.1f | %.0f | %.2f | %.1f, %.0f |" % (name, n, t, t / n, b / n / 1e6, gbs, gbs / peak, tm, bm / tm / 1e3 if tm else 0))
