"""Per-kernel timeline of one MSS loss forward + backward from an ncu CSV (gpu__time_duration, dram bytes)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ik, im, iv, iid, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[iid], {"name": r[ik]})[r[im]] = (float(r[iv].replace(",", "")), r[iu])
L = list(per.values()); n = len(L) // 3
tot = 0
for d in L[2 * n:]:
    t = d["gpu__time_duration.sum"]; rd = d["dram__bytes_read.sum"]; wr = d["dram__bytes_write.sum"]
    us = t[0] / (1000 if t[1].startswith("n") else 1)
    sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}
    tot += us
    print(f"{us:9.1f} us  rd {rd[0] * sc[rd[1]]:7.2f} MB  wr {wr[0] * sc[wr[1]]:7.2f} MB  {d['name'][:60]}")
print(f"total {tot:.1f} us")
