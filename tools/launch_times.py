"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list: per kernel name
count, mean duration, instructions, IPC."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[iid], {"name": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    n = d["name"].split("(")[0][-70:]
    a = agg.setdefault(n, {"n": 0, "ns": 0.0, "inst": 0.0, "ipc": 0.0})
    a["n"] += 1; a["ns"] += d.get("gpu__time_duration.sum", 0); a["inst"] += d.get("smsp__inst_executed.sum", 0)
    a["ipc"] += d.get("sm__inst_executed.avg.per_cycle_elapsed", 0)
tot = sum(a["ns"] for a in agg.values())
for n, a in agg.items():
    print(f"{a['ns']/a['n']/1e3:9.1f} us x{a['n']:3d}  {100*a['ns']/tot:5.1f}%  inst/launch {a['inst']/a['n']/1e6:7.2f}M  ipc/SM {a['ipc']/a['n']:.2f}  {n}")
print(f"total {tot/1e3:.1f} us")
