"""Per decoder pass: DRAM bytes (read + write) and device time per kernel, from the CSV of the capture described in
tools/prof_step_dram.py.  usage: python tools/dram_per_step.py launches.csv passes_in_csv [skip_passes]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
n_pass, skip = int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 8
h = rows[0]
ik, im, iv, iu, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.OrderedDict()
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}
for r in rows[1:]:
    per.setdefault(r[iid], {"name": r[ik]})[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
launches = [d for d in per.values() if "golf::" in d["name"]]
per_pass = len(launches) // n_pass
launches = launches[skip * per_pass:]
agg = collections.OrderedDict()
for d in launches:
    n = d["name"].split("(")[0].replace("void ", "").replace("golf::", "")
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("dram__bytes_read.sum", 0); a[2] += d.get("dram__bytes_write.sum", 0); a[3] += d.get("gpu__time_duration.sum", 0)
np_ = n_pass - skip
print(f"{per_pass} library launches per decoder pass; averages over {np_} passes (first {skip} skipped), 8 rotating input sets")
tr = tw = tt = 0.0
for n, a in agg.items():
    print(f"{a[0] / np_:4.1f} x  read {a[1] / np_ / 1e6:7.2f} MB  write {a[2] / np_ / 1e6:7.2f} MB  {a[3] / np_:7.1f} us  {n}")
    tr += a[1] / np_; tw += a[2] / np_; tt += a[3] / np_
print(f"per pass: read {tr / 1e6:.2f} MB + write {tw / 1e6:.2f} MB = {(tr + tw) / 1e6:.2f} MB DRAM, {tt:.1f} us of kernels")
