"""Summarise an `ncu --set full` report: per kernel the metrics DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/r1_full.ncu-rep profiles/r1_ncu_full_summary.txt profiles/ncu_summary.json
The JSON holds, per kernel name (template arguments stripped; the LONGEST launch of each name -- the refinement
launches of the stitch / solve kernels exit early when no sequence needs them), duration, DRAM
bytes per launch (read + write -- bench.py's roofline.traffic) and registers."""
import csv, io, json, subprocess, sys
rep, txt, js = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
M = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
     "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
     "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
M += [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in
      ("long_scoreboard", "short_scoreboard", "wait", "no_instruction", "not_selected", "math_pipe_throttle", "mio_throttle",
       "dispatch_stall", "barrier", "branch_resolving")]
M += ["sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
ik = h.index("Kernel Name")
out, summary = [], {}
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
for r in rows[2:]:
    name = r[ik]
    out.append("== " + name[:110])
    for m in M:
        if m in h:
            i = h.index(m)
            out.append(f"   {m:100s} {r[i]:>16s} {units[i]}")
    key = name.split("(")[0].split("<")[0].replace("void ", "").replace("golf::", "").strip()
    g = lambda m: (r[h.index(m)].replace(",", ""), units[h.index(m)])
    d, du = g("gpu__time_duration.sum")
    dur = float(d) * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(du, 1)
    if key not in summary or dur > summary[key]["duration_us"]:
        summary[key] = {"duration_us": float(d) * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(du, 1),
                        "dram_bytes_per_launch": to_bytes(*g("dram__bytes_read.sum")) + to_bytes(*g("dram__bytes_write.sum")),
                        "registers": int(float(g("launch__registers_per_thread")[0]))}
open(txt, "w").write("\n".join(out) + "\n")
json.dump(summary, open(js, "w"), indent=1)
print(json.dumps(summary, indent=1))
