"""Summarise an `ncu --page source --csv` export (SASS view): cumulative instructions executed and
stall samples in windows of N SASS instructions, to see which part of a kernel costs what.
usage: python tools/ncu_regions.py src.csv [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > max(isrc, ismp, iex) and r[0].startswith("0x")]
tot_i = sum(float(r[iex]) for r in body); tot_s = sum(float(r[ismp]) for r in body)
print(f"{len(body)} SASS instructions, {tot_i/1e6:.2f} M warp-instructions executed, {tot_s:.0f} stall samples")
for a in range(0, len(body), win):
    blk = body[a:a + win]
    ni = sum(float(r[iex]) for r in blk); ns = sum(float(r[ismp]) for r in blk)
    ops = {}
    for r in blk:
        op = r[isrc].split()[0 if not r[isrc].strip().startswith("@") else 1].split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = " ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print(f"[{a:5d}..{a+len(blk):5d})  inst {100*ni/tot_i:5.1f}%  samples {100*ns/tot_s:5.1f}%   {top}")
