"""Opcode histogram of one kernel's source page: ncu -i rep --page source --csv ... > file.csv; python tools/ncu_sass_hist.py file.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) > 10 and r[0].startswith("0x")]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
totI = sum(int(r[iI]) for r in data)
print("static instrs", len(data), "samples", tot, "warp-instr executed", totI)
c, ci = Counter(), Counter()
for r in data:
    t = r[iSrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[iS])
    ci[op] += int(r[iI])
for op, v in c.most_common(22):
    print(f"  {op:10s} samples {100 * v / tot:5.1f}%  instr {100 * ci[op] / totI:5.1f}%")
print("static instrs executed at least once:", sum(1 for r in data if int(r[iI]) > 0))
ex = sorted((int(r[iI]) for r in data), reverse=True)
s = 0
for i, v in enumerate(ex):
    s += v
    if s > 0.9 * totI:
        print("90% of dynamic instrs in", i, "static instrs")
        break
