"""Summarise an `ncu --page source --csv` dump: instruction mix and stall share per opcode / per source-ish bucket.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME | python benchmarks/ncu_mix.py [units]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
units = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
hdr = None
body = []
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if hdr is None:
            hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit():
        body.append(r)
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ia]) for r in body)
tots = sum(int(r[ist]) for r in body) or 1
print("warp-instructions %d  (%.1f per unit)  sass lines %d  stall samples %d" % (tot, tot / units, len(body), tots))
c, s = Counter(), Counter()
for r in body:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ia])
    s[op] += int(r[ist])
for op, n in c.most_common(28):
    print("%-12s %14d %5.1f%%   stall %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * s[op] / tots))
