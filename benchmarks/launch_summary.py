"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg, cnt = collections.Counter(), collections.Counter()
for r in data:
    name = r[kn][: int(sys.argv[2]) if len(sys.argv) > 2 else 90]
    agg[name] += float(r[mv].replace(",", ""))
    cnt[name] += 1
tot = sum(agg.values())
unit = rows[hi + 1][hdr.index("Metric Unit")] if len(rows) > hi + 1 else "?"
print("%d launches, total %.3f (%s summed; cold-cache, serialised)" % (len(data), tot, unit))
for n, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print("%12.1f %5.1f%% x%-4d %s" % (v, 100 * v / tot, cnt[n], n))
