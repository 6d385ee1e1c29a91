"""summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time per step.  args: csv steps"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    k = r[ki].split("(")[0][-48:]
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in agg.items():
    print("%-50s %9.1f us  x%d  %5.1f%%" % (k, v[0] / steps, v[1] // steps, 100 * v[0] / tot))
print("total per step: %.1f us" % (tot / steps))
