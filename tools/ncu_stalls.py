"""per-instruction stall samples of one kernel (all its device functions) from an ncu report, top N with the dominant reasons:
   python tools/ncu_stalls.py <report.ncu-rep> <kernel regex> [top]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
isamp = h.index("# Samples")
reasons = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
f = lambda x: float(x.replace(",", "")) if x else 0.0
data = []
tot = 0.0
agg = {}
for k, r in enumerate(rows[hdr + 1:]):
    if len(r) <= isamp or not r[0] or r[0] == "Address":
        continue
    v = f(r[isamp])
    tot += v
    rs = sorted(((f(r[i]), n) for i, n in reasons), reverse=True)
    for val, n in rs:
        agg[n] = agg.get(n, 0.0) + val
    data.append((v, k, r[1], rs[:3]))
print("total samples %d over %d instructions" % (tot, len(data)))
print("by reason:", ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(tot, 1)) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for v, k, src, rs in sorted(data, reverse=True)[:top]:
    print("%5.1f%%  #%-5d %-58s %s" % (100 * v / tot, k, src[:58], " ".join("%s=%d" % (n[6:], val) for val, n in rs if val > 0)))
