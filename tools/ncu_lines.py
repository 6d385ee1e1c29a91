"""per-source-line executed-instruction histogram of one kernel from an ncu report:
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object file> [top]
joins `ncu --page source --csv` (per-SASS-instruction counts) with `nvdisasm -g` (SASS -> source line) by order"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sect_re = sys.argv[5] if len(sys.argv) > 5 else kern  # regex for the mangled name in the object (template instances)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ie, isamp, ithr = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
sass = []
for r in rows[hdr + 1:]:
    if len(r) <= ie or not r[0]:
        break
    f = lambda x: float(x.replace(",", "")) if x else 0.0
    sass.append((r[1], f(r[ie]), f(r[isamp]), f(r[ithr])))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sect, cur, lines = "", "?", []
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        sect = line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = os.path.basename(m.group(1)) + ":" + m.group(2)
    if re.search(sect_re, sect) and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        lines.append(cur)
print("sass in report: %d, in object: %d" % (len(sass), len(lines)))
n = min(len(sass), len(lines))
cnt, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for i in range(n):
    cnt[lines[i]] += sass[i][1]
    smp[lines[i]] += sass[i][2]
    thr[lines[i]] += sass[i][3]
tot, tots = sum(cnt.values()), sum(smp.values())
print("warp instructions executed: %.4g, samples %d" % (tot, tots))
src = {}
for k, v in cnt.most_common(top):
    f, l = k.rsplit(":", 1)
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
        for cand in (os.path.join("bullet3_b200/csrc", f), p):
            if os.path.exists(cand):
                src[f] = open(cand).read().splitlines()
                break
        else:
            src[f] = []
    text = src[f][int(l) - 1].strip()[:80] if src[f] and int(l) <= len(src[f]) else ""
    print("%5.1f%% instr %5.1f%% samples  lanes %4.1f  %-22s %s" % (100 * v / tot, 100 * smp[k] / max(tots, 1), thr[k] / max(v, 1), k, text))
