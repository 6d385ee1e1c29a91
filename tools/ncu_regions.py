"""executed warp instructions per source REGION (inlined helpers charged to the most recent line of the main file):
   python tools/ncu_regions.py <report> <kernel regex> <obj> <mangled regex> <main source> [bucket]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern, obj, sect_re, main = sys.argv[1:6]
bucket = int(sys.argv[6]) if len(sys.argv) > 6 else 20
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern] + (["--launch-skip", os.environ["NCU_SKIP"], "--launch-count", "1"] if os.environ.get("NCU_SKIP") else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
want = os.environ.get("NCU_KERNEL_SUBSTR", "")  # e.g. "(bool)1" to pick one template instance
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]) if want else 0
hdr = next(i for i, r in enumerate(rows) if i >= start and r and r[0] == "Address")
end = next((i for i, r in enumerate(rows) if i > hdr and r and r[0] == "Kernel Name"), len(rows))
rows = rows[:end]
h = rows[hdr]
ie, ithr = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
f = lambda x: float(x.replace(",", "")) if x else 0.0
sass = [(f(r[ie]), f(r[ithr])) for r in rows[hdr + 1:] if len(r) > ie and r[0]]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [x for x in os.listdir(d) if x.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sect, cur, lines = "", 0, []
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        sect = line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m and os.path.basename(m.group(1)) == main:
        cur = int(m.group(2))
    if re.search(sect_re, sect) and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
cnt, thr = collections.Counter(), collections.Counter()
for l, (e, t) in zip(lines, sass):
    cnt[l // bucket * bucket] += e
    thr[l // bucket * bucket] += t
tot = sum(cnt.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", main)).read().splitlines()
for k in sorted(cnt):
    if cnt[k] > 0.003 * tot:
        print("%5.1f%%  lanes %4.1f  lines %4d-%4d  %s" % (100 * cnt[k] / tot, thr[k] / max(cnt[k], 1), k, k + bucket - 1, src[k].strip()[:80] if k < len(src) else ""))
print("total %.4g" % tot)
