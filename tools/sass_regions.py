"""static SASS per source REGION of one kernel: instructions from inlined helpers (common.cuh, intrinsics) are charged to the most
recent line of the main source file.  python tools/sass_regions.py <obj> <mangled regex> <main source> [bucket]"""
import collections, os, re, subprocess, sys, tempfile
obj, kern, main = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 20
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sect, cur = "", 0
cnt = collections.Counter()
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        sect = line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m and os.path.basename(m.group(1)) == main:
        cur = int(m.group(2))
    if re.search(kern, sect) and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur // bucket * bucket] += 1
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", main)).read().splitlines()
for k in sorted(cnt):
    print("%5d  lines %4d-%4d  %s" % (cnt[k], k, k + bucket - 1, src[k].strip()[:90] if k < len(src) else ""))
print("total", sum(cnt.values()))
