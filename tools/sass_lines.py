"""histogram of SASS instructions per source line for one kernel: python tools/sass_lines.py <obj> <kernel-substring>"""
import re, subprocess, sys, os, tempfile, collections
obj, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sect, cur, cnt = "", "?", collections.Counter()
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        sect = line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = os.path.basename(m.group(1)) + ":" + m.group(2)
    if kern in sect and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total instructions", tot, "=", tot * 16 / 1024, "KB")
for k, v in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(v, k)
