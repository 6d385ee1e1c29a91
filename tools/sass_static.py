"""static SASS instruction count per source line / per line range of one kernel: python tools/sass_static.py <obj> <mangled-name regex> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
sect, cur = "", "?"
cnt = collections.Counter()
for line in out.splitlines():
    if line.startswith("//---") and ".text." in line:
        sect = line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = os.path.basename(m.group(1)) + ":" + m.group(2)
    if re.search(kern, sect) and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
print("total", sum(cnt.values()))
for k, v in cnt.most_common(top):
    print(v, k)
