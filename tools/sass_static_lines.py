#!/usr/bin/env python3
"""Static instruction count per source line (and per source function range) of one kernel: sass_static_lines.py obj.o kernel_substr [top]"""
import collections, os, re, subprocess, sys, tempfile

obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
counts = collections.Counter()
inside, line = False, None
for l in dis.splitlines():
    m = re.match(r"^(\$?[_A-Za-z0-9\$]+):\s*$", l)
    if m and not m.group(1).startswith(".L"):
        inside = kern in m.group(1) and "$" not in m.group(1)[1:]
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l):
        counts[line] += 1
total = sum(counts.values())
print("total", total)
for (k, v) in counts.most_common(top):
    print(f"{v:6d} {100.0 * v / total:5.1f}%  {k}")
