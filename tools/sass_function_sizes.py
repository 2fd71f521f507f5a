#!/usr/bin/env python3
"""Instruction count per function (kernel + its non-inlined device functions) of an object file: sass_function_sizes.py obj.o"""
import os, re, subprocess, sys, tempfile

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(sys.argv[1])], cwd=tmp, capture_output=True)
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    cur, sizes, order = None, {}, []
    for line in dis.splitlines():
        m = re.match(r"^(\$?[_A-Za-z0-9\$]+):\s*$", line)
        if m and not m.group(1).startswith(".L"):
            cur = m.group(1); sizes[cur] = 0; order.append(cur); continue
        if cur and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
            sizes[cur] += 1
    for k in order:
        if sizes[k]:
            name = subprocess.run(["c++filt", k.split("$")[-1]], capture_output=True, text=True).stdout.strip()
            print(f"{sizes[k]:7d}  {name[:110]}")
