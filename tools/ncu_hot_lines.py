#!/usr/bin/env python3
"""Attribute ncu warp-stall samples / executed instructions (source page, SASS view) to SOURCE LINES through the line
table nvdisasm prints for the cubin of the object file.  usage: ncu_hot_lines.py report.ncu-rep obj.o kernel_substr [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main(rep, obj, kern, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, isamp, iinst, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((int(r[ia], 16), int(r[isamp] or 0), int(r[iinst] or 0), int(r[ithr] or 0)))
        except (ValueError, IndexError):
            pass
    base = min(a for a, *_ in data)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # walk the listing: section of the kernel, "//## File "x", line N" markers, /*addr*/ instructions
    line_of = {}
    in_k = False
    cur = ("?", 0)
    for l in dis.splitlines():
        if l.startswith("//--------------------- .text."):
            in_k = kern in l
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for a, s, n, t in data:
        k = line_of.get(a - base, ("?", 0))
        agg[k][0] += s
        agg[k][1] += n
        agg[k][2] += t
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"{'samples%':>8} {'inst%':>7} {'thr/inst':>8}  line")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / ts:8.2f} {100 * v[1] / ti:7.2f} {v[2] / max(v[1], 1):8.2f}  {k[0]}:{k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
