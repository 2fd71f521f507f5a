#!/usr/bin/env python3
"""Attribute ncu warp-stall samples (source page, SASS view) to the device sub-functions of a kernel, using the
symbol table of the cubin extracted from the object file.  usage: ncu_hot_functions.py report.ncu-rep obj.o kernel_substr"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main(rep, obj, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
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
    sym = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
    funcs = []
    for line in sym.splitlines():
        p = line.split()
        if len(p) >= 8 and p[3] == "FUNC" and kern in p[7]:
            size = int(p[2], 16) if p[2].startswith("0x") else int(p[2])
            name = p[7].split("$")[-1] if "$" in p[7] else "<kernel body>"
            funcs.append((int(p[1], 16), size, name))
    funcs.sort()
    kbase = 0  # symbol values are section offsets; the source page starts at the section start
    agg = {}
    for a, s, n, t in data:
        off = a - base + kbase
        name = "<kernel body>"
        for fa, fs, fn in funcs:
            if fn != "<kernel body>" and fa <= off < fa + fs:
                name = fn
        d = agg.setdefault(name, [0, 0, 0])
        d[0] += s
        d[1] += n
        d[2] += t
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"{'samples%':>9} {'inst%':>7} {'thr/inst':>8}  function")
    for name, (s, n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() if name.startswith("_Z") else name
        dem = re.sub(r"\(.*", "", dem)
        print(f"{100.0 * s / tot:9.2f} {100.0 * n / toti:7.2f} {t / max(n, 1):8.2f}  {dem}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
