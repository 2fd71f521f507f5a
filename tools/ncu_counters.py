#!/usr/bin/env python3
"""Reads `ncu --set full` reports (gpurun_out/*.ncu-rep, read here with `ncu -i`, no GPU needed) and writes
profiles/ncu_counters.json: per codec, the per-encode sums over all launches of the capture (AMD BC7 = one launch per
mode) of DRAM bytes, executed warp / thread instructions and kernel time, plus the block count of the captured image.
bench.py scales these per-block figures to its workload for `roofline.traffic` and the `alu` object.

usage: ncu_counters.py codec=report.ncu-rep:blocks [...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def read(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def col(name, r):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    out = dict(launches=0, dram_bytes=0.0, warp_inst=0.0, thread_inst=0.0, time_s=0.0, kernels=[])
    for r in rows[2:]:
        out["launches"] += 1
        out["dram_bytes"] += col("dram__bytes_read.sum", r) + col("dram__bytes_write.sum", r)
        out["warp_inst"] += col("smsp__inst_executed.sum", r)
        out["thread_inst"] += col("thread_inst_executed", r)
        out["time_s"] += col("gpu__time_duration.sum", r)
        k = r[hdr.index("Kernel Name")]
        if k not in out["kernels"]:
            out["kernels"].append(k)
    return out


def main(args):
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    for a in args:
        codec, rest = a.split("=")
        rep, blocks = rest.rsplit(":", 1)
        c = read(rep)
        b = int(blocks)
        db[codec] = {"report": os.path.basename(rep), "blocks": b, "launches_per_encode": c["launches"], "kernels": c["kernels"],
                     "dram_bytes_per_block": c["dram_bytes"] / b, "warp_inst_per_block": c["warp_inst"] / b,
                     "thread_inst_per_block": c["thread_inst"] / b, "kernel_time_s_under_ncu": c["time_s"]}
        print(codec, db[codec])
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
