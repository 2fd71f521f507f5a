#!/usr/bin/env python3
"""Reads ncu metric captures (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,
smsp__thread_inst_executed.sum,gpu__time_duration.sum --csv --log-file X tools/ncu_capture.py codec size`) and writes
profiles/ncu_counters.json: per codec, the per-encode sums over all launches of the capture (AMD BC7 = one launch per
mode) of DRAM bytes, executed warp / thread instructions and kernel time, plus the block count of the captured image.
bench.py scales these per-block figures to its workload for `roofline.traffic` and the `alu` object.

usage: ncu_counters.py codec=capture.csv:blocks [...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def read(path):
    """`path` is the --csv --log-file of `ncu --metrics ...` (one row per launch and metric)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID" and not r[0].startswith("==")]
    # tools/ncu_capture.py runs the encode twice (warm-up, then the one that counts): keep the second half of the launches
    all_ids = sorted({int(r[0]) for r in rows})
    keep = set(all_ids[len(all_ids) // 2:])
    rows = [r for r in rows if int(r[0]) in keep]
    out = dict(launches=0, dram_bytes=0.0, warp_inst=0.0, thread_inst=0.0, time_s=0.0, kernels=[], per_launch={})
    ids = set()
    for r in rows:
        kid, kernel, metric, unit, value = r[0], r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
        value *= UNIT.get(unit, 1.0)
        ids.add(kid)
        pl = out["per_launch"].setdefault(int(kid), dict(kernel=kernel.split("(")[0].split("::")[-1], dram_bytes=0.0, warp_inst=0.0, thread_inst=0.0, time_s=0.0))
        if metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            pl["dram_bytes"] += value
        elif metric == "smsp__inst_executed.sum":
            pl["warp_inst"] += value
        elif metric == "smsp__thread_inst_executed.sum":
            pl["thread_inst"] += value
        elif metric == "gpu__time_duration.sum":
            pl["time_s"] += value
        if kernel not in out["kernels"]:
            out["kernels"].append(kernel)
        if metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            out["dram_bytes"] += value
        elif metric == "smsp__inst_executed.sum":
            out["warp_inst"] += value
        elif metric == "smsp__thread_inst_executed.sum":
            out["thread_inst"] += value
        elif metric == "gpu__time_duration.sum":
            out["time_s"] += value
    out["launches"] = len(ids)
    out["per_launch"] = [out["per_launch"][k] for k in sorted(out["per_launch"])]
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
        if c["launches"] > 1:
            # AMD BC7: 21 launches per chunk of <= 2^19 blocks, in a fixed (mode, phase) order (bench.py AMD_LAUNCHES); the
            # records of the chunks are folded into one total per (mode, phase) over the whole encode
            pl = c["per_launch"]
            if codec == "bc7_amd" and len(pl) % 21 == 0:
                fold = []
                for k in range(21):
                    grp = pl[k::21]
                    fold.append(dict(kernel=grp[0]["kernel"], launches=len(grp), **{m: sum(g[m] for g in grp) for m in ("dram_bytes", "warp_inst", "thread_inst", "time_s")}))
                pl = fold
            db[codec]["per_launch"] = pl
        print(codec, db[codec])
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
