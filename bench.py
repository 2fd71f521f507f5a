#!/usr/bin/env python3
"""bench.py -- BCn block-compression throughput on B200 (BASELINE.json metric: BC7 Mpix/s at 8192^2 RGBA8).

A "step" = one encode of one 8192x8192 RGBA8 synthetic image (left half translucent, right half opaque --
BASELINE config[2]) per rank.  Ranks are independent (block/texture sharding, no data-path collective), so the
N-GPU run encodes N images ("weak" scaling) and `value` = N * Mpix per step / max-over-ranks step time.

  value        device-resident: input already in HBM, CUDA events on the launching stream
  e2e          same metric through the host-buffer C-ABI (b200ic_encode_host, what Image_Compress* calls):
               pinned host texels in, host blocks out, H2D + D2H inside the timed region
  roofline     dominant kernel: algorithmic bytes (5 B/px for RGBA8 -> 16-byte blocks) / event time vs measured HBM
               peak.  The search is ALU bound by design, so frac is small; `alu` gives context.
  cpu_baseline the unmodified reference (oracle/_ref/libref_oracle.so) on this host's cores, bounded sample
  --impl reference   times only that CPU reference on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CODECS = {"bc1": 1, "bc4": 4, "bc5": 5, "bc6h": 6, "bc7_amd": 7, "bc7_rg": 8}
REF_CODEC = {1: 1, 4: 4, 5: 5, 6: 6, 7: 7, 8: 8}  # oracle.ref numbering is identical


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def _counters(cname):
    """Per-block counters of the codec's kernel(s) from one `ncu --set full` capture (tools/ncu_counters.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as f:
            return json.load(f).get(cname)
    except Exception:
        return None


def _alu(cname, blocks_per_s, sm_mhz):
    """Issue side of the roofline: executed thread-instructions (lane-ops) per second against 148 SMs x 128 lanes x clock.
    The kernels are built --fmad=false, so one lane-op is at most one flop."""
    c = _counters(cname)
    if not c:
        return None
    peak = 148 * 128 * float(sm_mhz or 1965.0) * 1e6
    ach = c["thread_inst_per_block"] * blocks_per_s
    return {"achieved": ach / 1e12, "peak": peak / 1e12, "unit": "T lane-ops/s", "frac": ach / peak,
            "thread_inst_per_block": c["thread_inst_per_block"], "warp_inst_per_block": c["warp_inst_per_block"],
            "lanes_per_inst": c["thread_inst_per_block"] / max(c["warp_inst_per_block"], 1.0), "source": "profiles/ncu_counters/" + c["report"]}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


_FULL = {}


def _full(kind, size, seed):
    from gfx_imagecompress_b200 import synth
    key = (kind, size, seed)
    if key not in _FULL:
        _FULL.clear()
        _FULL[key] = synth.height_rg8(size, size, seed) if kind == "h" else synth.hdr_rgba16f(size, size, seed)
    return _FULL[key]


def workload(codec: int, size: int):
    from gfx_imagecompress_b200 import synth
    if codec in (4, 5):
        return dict(name=f"height_rg8 {size}x{size} RG8", fmt=synth.FMT_RG8, bpt=2,
                    gen=lambda seed, y0=0, rows=None: _full("h", size, seed)[y0:(y0 + rows) if rows else None])
    if codec == 6:
        return dict(name=f"hdr_rgba16f {size}x{size} RGBA16F(unsigned)", fmt=synth.FMT_RGBA16UF, bpt=8,
                    gen=lambda seed, y0=0, rows=None: _full("f", size, seed)[y0:(y0 + rows) if rows else None])
    alpha = "punch" if codec == 1 else "lefthalf"
    return dict(name=f"rgba8_gradnoise({alpha}) {size}x{size} RGBA8", fmt=synth.FMT_RGBA8, bpt=4,
                gen=lambda seed, y0=0, rows=None: synth.rgba8_gradnoise(size, size, seed, alpha, y0, rows))


def cpu_reference_sample(codec: int, size: int, seed: int, budget_s: float, steps: int = 1):
    """Times the unmodified reference on `steps` bounded samples: evenly spaced block-rows of the workload image,
    all host threads.  Returns (Mpix/s, dict)."""
    import oracle
    if not oracle.have_ref():
        return None, {"unavailable": "oracle/_ref/libref_oracle.so missing"}
    ref = oracle.RefOracle()
    wl = workload(codec, size)
    threads = ref.hw_threads()
    blocks_y = size // 4
    # probe: one strip of 4 block-rows' worth split over threads to estimate the rate
    probe_rows = max(1, min(blocks_y, threads // 8 or 1))

    def strip_image(rows_idx):
        parts = [wl["gen"](seed, 4 * r, 4) for r in rows_idx]
        return np.ascontiguousarray(np.concatenate(parts, axis=0))

    idx = [int(i * blocks_y / probe_rows) for i in range(probe_rows)]
    img = strip_image(idx)
    t0 = time.perf_counter()
    ref.encode(REF_CODEC[codec], img, wl["fmt"] if wl["fmt"] != 11 else 11, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = probe_rows / dt  # block-rows per second
    n_rows = int(max(threads // 4 or 1, min(blocks_y, rate * budget_s)))
    idx = [int(i * blocks_y / n_rows) for i in range(n_rows)]
    img = strip_image(idx)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref_blocks = ref.encode(REF_CODEC[codec], img, wl["fmt"], threads=threads)
        times.append(time.perf_counter() - t0)
    cpu_reference_sample.last = (idx, ref_blocks, img)
    dt = sum(times) / len(times)
    mpix = n_rows * 4 * size / dt / 1e6
    return mpix, {"value": mpix, "unit": "Mpix/s", "cores": threads, "kind": "reference",
                  "sample": f"{n_rows} evenly spaced block-rows ({n_rows * (size // 4)} blocks) of the {size}x{size} workload image, "
                            f"{dt:.2f} s per pass, unmodified reference via oracle/_ref, {threads} host threads over row strips"}


def bench_batch(args, codec, cname, rank, local_rank, world):
    """BASELINE config[4]: a batch of RGBA8 textures with full box-filtered mip chains, BC7, block-row sharded over the
    ranks by b200ic_plan_shards (strong scaling: the batch is fixed, each rank encodes its shards; no data-path
    collective, the gather of the blocks is outside the timed region)."""
    import torch
    import torch.distributed as dist
    import gfx_imagecompress_b200 as g
    from gfx_imagecompress_b200 import sharded, synth
    if args.impl == "reference":
        if rank == 0:
            mpix, info = cpu_reference_sample(codec, args.tex_size, 100, args.cpu_budget, max(1, args.steps))
            print(json.dumps({"impl": "reference", "metric": f"{cname.upper()} Mpix/s, texture batch", "value": mpix, "unit": "Mpix/s",
                              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "data": "synthetic", "cpu_baseline": info,
                              "config": {"workload": f"top level of one {args.tex_size}^2 texture (sample)"},
                              "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g.load_library()
    g.init(local_rank)
    unique = 8  # 8 x 21 MiB of distinct mip chains: the input working set exceeds the 126 MB L2
    chains = []
    for u in range(unique):
        top = torch.from_numpy(synth.rgba8_gradnoise(args.tex_size, args.tex_size, 100 + u, "lefthalf")).to(dev)
        chains.append(sharded.box_mips(top))
    images = [lvl for t in range(args.textures) for lvl in chains[t % unique]]
    dims = [(int(t.shape[1]), int(t.shape[0])) for t in images]
    mine = g.plan_shards(dims, world, rank)
    bb = g.BLOCK_BYTES[codec]
    # outputs only for the images this rank touches
    touched = sorted({i for i, _, _ in mine})
    outs = [None] * len(images)
    dummy = torch.empty((1, bb), dtype=torch.uint8, device=dev)
    for i in range(len(images)):
        outs[i] = dummy
    for i in touched:
        outs[i] = torch.empty((((dims[i][0] + 3) // 4) * ((dims[i][1] + 3) // 4), bb), dtype=torch.uint8, device=dev)
    my_blocks = sum(((dims[i][0] + 3) // 4) * (b - a) for i, a, b in mine)
    total_pix = sum(w * h for w, h in dims)

    def step():
        g.encode_batch_device(codec, images, synth.FMT_RGBA8, outs=outs, shards=mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) if codec != 7 else 1):
        step()
    barrier()
    n0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    launches = g.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    if rank == 0:
        peaks, how = _peaks()
        value = total_pix / 1e6 * args.steps / (total_ms / 1e3)
        alg = total_pix * 4 + sum(((w + 3) // 4) * ((h + 3) // 4) for w, h in dims) * bb
        achieved = alg * args.steps / (total_ms / 1e3) / 1e9
        print(json.dumps({
            "metric": f"{cname.upper()} Mpix/s, batch of {args.textures} x {args.tex_size}^2 RGBA8 textures with full mip chains",
            "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if codec == 7 else "f32", "data": f"synthetic ({unique} unique textures cycled)",
            "config": {"workload": f"{cname} encode of {args.textures} textures x {len(chains[0])} mip levels, block-row sharded "
                                   f"over {world} rank(s) by b200ic_plan_shards (BASELINE config[4] shape)", "codec": cname,
                       "textures": args.textures, "tex_size": args.tex_size, "levels": len(chains[0]),
                       "rank0_shards": len(mine), "rank0_blocks": my_blocks,
                       "l2_policy": "distinct outputs per texture; inputs cycle over 8 x 21 MiB distinct chains (178 MB > the 126 MB L2)"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": None, "peak_source": how, "kernel": f"{cname}_kernel"}}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _parity(codec, size, gpu_blocks, bb):
    """The block-rows the CPU reference just encoded for cpu_baseline, compared with the SAME rows of the GPU result of
    the timed workload (rank 0's image): identical-block fraction and, for BC7 / BC1, decoded PSNR of both."""
    last = getattr(cpu_reference_sample, "last", None)
    if last is None:
        return None
    idx, ref_blocks, img = last
    bx = size // 4
    g = gpu_blocks.reshape(size // 4, bx, bb)[idx].reshape(-1, bb)
    same = float((g == ref_blocks).all(axis=1).mean())
    out = {"rows": len(idx), "blocks": int(len(ref_blocks)), "identical_fraction": same}
    try:
        from oracle import metrics
        if codec in (7, 8):
            out["psnr_gpu_db"] = metrics.psnr_bc7(g, img)
            out["psnr_reference_db"] = metrics.psnr_bc7(ref_blocks, img)
        elif codec == 1:
            out["psnr_gpu_db"] = metrics.psnr_bc1(g, img)
            out["psnr_reference_db"] = metrics.psnr_bc1(ref_blocks, img)
        if "psnr_gpu_db" in out:
            out["delta_psnr_db"] = out["psnr_gpu_db"] - out["psnr_reference_db"]
    except Exception as e:  # the decoders are test infrastructure; parity by bytes stands without them
        out["psnr_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--codec", default="auto", choices=["auto"] + list(CODECS))
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU reference work for cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="image", choices=["image", "batch"],
                    help="image: one --size^2 image per rank per step (headline); batch: --textures x (--tex-size^2 + full mip "
                         "chain), block-row sharded over the ranks (BASELINE config[4])")
    ap.add_argument("--textures", type=int, default=256)
    ap.add_argument("--tex-size", type=int, default=2048)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import gfx_imagecompress_b200 as g
    if args.codec == "auto":
        g.load_library()
        codec = next(c for c in (7, 8, 1, 5) if g.codec_available(c))
    else:
        codec = CODECS[args.codec]
    cname = {v: k for k, v in CODECS.items()}[codec]
    if args.workload == "batch":
        return bench_batch(args, codec, cname, rank, local_rank, world)
    size = args.size
    wl = workload(codec, size)
    metric = f"{cname.upper()} Mpix/s at {size}^2 {wl['name'].split()[-1]}"
    in_bytes = size * size * wl["bpt"]
    need_flush = in_bytes <= (126 << 20)
    base_cfg = {1: 0, 4: 1, 5: 1, 6: 3, 7: 2, 8: 2}[codec]
    cfg = {"workload": f"{cname} encode of {wl['name']}, one image per rank per step (BASELINE config[{base_cfg}] shape)",
           "codec": cname, "image": [size, size],
           "l2_policy": (f"input {in_bytes >> 20} MiB fits the 126 MB L2: a 256 MiB buffer is overwritten between the timed steps "
                         "(outside the per-step events)") if need_flush else
                        f"input {in_bytes >> 20} MiB per step exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_sample(codec, size, 3, min(args.cpu_budget, 2.0), 1)
        mpix, info = cpu_reference_sample(codec, size, 3, args.cpu_budget, steps)
        if mpix is None:
            print(json.dumps({"impl": "reference", **info}))
            return 0
        print(json.dumps({"impl": "reference", "metric": metric, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": size * size / 1e6 / mpix * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if codec == 7 else "f32",
                          "data": "synthetic", "config": cfg, "cpu_baseline": info,
                          "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g.load_library()
    g.init(local_rank)

    # ---- inputs: one image per rank (seed differs per rank); pinned host copy for the e2e leg
    host = wl["gen"](3 + rank)
    host_t = torch.from_numpy(host).pin_memory()
    d_src = host_t.to(dev, non_blocking=True)
    nblocks = (size // 4) * (size // 4)
    bb = g.BLOCK_BYTES[codec]
    d_dst = torch.empty((nblocks, bb), dtype=torch.uint8, device=dev)
    h_dst = torch.empty((nblocks, bb), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        g.encode_device(codec, d_src, wl["fmt"], size, size, 1, out=d_dst)

    host_np = host_t.numpy()
    h_dst_np = h_dst.numpy()

    def step_host():
        g.encode_host(codec, host_np, wl["fmt"], out=h_dst_np)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()

    # ---- device-resident timing (CUDA events on the launching = torch current stream)
    n0 = g.launch_count()
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None
    barrier()
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            if flush is not None:
                flush.fill_(i & 0xFF)  # evicts the image from L2; not inside the step's events
            ev_s[i].record()
            step_device()
            ev_e[i].record()
        barrier()
    launches = g.launch_count() - n0
    per_launch_ms = [ev_s[i].elapsed_time(ev_e[i]) for i in range(args.steps)]
    total_ms = sum(per_launch_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())

    # ---- end to end through the host-buffer C-ABI (pinned host in, host out)
    for _ in range(1):
        step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    # the device result of the e2e leg must equal the device-resident result
    same = bool(torch.equal(torch.from_numpy(h_dst_np).to(dev), d_dst))

    if rank == 0:
        mpix_step = size * size / 1e6
        value = world * mpix_step * args.steps / (total_ms_max / 1e3)
        e2e_v = world * mpix_step * e2e_steps / float(e2e_s.item())
        peaks, how = _peaks()
        alg_bytes = size * size * wl["bpt"] + nblocks * bb
        k_ms = sum(per_launch_ms) / len(per_launch_ms)
        achieved = alg_bytes / (k_ms / 1e3) / 1e9
        ctr = _counters(cname)
        clk = clocks.summary()
        out = {"metric": metric, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64" if codec == 7 else "f32", "data": "synthetic", "config": cfg,
               "e2e": {"value": e2e_v, "unit": "Mpix/s", "h2d_bytes_per_step": int(host_np.nbytes), "d2h_bytes_per_step": int(nblocks * bb),
                       "matches_device_path": same},
               "gpu_launches": int(launches), "clocks": clk,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / peaks["hbm_gbs"],
                            "traffic": (ctr["dram_bytes_per_block"] * nblocks) if ctr else None, "peak_source": how,
                            "kernel": f"{cname}_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                            "launches_per_step": int(launches) // max(args.steps, 1),
                            "note": "achieved / traffic / algorithmic bytes are per step (= per image: every launch of the step "
                                    "together; AMD BC7 is one launch per mode); traffic = ncu dram bytes per block of "
                                    + (f"profiles/ncu_counters/{ctr['report']} ({ctr['blocks']} blocks)" if ctr else "n/a") +
                                    " x the blocks of this workload. The per-block search is ALU bound (SURVEY.md 8d): see `alu`"},
               "blocks_per_s": world * nblocks * args.steps / (total_ms_max / 1e3)}
        out["alu"] = _alu(cname, out["blocks_per_s"] / world, clk.get("sm_mhz"))
        if not args.no_cpu and world == 1:
            _, info = cpu_reference_sample(codec, size, 3, args.cpu_budget, 1)
            out["cpu_baseline"] = info
            out["parity"] = _parity(codec, size, d_dst.cpu().numpy(), bb)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
