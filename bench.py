#!/usr/bin/env python3
"""bench.py -- BCn block-compression throughput on B200 (BASELINE.json metric: BC7 Mpix/s at 8192^2 RGBA8).

A "step" = one encode of ONE 8192x8192 RGBA8 synthetic image (left half translucent, right half opaque -- BASELINE
config[2]) with the AMD BC7 encoder.

  N = 1   the whole image on one GPU.
  N > 1   STRONG scaling (one rank per GPU under torchrun): the block-rows of the one image are dealt to the ranks by
          b200ic_plan_shards, every rank encodes its shards, no data-path collective; `value` = image pixels / the
          slowest rank's time.  The only exchange is the final gather of the 16-byte blocks to rank 0 (NCCL), which is
          inside the `e2e` timer and reported on its own (`gather_ms`).

  value        device-resident: input already in HBM, CUDA events on the launching stream, per step
  e2e          same metric through the reference's own entry point: Image_CompressAMDBC7(Image_ImageHeader*) on a
               pageable (malloc) source image, returning a malloc'd compressed image -- host->device and device->host copies
               inside the timed region, every step
  roofline     the dominant kernel of the step (AMD BC7: the ep_shaker_d cube kernel of mode 0), ALU issue side: executed
               lane-instructions of that launch (ncu count, profiles/ncu_counters.json, scaled by blocks) / its duration
               measured live with CUDA events on the launching stream (b200ic_profile) vs 148 SMs x 128 lanes x SM clock.
               `hbm` is the same launch against the measured copy bandwidth: the search is ALU bound by design
  cpu_baseline the unmodified reference (oracle/_ref/libref_oracle.so) on ALL host threads (threads_used is reported and
               must equal cores), bounded sample of evenly spaced block-rows of the same image
  parity       64 evenly spaced block-rows of the GPU result vs the reference's bytes (+ decoded PSNR of both)
  configs      the other BASELINE configs (BC1 1024^2, BC4 / BC5 4096^2, BC6H 4096^2, bc7enc16 8192^2), each with its own
               device / e2e / ALU / parity / clocks record (N = 1 only)
  --impl reference   times only the CPU reference on the same workload; never loads the CUDA library.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

CODECS = {"bc1": 1, "bc2": 2, "bc3": 3, "bc4": 4, "bc5": 5, "bc6h": 6, "bc7_amd": 7, "bc7_rg": 8}
NAMES = {v: k for k, v in CODECS.items()}
BLOCK_BYTES = {1: 8, 2: 16, 3: 16, 4: 8, 5: 16, 6: 16, 7: 16, 8: 16}
DTYPE = {1: "f32", 2: "f32", 3: "f32", 4: "f32", 5: "f32", 6: "f32", 7: "f64+int32", 8: "int32+f32"}
BASE_CFG = {1: 0, 4: 1, 5: 1, 6: 3, 7: 2, 8: 2, 2: 0, 3: 0}
FMT_RG8, FMT_RGBA8, FMT_RGBA16UF = 3, 7, 11
# AMD BC7 launch order of one encode (csrc/bc7amd.cu): (mode, kind) with kind 0 quantise, 1 cube, 2 window, 3 thread-per-block
AMD_LAUNCHES = [(6, 3)] + [(m, k) for m in (4, 3, 1, 2, 0) for k in (0, 1, 2)] + [(7, 0), (7, 2)] + [(5, 0), (5, 1), (5, 2)]
KIND = ["amd_quant_kernel", "amd_cube_kernel", "amd_window_kernel", "bc7amd_serial_kernel"]


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def _counters(cname):
    """Per-block / per-launch counters of the codec's kernel(s) from `ncu --metrics` captures (tools/ncu_counters.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as f:
            return json.load(f).get(cname)
    except Exception:
        return None


def _alu_peak(sm_mhz):
    return 148 * 128 * float(sm_mhz or 1965.0) * 1e6


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period: float = 0.1):
        self.index, self.period = index, period
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
            self._stop.wait(self.period)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
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
        return dict(name=f"height_rg8 {size}x{size} RG8", fmt=FMT_RG8, bpt=2,
                    gen=lambda seed, y0=0, rows=None: _full("h", size, seed)[y0:(y0 + rows) if rows else None])
    if codec == 6:
        return dict(name=f"hdr_rgba16f {size}x{size} RGBA16F(unsigned)", fmt=FMT_RGBA16UF, bpt=8,
                    gen=lambda seed, y0=0, rows=None: _full("f", size, seed)[y0:(y0 + rows) if rows else None])
    alpha = "punch" if codec == 1 else "lefthalf"
    return dict(name=f"rgba8_gradnoise({alpha}) {size}x{size} RGBA8", fmt=FMT_RGBA8, bpt=4,
                gen=lambda seed, y0=0, rows=None: synth.rgba8_gradnoise(size, size, seed, alpha, y0, rows))


# ---- CPU reference legs (the only code that may execute oracle/) -------------------------------------------------------

class CpuRef:
    """The unmodified reference on the host cores: evenly spaced block-rows of the workload image, every host thread busy
    (oracle/ref_harness.cpp cuts the rows into block-column tiles when there are fewer rows than threads)."""

    def __init__(self, codec, size, seed):
        import oracle
        self.ok = oracle.have_ref()
        self.codec, self.size, self.seed = codec, size, seed
        if not self.ok:
            return
        self.ref = oracle.RefOracle()
        self.ref.lib.ref_last_threads_used.restype = C.c_int
        self.wl = workload(codec, size)
        self.threads = self.ref.hw_threads()
        self.blocks_y = size // 4

    def rows_image(self, rows_idx):
        return np.ascontiguousarray(np.concatenate([self.wl["gen"](self.seed, 4 * r, 4) for r in rows_idx], axis=0))

    def encode_rows(self, rows_idx):
        img = self.rows_image(rows_idx)
        t0 = time.perf_counter()
        blocks = self.ref.encode(self.codec, img, self.wl["fmt"], threads=self.threads)
        dt = time.perf_counter() - t0
        return blocks, img, dt, int(self.ref.lib.ref_last_threads_used())

    def grid(self, n):
        n = max(1, min(n, self.blocks_y))
        return [int(i * self.blocks_y / n) for i in range(n)]

    def rows_for_budget(self, budget_s, lo=1, hi=None, multiple=1):
        """Probe one block-row, then pick how many fit in `budget_s` seconds (a multiple of `multiple`)."""
        _, _, dt, _ = self.encode_rows(self.grid(1))
        n = int(budget_s / max(dt, 1e-4))
        n = max(lo, min(n, hi or self.blocks_y))
        return max(multiple, (n // multiple) * multiple)

    def timed(self, n_rows, steps=1):
        idx = self.grid(n_rows)
        times, used = [], 0
        for _ in range(steps):
            blocks, img, dt, used = self.encode_rows(idx)
            times.append(dt)
        dt = sum(times) / len(times)
        mpix = len(idx) * 4 * self.size / dt / 1e6
        info = {"value": mpix, "unit": "Mpix/s", "cores": self.threads, "threads_used": used, "kind": "reference",
                "sample": f"{len(idx)} evenly spaced block-rows ({len(idx) * (self.size // 4)} blocks) of the {self.size}x{self.size} "
                          f"workload image, {dt:.2f} s per pass, unmodified reference via oracle/_ref/libref_oracle.so, "
                          f"{used} of {self.threads} host threads busy (block-row x block-column tiles)"}
        assert used == min(self.threads, len(idx) * (self.size // 4)), f"CPU reference used {used} of {self.threads} threads"
        return mpix, info, (idx, blocks, img)


def parity(codec, size, gpu_blocks, idx, ref_blocks, img):
    """Block-rows `idx` of the GPU result against the reference's bytes; decoded PSNR of both where a decoder applies."""
    bb = BLOCK_BYTES[codec]
    bx = size // 4
    g = gpu_blocks.reshape(size // 4, bx, bb)[idx].reshape(-1, bb)
    same = (g == ref_blocks).all(axis=1)
    out = {"rows": len(idx), "blocks": int(len(ref_blocks)), "identical_fraction": float(same.mean()),
           "differing_blocks": int((~same).sum())}
    try:
        from oracle import metrics
        if codec in (7, 8):
            out["psnr_gpu_db"], out["psnr_reference_db"] = metrics.psnr_bc7(g, img), metrics.psnr_bc7(ref_blocks, img)
        elif codec == 1:
            out["psnr_gpu_db"], out["psnr_reference_db"] = metrics.psnr_bc1(g, img), metrics.psnr_bc1(ref_blocks, img)
        elif codec == 6:
            h = img.view(np.float16)
            out["psnr_gpu_db"], out["psnr_reference_db"] = metrics.psnr_bc6h(g, h), metrics.psnr_bc6h(ref_blocks, h)
        if "psnr_gpu_db" in out:
            out["delta_psnr_db"] = out["psnr_gpu_db"] - out["psnr_reference_db"]
    except Exception as e:  # the decoders are test infrastructure; parity by bytes stands without them
        out["psnr_error"] = repr(e)
    return out


def reference_arm(args, codec, size, rank):
    """--impl reference: the reference's own CPU implementation on this host, all threads, bounded sample per step."""
    if rank != 0:
        return 0
    cname = NAMES[codec]
    wl = workload(codec, size)
    cpu = CpuRef(codec, size, 3)
    if not cpu.ok:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so missing"}))
        return 0
    steps = max(1, args.steps)
    # whole run (warm-up + steps) within a few minutes: at most ~150 s of CPU work in total
    per_step = min(args.cpu_budget, 150.0 / (steps + min(args.warmup, 1)))
    n_rows = cpu.rows_for_budget(per_step, lo=1)
    for _ in range(min(args.warmup, 1)):
        cpu.timed(n_rows, 1)
    mpix, info, _ = cpu.timed(n_rows, steps)
    print(json.dumps({"impl": "reference", "metric": f"{cname.upper()} Mpix/s at {size}^2 {wl['name'].split()[-1]}", "value": mpix,
                      "unit": "Mpix/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
                      "ms_per_step": size * size / 1e6 / mpix * 1e3, "higher_is_better": True,
                      "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": DTYPE[codec], "data": "synthetic",
                      "config": {"workload": f"{cname} encode of {wl['name']} (BASELINE config[{BASE_CFG[codec]}] shape); each step = a bounded "
                                             "sample of evenly spaced block-rows, scaled to the image", "codec": cname, "image": [size, size]},
                      "cpu_baseline": info,
                      "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


# ---- GPU legs ----------------------------------------------------------------------------------------------------------

class ImageApi:
    """The reference-facing entry points on a pageable Image_ImageHeader (48-byte header + texels in one malloc'd-style
    buffer), the returned compressed image viewed in place and freed with libc free (= Image_Destroy of the shim)."""
    FN = {1: "Image_CompressAMDBC1", 2: "Image_CompressAMDBC2", 3: "Image_CompressAMDBC3", 4: "Image_CompressAMDBC4",
          5: "Image_CompressAMDBC5", 6: "Image_CompressAMDBC6H", 7: "Image_CompressAMDBC7", 8: "Image_CompressRichGel999BC7"}

    def __init__(self, g, codec, pixels, fmt):
        self.g, self.codec = g, codec
        self.img = g.Image(pixels, fmt)  # ctypes buffer: ordinary pageable memory
        self.fn = getattr(g.library(), self.FN[codec])
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        self.nargs = {1: 5, 4: 3, 5: 3}.get(codec, 4)

    def __call__(self, keep=False):
        addr = self.fn(self.img.ptr, *([None] * (self.nargs - 1)))
        if not addr:
            raise SystemExit(f"{self.FN[self.codec]} failed: {self.g.library().b200ic_last_error().decode()}")
        hdr = self.g.api._ImageHeader.from_address(addr)
        out = np.ctypeslib.as_array((C.c_uint8 * hdr.dataSize).from_address(addr + 48)).copy() if keep else None
        self.libc.free(C.c_void_p(addr))
        return out


def amd_profile(g, nblocks, steps, sm_mhz, peaks):
    """Reads the per-kernel event times recorded by the library (b200ic_profile) and builds the roofline of the dominant
    kernel from the ncu instruction / DRAM counts of the same kernel (profiles/ncu_counters.json).  An encode runs the 21
    kernels of AMD_LAUNCHES once per chunk of <= 2^19 blocks; both sides are summed over the chunks of one image."""
    L = g.library()
    ms = (C.c_double * 32)()
    cnt = (C.c_uint64 * 32)()
    L.b200ic_profile_read(ms, cnt)
    per = {(m, k): (ms[m * 4 + k], int(cnt[m * 4 + k])) for m in range(8) for k in range(4) if cnt[m * 4 + k]}
    if not per:
        return None, None
    total = sum(v[0] for v in per.values())
    (m, k), (t, n) = max(per.items(), key=lambda kv: kv[1][0])
    ms_per_image = t / steps
    table = [{"kernel": KIND[kk], "mode": mm, "ms_per_step": v[0] / steps, "launches_per_step": v[1] // steps, "share": v[0] / total}
             for (mm, kk), v in sorted(per.items(), key=lambda kv: -kv[1][0])]
    ctr = _counters("bc7_amd")
    roof = None
    if ctr and "per_launch" in ctr and len(ctr["per_launch"]) == len(AMD_LAUNCHES):
        pl = ctr["per_launch"][AMD_LAUNCHES.index((m, k))]
        scale = nblocks / ctr["blocks"]
        lane_ops = pl["thread_inst"] * scale
        peak = _alu_peak(sm_mhz)
        ach = lane_ops / (ms_per_image / 1e3)
        alg = nblocks * 80 + (nblocks * 584 if k != 3 else 0)  # texels + block (+ the phase kernels' hand-over words)
        roof = {"bound": "alu", "kernel": f"{KIND[k]} (mode {m})", "achieved": ach / 1e12, "peak": peak / 1e12, "unit": "T lane-ops/s",
                "frac": ach / peak, "traffic": pl["dram_bytes"] * scale, "ms_per_launch": t / n, "launches_per_step": n // steps,
                "ms_per_step": ms_per_image, "share_of_step": t / total, "lane_ops_per_step": lane_ops,
                "lanes_per_inst": pl["thread_inst"] / max(pl["warp_inst"], 1.0),
                "peak_source": "148 SMs x 128 FP32/INT32 lanes x median SM clock sampled during the timed region",
                "counters": "profiles/ncu_counters/" + ctr["report"],
                "hbm": {"achieved": alg / (ms_per_image / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": alg / (ms_per_image / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_step": alg},
                "note": "duration = CUDA events around this kernel's launches on the launching stream inside the timed steps "
                        "(b200ic_profile), summed over the chunks of one image; lane-ops / DRAM bytes = ncu counts of the same kernel "
                        "on the same workload, summed the same way. The search is ALU bound (SURVEY.md 8d): the HBM fraction is "
                        "the evidence, not the target"}
    return roof, table


def single_kernel_roofline(cname, nblocks, avg_ms, alg_bytes, sm_mhz, peaks, how):
    ctr = _counters(cname)
    hbm = {"achieved": alg_bytes / (avg_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "frac": alg_bytes / (avg_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes, "peak_source": how}
    if not ctr:
        return {"bound": "hbm", **hbm, "traffic": None, "kernel": f"{cname}_kernel"}
    lane_ops = ctr["thread_inst_per_block"] * nblocks
    peak = _alu_peak(sm_mhz)
    ach = lane_ops / (avg_ms / 1e3)
    return {"bound": "alu", "kernel": ctr["kernels"][0] if ctr.get("kernels") else cname, "achieved": ach / 1e12, "peak": peak / 1e12,
            "unit": "T lane-ops/s", "frac": ach / peak, "traffic": ctr["dram_bytes_per_block"] * nblocks, "ms_per_launch": avg_ms,
            "lanes_per_inst": ctr["thread_inst_per_block"] / max(ctr["warp_inst_per_block"], 1.0),
            "counters": "profiles/ncu_counters/" + ctr["report"], "hbm": hbm}


def run_single_gpu(g, torch, dev, codec, size, steps, warmup, cpu_budget, with_cpu, parity_rows, local_rank, full=True):
    """Device-resident + e2e (Image API, pageable) + roofline + cpu baseline + parity of one codec / size on one GPU."""
    cname = NAMES[codec]
    wl = workload(codec, size)
    bb = BLOCK_BYTES[codec]
    nblocks = (size // 4) * (size // 4)
    host = wl["gen"](3)
    d_src = torch.from_numpy(host).to(dev)
    d_dst = torch.empty((nblocks, bb), dtype=torch.uint8, device=dev)
    in_bytes = host.nbytes
    need_flush = in_bytes <= (126 << 20)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None

    def step_device():
        g.encode_device(codec, d_src, wl["fmt"], size, size, 1, out=d_dst)

    for _ in range(max(warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    n0 = g.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    L = g.library()
    L.b200ic_profile(1 if codec == 7 else 0)
    with ClockSampler(local_rank, 0.05 if codec != 7 else 0.1) as clocks:
        for i in range(steps):
            if flush is not None:
                flush.fill_(i & 0xFF)  # evicts the image from L2; outside the step's events
            ev[i][0].record()
            step_device()
            ev[i][1].record()
        torch.cuda.synchronize()
        launches = g.launch_count() - n0
        if codec != 7:  # sub-10 ms kernels: keep the sampler alive long enough for >= 10 clock samples under load
            t_end = time.perf_counter() + 2.5
            while time.perf_counter() < t_end:
                step_device()
            torch.cuda.synchronize()
    L.b200ic_profile(0)
    per_ms = [a.elapsed_time(b) for a, b in ev]
    clk = clocks.summary()
    peaks, how = _peaks()
    alg_bytes = in_bytes + nblocks * bb
    if codec == 7:
        roof, table = amd_profile(g, nblocks, steps, clk.get("sm_mhz"), peaks)
    else:
        roof, table = single_kernel_roofline(cname, nblocks, sum(per_ms) / len(per_ms), alg_bytes, clk.get("sm_mhz"), peaks, how), None
    gpu_blocks = d_dst.cpu().numpy()

    # ---- end to end through the reference's entry point on pageable memory, every step
    g.set_devices(1)
    api = ImageApi(g, codec, host, wl["fmt"])
    e2e_blocks = api(keep=True).reshape(-1, bb)  # (also the warm-up: staging buffers, contexts)
    t0 = time.perf_counter()
    for _ in range(steps):
        api()
    e2e_s = time.perf_counter() - t0
    mpix_step = size * size / 1e6
    out = {"codec": cname, "image": [size, size], "value": mpix_step * steps / (sum(per_ms) / 1e3), "unit": "Mpix/s",
           "ms_per_step": sum(per_ms) / steps, "steps": steps,
           "e2e": {"value": mpix_step * steps / e2e_s, "unit": "Mpix/s", "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(nblocks * bb),
                   "path": ImageApi.FN[codec], "host_memory": "pageable", "steps": steps, "ms_per_step": e2e_s / steps * 1e3,
                   "matches_device_path": bool(np.array_equal(e2e_blocks, gpu_blocks))},
           "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
           "l2_policy": ("a 256 MiB buffer is overwritten between the timed steps (outside the per-step events)" if need_flush
                         else f"input {in_bytes >> 20} MiB per step exceeds the 126 MB L2"),
           "blocks_per_s": nblocks * steps / (sum(per_ms) / 1e3)}
    if table:
        out["kernels"] = table
    if with_cpu:
        cpu = CpuRef(codec, size, 3)
        if cpu.ok:
            # parity grid first (64 rows where the budget allows: BC7-AMD costs ~0.5 ms per block and thread), the timed sample
            # is a sub-grid of it
            timed_rows = cpu.rows_for_budget(cpu_budget, lo=min(16, cpu.blocks_y), hi=parity_rows, multiple=1)
            _, info, (idx, ref_blocks, img) = cpu.timed(timed_rows, 1)
            out["cpu_baseline"] = info
            if timed_rows < parity_rows and full:
                idx = cpu.grid(parity_rows)
                ref_blocks, img, _, _ = cpu.encode_rows(idx)
            out["parity"] = parity(codec, size, gpu_blocks, idx, ref_blocks, img)
    return out


def bench_batch(args, codec, cname, rank, local_rank, world):
    """BASELINE config[4]: a batch of RGBA8 textures with full box-filtered mip chains, BC7, block-row sharded over the
    ranks by b200ic_plan_shards (strong scaling: the batch is fixed, each rank encodes its shards; no data-path
    collective, the gather of the blocks is outside the timed region)."""
    import torch
    import torch.distributed as dist
    import gfx_imagecompress_b200 as g
    from gfx_imagecompress_b200 import sharded, synth
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
    touched = sorted({i for i, _, _ in mine})
    dummy = torch.empty((1, bb), dtype=torch.uint8, device=dev)
    outs = [dummy] * len(images)
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
            "dtype": DTYPE[codec], "data": f"synthetic ({unique} unique textures cycled)",
            "config": {"workload": f"{cname} encode of {args.textures} textures x {len(chains[0])} mip levels, block-row sharded "
                                   f"over {world} rank(s) by b200ic_plan_shards (BASELINE config[4] shape; config[4] names 1024 textures)",
                       "codec": cname, "textures": args.textures, "tex_size": args.tex_size, "levels": len(chains[0]),
                       "rank0_shards": len(mine), "rank0_blocks": my_blocks,
                       "l2_policy": "distinct outputs per texture; inputs cycle over 8 x 21 MiB distinct chains (178 MB > the 126 MB L2)"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": None, "peak_source": how, "kernel": f"{cname} kernels of the batch"}}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_sharded_image(args, g, torch, dist, codec, size, rank, local_rank, world, dev):
    """N > 1: ONE image, block-rows dealt to the ranks (strong scaling)."""
    cname = NAMES[codec]
    wl = workload(codec, size)
    bb = BLOCK_BYTES[codec]
    bx = size // 4
    nblocks = bx * bx
    plans = [g.plan_shards([(size, size)], world, r) for r in range(world)]
    mine = plans[rank]
    host = wl["gen"](3)  # every rank generates the same image and reads only its rows
    d_src = torch.from_numpy(host).to(dev)
    d_dst = torch.zeros((nblocks, bb), dtype=torch.uint8, device=dev)
    my_rows = sum(b - a for _, a, b in mine)
    cap = max(sum(b - a for _, a, b in p) for p in plans) * bx * bb
    flat = torch.zeros(cap, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev) if rank == 0 else None
    h_out = torch.empty((nblocks, bb), dtype=torch.uint8).pin_memory() if rank == 0 else None

    def step_device():
        g.encode_batch_device(codec, [d_src], wl["fmt"], outs=[d_dst], shards=mine)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    n0 = g.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    g.library().b200ic_profile(1 if (codec == 7 and rank == 0) else 0)
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            ev[i][0].record()
            step_device()
            ev[i][1].record()
        barrier()
    g.library().b200ic_profile(0)
    launches = g.launch_count() - n0
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    roof = None
    if codec == 7 and rank == 0:  # rank 0's share of the image through the same per-kernel event timing as the N = 1 line
        roof, _ = amd_profile(g, my_rows * bx, args.steps, clocks.summary().get("sm_mhz"), _peaks()[0])
        if roof:
            roof["note"] = "rank 0's launches (its share of the block-rows); " + roof["note"]

    # ---- e2e: pageable host rows -> this rank's GPU -> encode -> gather of the blocks to rank 0 -> host, every step
    def step_e2e():
        pos = 0
        for _, r0, r1 in mine:
            rows = torch.from_numpy(host[4 * r0:4 * r1])  # pageable numpy memory
            d_src[4 * r0:4 * r1].copy_(rows, non_blocking=False)
        step_device()
        for _, r0, r1 in mine:
            n = (r1 - r0) * bx * bb
            flat[pos:pos + n] = d_dst[r0 * bx:r1 * bx].reshape(-1)
            pos += n
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.gather(flat, [gathered[r * cap:(r + 1) * cap] for r in range(world)] if rank == 0 else None, dst=0)
        g1.record()
        if rank == 0:
            for r in range(world):
                pos = r * cap
                for _, r0, r1 in plans[r]:
                    n = (r1 - r0) * bx * bb
                    h_out[r0 * bx:r1 * bx].copy_(gathered[pos:pos + n].reshape(-1, bb), non_blocking=True)
                    pos += n
        torch.cuda.synchronize()
        return g0.elapsed_time(g1)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    gather_ms = 0.0
    for _ in range(args.steps):
        gather_ms += step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if rank == 0:
        mpix_step = size * size / 1e6
        same = bool(torch.equal(h_out.to(dev)[mine[0][1] * bx:mine[0][2] * bx], d_dst[mine[0][1] * bx:mine[0][2] * bx]))
        clk = clocks.summary()
        print(json.dumps({
            "metric": f"{cname.upper()} Mpix/s at {size}^2 {wl['name'].split()[-1]}", "value": mpix_step * args.steps / (total_ms / 1e3),
            "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE[codec], "data": "synthetic",
            "config": {"workload": f"{cname} encode of ONE {wl['name']} (BASELINE config[{BASE_CFG[codec]}] shape), block-rows sharded over "
                                   f"{world} ranks by b200ic_plan_shards (chunks of 64 block-rows, one contiguous run per rank); no data-path collective",
                       "codec": cname, "image": [size, size], "rank0_block_rows": my_rows,
                       "l2_policy": f"input {host.nbytes >> 20} MiB per step exceeds the 126 MB L2"},
            "e2e": {"value": mpix_step * args.steps / float(e2e_s.item()), "unit": "Mpix/s",
                    "h2d_bytes_per_step": int(host.nbytes), "d2h_bytes_per_step": int(nblocks * bb),
                    "path": "per rank: pageable rows -> H2D -> b200ic_encode_batch_device(shards) -> NCCL gather to rank 0 -> D2H",
                    "host_memory": "pageable in, pinned out", "gather_ms_per_step": gather_ms / args.steps,
                    "gather_bytes_per_step": int(nblocks * bb), "ms_per_step": float(e2e_s.item()) / args.steps * 1e3,
                    "rank0_rows_match_device_path": same},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": roof or {"bound": "alu", "kernel": "see the N = 1 line (same kernels, 1/N of the blocks per rank)", "achieved": None,
                                 "peak": None, "unit": "T lane-ops/s", "frac": None, "traffic": None}}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--codec", default="auto", choices=["auto"] + list(CODECS))
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU reference work for the timed cpu_baseline sample")
    ap.add_argument("--parity-rows", type=int, default=64, help="evenly spaced block-rows compared with the CPU reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs after the headline")
    ap.add_argument("--workload", default="image", choices=["image", "batch"],
                    help="image: one --size^2 image per step (headline; sharded by block-row for N > 1); batch: --textures x "
                         "(--tex-size^2 + full mip chain), block-row sharded over the ranks (BASELINE config[4])")
    ap.add_argument("--textures", type=int, default=256)
    ap.add_argument("--tex-size", type=int, default=2048)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    codec = 7 if args.codec == "auto" else CODECS[args.codec]
    cname = NAMES[codec]
    size = args.size
    if args.impl == "reference":  # the CPU reference only: the CUDA library is never loaded in this arm
        return reference_arm(args, codec, size, rank)

    import gfx_imagecompress_b200 as g
    if args.workload == "batch":
        return bench_batch(args, codec, cname, rank, local_rank, world)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    g.load_library()
    g.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        rc = bench_sharded_image(args, g, torch, dist, codec, size, rank, local_rank, world, dev)
        dist.destroy_process_group()
        return rc

    wl = workload(codec, size)
    head = run_single_gpu(g, torch, dev, codec, size, args.steps, args.warmup, args.cpu_budget, not args.no_cpu, args.parity_rows, local_rank)
    out = {"metric": f"{cname.upper()} Mpix/s at {size}^2 {wl['name'].split()[-1]}", "value": head["value"], "unit": "Mpix/s", "n_gpus": 1,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": DTYPE[codec], "data": "synthetic",
           "config": {"workload": f"{cname} encode of {wl['name']}, one image per step (BASELINE config[{BASE_CFG[codec]}] shape)",
                      "codec": cname, "image": [size, size], "l2_policy": head["l2_policy"]},
           "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
           "blocks_per_s": head["blocks_per_s"]}
    for k in ("kernels", "cpu_baseline", "parity"):
        if k in head:
            out[k] = head[k]
    if args.codec == "auto" and not args.no_configs:
        # the other BASELINE configs, same contract per entry (short: their kernels run in milliseconds)
        out["configs"] = []
        for c, s in ((1, 1024), (4, 4096), (5, 4096), (6, 4096), (8, 8192)):
            try:
                # BC1 / BC4 / BC5 are cheap on the CPU: 64 parity rows; BC6H / bc7enc16: the rows the CPU budget allows
                r = run_single_gpu(g, torch, dev, c, s, max(args.steps, 10), args.warmup, min(args.cpu_budget, 6.0), not args.no_cpu,
                                   args.parity_rows, local_rank, full=(c in (1, 4, 5)))
                r["baseline_config"] = BASE_CFG[c]
                out["configs"].append(r)
            except Exception as e:  # a failing side config must not take the headline down; it is reported
                out["configs"].append({"codec": NAMES[c], "error": repr(e)})
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
