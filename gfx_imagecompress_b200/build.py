"""In-tree build of libgfx_imagecompress_b200.so (sm_100a only) with nvcc.

One translation unit per codec: the bit-exact codecs (BC1/BC4/BC5/bc7enc16) are compiled with --fmad=false
so that no FP32 multiply-add is contracted (the reference's output changes under contraction),
the AMD BC7 and BC6H kernels likewise (they then reproduce the FP64 / FP32 reference bit for bit).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("B200IC_BUILD_DIR") or os.path.join(HERE, "lib")  # (debug builds go elsewhere: B200IC_BUILD_DIR)
LIB = os.path.join(OUT, "libgfx_imagecompress_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
          "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "compat"), "-I" + CSRC]

# (source, extra flags)
UNITS = [
    ("api.cu", []),
    ("container.cu", []),
    ("decode.cu", []),
    ("bc45.cu", ["--fmad=false"]),
    ("bc1.cu", ["--fmad=false"]),
    ("bc7rg.cu", ["--fmad=false"]),
    ("bc7amd.cu", ["--fmad=false"]),
    ("bc6h.cu", ["--fmad=false"]),
    ("image_shim.cpp", []),
]

def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", "b200ic.h"),
                os.path.join(ROOT, "include", "gfx_imagecompress", "imagecompress.h"), os.path.abspath(__file__)]
    present = [(s, f) for s, f in UNITS if os.path.exists(os.path.join(CSRC, s))]
    defines = ["-D" + d for d in os.environ.get("B200IC_EXTRA_DEFS", "").split() if d]  # e.g. B200IC_AMD_TIMING (debug builds)
    cc = nvcc()
    jobs = []
    objs = []
    for src, extra in present:
        obj = os.path.join(OUT, "obj", src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        sp = os.path.join(CSRC, src)
        if force or _stale(obj, [sp] + headers):
            cmd = [cc] + ARCH + COMMON + defines + extra + (["-Xptxas", "-v"] if ptxas_info else [])
            if src.endswith(".cpp"):
                cmd += ["-x", "cu"]
            cmd += ["-c", sp, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if (verbose or ptxas_info) and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, flush=True)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([cc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xlinker", "-Bsymbolic", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
