"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the compiled, unmodified reference).

CPU: the compiled reference still reproduces them (drift guard, build container only), the plain-C restatement and the
host builds of the kernels' per-block cores reproduce them.  GPU: the CUDA kernels through the C-ABI reproduce them --
the one parity check that needs nothing but the repository."""
import glob
import json
import os

import numpy as np
import pytest

from gfx_imagecompress_b200 import synth
from oracle.ref import BC1, BC4, BC5, BC6H, BC7, BC7_RG, default_opts

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
# oracle.ref codec numbering -> b200ic_codec
ENGINE_CODEC = {BC1: 1, BC4: 4, BC5: 5, BC6H: 6, BC7: 7, BC7_RG: 8}


def load(path):
    z = np.load(path)
    px = z["pixels"]
    if bool(z["is_half"]):
        px = px.view(np.float16)
    return dict(name=os.path.basename(path)[:-4], pixels=np.ascontiguousarray(px), fmt=int(z["fmt"]), codec=int(z["codec"]),
                blocks=z["blocks"], opts=json.loads(str(z["opts"])))


def gather_rgba_f32(px: np.ndarray) -> np.ndarray:
    """Replicate-edge 4x4 gather exactly as the compat shim + src/block_utils.cpp:7-41: (nblocks, 64) float32 RGBA,
    u8 -> x / 255.0f, missing channels g=b=0, a=1."""
    h, w, c = px.shape
    by, bx = (h + 3) // 4, (w + 3) // 4
    ys = np.minimum(np.arange(by * 4), h - 1)
    xs = np.minimum(np.arange(bx * 4), w - 1)
    p = px[ys][:, xs]
    f = p.astype(np.float32) / np.float32(255) if px.dtype == np.uint8 else p.astype(np.float32)
    full = np.zeros((by * 4, bx * 4, 4), np.float32)
    full[..., 3] = 1.0
    full[..., :c] = f
    return np.ascontiguousarray(full.reshape(by, 4, bx, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64))


def test_fixtures_present():
    assert len(GOLDEN) >= 16


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_reference_reproduces_golden(ref, path):
    g = load(path)
    got = ref.encode(g["codec"], g["pixels"], g["fmt"], opts=default_opts(**g["opts"]) if g["opts"] else None)
    assert np.array_equal(got, g["blocks"])


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cpu_restatements_reproduce_golden(restated, path):
    """oracle/restate_bc4.c for BC4/BC5; the g++ builds of csrc/*_core.cuh (tests/hostbuild) for the others."""
    import hostbuild
    g = load(path)
    px, o = g["pixels"], g["opts"]
    if g["codec"] in (BC4, BC5):
        got = restated.bc4(px) if g["codec"] == BC4 else restated.bc5(px)
    else:
        L = hostbuild.load()
        fb = gather_rgba_f32(px)
        if g["codec"] == BC1:
            got = hostbuild.bc1_blocks(L, fb, o.get("bc1_alpha_threshold", 128) / 255.0, o.get("amd_refinement_steps", 1))
        elif g["codec"] == BC7:
            got, _ = hostbuild.bc7amd_blocks(L, fb, o.get("amd_mode_mask", 0xFF), u8_path=px.dtype == np.uint8)
        elif g["codec"] == BC6H:
            got = hostbuild.bc6h_blocks(L, fb)
        else:
            u8 = (fb * np.float32(255) + np.float32(0.5)).astype(np.uint8).reshape(-1, 16, 4)
            got = hostbuild.bc7rg_blocks(L, np.ascontiguousarray(u8).view(np.uint32).reshape(-1, 16),
                                         bool(o.get("rg_perceptual", 1)), bool(o.get("rg_fast", 0)))
    bad = np.flatnonzero((got != g["blocks"]).any(axis=1))
    assert bad.size == 0, f"{g['name']}: {bad.size}/{len(got)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_reproduces_golden(engine, path):
    g = load(path)
    o = g["opts"]
    kw = {}
    if "bc1_alpha_threshold" in o:
        kw["bc1_alpha_threshold"] = o["bc1_alpha_threshold"] / 255.0
    for k in ("amd_refinement_steps", "amd_mode_mask", "rg_perceptual", "rg_fast"):
        if k in o:
            kw[k] = o[k]
    got = engine.encode_host(ENGINE_CODEC[g["codec"]], g["pixels"], g["fmt"], engine.Opts.default(**kw) if kw else None)
    bad = np.flatnonzero((got != g["blocks"]).any(axis=1))
    assert bad.size == 0, f"{g['name']}: {bad.size}/{len(got)} blocks differ, first {bad[:5]}"
