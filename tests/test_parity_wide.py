"""Parity where round 1 was thin ("What's weak" 1): the headline size, the mip tail, the batch path against
the ORACLE (not against the GPU's own host path), NPOT images under the same gate, and the format-pick function.

Gate for the float-search encoders (AMD BC7, BC6H): the blocks are byte-identical to the compiled reference, or the
differing blocks are listed and the decoded PSNR of the image stays within 0.02 dB of the reference's
(BASELINE.json north_star).  Everything else is bit-exact."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import sharded, synth
from oracle import metrics
from oracle.ref import BC1, BC4, BC5, BC6H, BC7, BC7_RG

PSNR_TOL_DB = 0.02
REF_OF = {1: BC1, 4: BC4, 5: BC5, 6: BC6H, 7: BC7, 8: BC7_RG}


def same_or_within_tolerance(name, got, want, px, codec):
    """Identical bytes, or: report which blocks differ and accept only a decoded-image PSNR within the stated tolerance."""
    bad = np.flatnonzero((got != want).any(axis=1))
    if bad.size == 0:
        return 0
    assert codec in (6, 7), f"{name}: {bad.size} blocks differ in a bit-exact codec, first {bad[:8]}"
    if codec == 7:
        rgba = px if px.shape[2] == 4 else np.concatenate([px, np.full(px.shape[:2] + (1,), 255, np.uint8)], 2)
        d = metrics.psnr_bc7(got, rgba) - metrics.psnr_bc7(want, rgba)
    else:
        h = px.view(np.float16)
        d = metrics.psnr_bc6h(got, h) - metrics.psnr_bc6h(want, h)
    assert abs(d) <= PSNR_TOL_DB, f"{name}: blocks {bad[:8]} (+{max(0, bad.size - 8)} more) differ and PSNR moves by {d:+.4f} dB"
    return bad.size


def mip_tail(top):
    """Box-filtered chain of an (H, W, C) uint8 image down to 1x1, every level filtered from the previous (quantised)
    level with floor(x + 0.5) rounding (numpy twin of b200ic_box_mip_rgba8_device)."""
    chain = [np.ascontiguousarray(top)]
    while chain[-1].shape[0] > 1 or chain[-1].shape[1] > 1:
        cur = chain[-1].astype(np.float32)
        h, w = cur.shape[:2]
        if h > 1:
            cur = (cur[0:h // 2 * 2:2] + cur[1:h // 2 * 2:2]) * 0.5
        if w > 1:
            cur = (cur[:, 0:w // 2 * 2:2] + cur[:, 1:w // 2 * 2:2]) * 0.5
        chain.append(np.ascontiguousarray(np.floor(cur + 0.5).clip(0, 255).astype(np.uint8)))
    return chain


@pytest.mark.gpu
@pytest.mark.parametrize("codec", [1, 7, 8], ids=["bc1", "bc7_amd", "bc7_rg"])
def test_mip_tail_levels_match_reference(engine, ref, codec):
    """Every level of a 64^2 chain -- 64, 32, 16, 8 and the sub-block levels 4x4 (exact), 2x2, 1x1 (replicate-edge
    gather, src/block_utils.cpp:19,22) -- as its own image, exactly how a caller of the reference compresses a chain."""
    for lvl in mip_tail(synth.rgba8_gradnoise(64, 64, 21, "lefthalf")):
        got = engine.encode_host(codec, lvl, synth.FMT_RGBA8)
        want = ref.encode(REF_OF[codec], lvl, synth.FMT_RGBA8)
        same_or_within_tolerance(f"level {lvl.shape[1]}x{lvl.shape[0]}", got, want, lvl, codec)
    odd = synth.rgba8_gradnoise(7, 3, 5, "ramp")  # ragged in both directions
    same_or_within_tolerance("7x3", engine.encode_host(codec, odd, synth.FMT_RGBA8), ref.encode(REF_OF[codec], odd, synth.FMT_RGBA8), odd, codec)


@pytest.mark.gpu
def test_mip_tail_bc6h(engine, ref):
    top = synth.hdr_rgba16f(32, 32, 4).view(np.float16).astype(np.float32)
    cur = top
    while True:
        lvl = np.ascontiguousarray(cur.astype(np.float16).view(np.uint16))
        got = engine.encode_host(6, lvl, synth.FMT_RGBA16UF)
        want = ref.encode(BC6H, lvl, synth.FMT_RGBA16UF)
        same_or_within_tolerance(f"hdr level {cur.shape[1]}", got, want, lvl, 6)
        if cur.shape[0] == 1:
            break
        cur = (cur[0::2, 0::2] + cur[1::2, 0::2] + cur[0::2, 1::2] + cur[1::2, 1::2]) * 0.25
        cur = cur + np.float32(1e-3) * np.arange(cur.shape[1], dtype=np.float32)[None, :, None]  # (no exactly flat blocks: reference UB)


@pytest.mark.gpu
def test_batch_path_matches_the_oracle(engine, ref):
    """b200ic_encode_batch_device over a full mip chain (device-resident, forked streams) against the compiled
    reference level by level -- not against the GPU's own per-image path."""
    import torch
    dev = torch.device("cuda", 0)
    top = synth.rgba8_gradnoise(128, 128, 7, "lefthalf")
    chain = sharded.box_mips(torch.from_numpy(top).to(dev))
    assert tuple(chain[-1].shape[:2]) == (1, 1)
    for codec in (8, 1, 7):
        outs = engine.encode_batch_device(codec, chain, synth.FMT_RGBA8)
        torch.cuda.synchronize()
        for t, o in zip(chain, outs):
            px = t.cpu().numpy()
            same_or_within_tolerance(f"codec {codec} level {px.shape[1]}", o.cpu().numpy(), ref.encode(REF_OF[codec], px, synth.FMT_RGBA8), px, codec)


@pytest.mark.gpu
def test_amd_images_identical_or_within_tolerance_including_npot(engine, ref):
    report = {}
    for name, px, fmt in cases.rgba_cases(small=True):
        got = engine.encode_host(7, px, fmt)
        want = ref.encode(BC7, px, fmt)
        if px.shape[0] % 4 or px.shape[1] % 4:  # PSNR on the padded image: pad by edge replication like the gather does
            ph, pw = (-px.shape[0]) % 4, (-px.shape[1]) % 4
            px = np.ascontiguousarray(np.pad(px, ((0, ph), (0, pw), (0, 0)), mode="edge"))
        report[name] = same_or_within_tolerance(name, got, want, px, 7)
    print("differing blocks per image:", report)


@pytest.mark.gpu
@pytest.mark.parametrize("codec", [7, 8], ids=["bc7_amd", "bc7_rg"])
def test_headline_size_64_block_rows(engine, ref, codec):
    """BASELINE config[2] at its full size: the 8192^2 image encoded whole on the GPU, 64 evenly spaced block-rows
    (131 072 blocks) against the compiled reference."""
    size = 8192
    px = synth.rgba8_gradnoise(size, size, 3, "lefthalf")
    got = engine.encode_host(codec, px, synth.FMT_RGBA8).reshape(size // 4, size // 4, 16)
    rows = [int(i * (size // 4) / 64) for i in range(64)]
    strip = np.ascontiguousarray(np.concatenate([px[4 * r:4 * r + 4] for r in rows], axis=0))
    want = ref.encode(REF_OF[codec], strip, synth.FMT_RGBA8)
    same_or_within_tolerance(f"{size}^2 rows", got[rows].reshape(-1, 16), want, strip, codec)


def test_pick_compression_type_matches_reference(ref):
    """ImageCompress_PickCompressionType over every flag combination x source format (src/imagecompress.cpp:52-116).
    Host-only logic: runs without a GPU."""
    import ctypes as C
    import gfx_imagecompress_b200 as g
    L = g.load_library()
    ref.lib.ImageCompress_PickCompressionType.argtypes = [C.c_int, C.c_void_p]
    ref.lib.ImageCompress_PickCompressionType.restype = C.c_int
    L.ImageCompress_PickCompressionType.restype = C.c_int
    fmts = [(synth.FMT_R8, 1, np.uint8), (synth.FMT_RG8, 2, np.uint8), (synth.FMT_RGB8, 3, np.uint8), (synth.FMT_RGBA8, 4, np.uint8),
            (8, 4, np.uint8), (synth.FMT_RGBA16F, 4, np.uint16), (synth.FMT_RGBA32F, 4, np.float32)]
    for fmt, ch, dt in fmts:
        img = g.Image(np.zeros((4, 4, ch), dt), fmt)
        for flags in range(16):
            want = ref.lib.ImageCompress_PickCompressionType(flags, img.ptr)
            assert g.ImageCompress_PickCompressionType(flags, img) == want, (fmt, flags)
