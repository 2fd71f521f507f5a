"""The host-buffer path of the C-ABI (b200ic_encode_host / b200ic_encode_host_sharded, csrc/api.cu): pageable and pinned
caller buffers, chunk pipeline, per-block-row progress with the reference's percentages and cancellation
(reference src/amd_bc1_compressor.cpp:64-68), block-row sharding over the GPUs of one process
(the loop of src/amd_bc7_compressor.cpp:48-77 split by block-row)."""
import numpy as np
import pytest

from gfx_imagecompress_b200 import synth


@pytest.mark.gpu
def test_pageable_and_pinned_buffers_agree(engine):
    import torch
    px = synth.height_rg8(1024, 4096, 2)  # 8 MiB: several chunks of the ring
    a = engine.encode_host(engine.BC5, px, synth.FMT_RG8)  # numpy = pageable: staged through pinned memory
    pin = torch.from_numpy(px).pin_memory()
    out = torch.empty((a.shape[0], 16), dtype=torch.uint8).pin_memory()
    engine.encode_host(engine.BC5, pin.numpy(), synth.FMT_RG8, out=out.numpy())
    assert np.array_equal(a, out.numpy())
    dev = engine.encode_device(engine.BC5, torch.from_numpy(px).cuda(), synth.FMT_RG8, 1024, 4096).cpu().numpy()
    assert np.array_equal(a, dev)


@pytest.mark.gpu
def test_progress_matches_reference_percentages_and_cancels(engine):
    px = synth.rgba8_gradnoise(64, 132, 3, "opaque")  # 33 block-rows
    seen = []
    engine.encode_host(engine.BC1, px, synth.FMT_RGBA8, progress=lambda pct: seen.append(pct) and False)
    bx, by = 16, 33
    want = [np.float32(100.0) * np.float32(y * bx) / np.float32(bx * by) for y in range(by)]
    assert len(seen) == by and np.allclose(seen, want, rtol=1e-6)
    calls = []
    out = engine.encode_host(engine.BC1, px, synth.FMT_RGBA8, progress=lambda pct: calls.append(pct) or len(calls) >= 3)
    assert out is None and len(calls) == 3  # cancelled: the reference returns nullptr
    img = engine.Image(px, synth.FMT_RGBA8)
    assert engine.Image_CompressAMDBC1(img, progress=lambda pct: True) is None


@pytest.mark.gpu
def test_slices_and_ragged_chunks(engine):
    px = np.stack([synth.rgba8_gradnoise(36, 20, s, "ramp") for s in range(3)])  # 3 slices, NPOT
    whole = engine.encode_host(engine.BC7_RG, px, synth.FMT_RGBA8)
    per = np.concatenate([engine.encode_host(engine.BC7_RG, np.ascontiguousarray(px[s]), synth.FMT_RGBA8) for s in range(3)])
    assert np.array_equal(whole, per)


@pytest.mark.gpu
def test_sharded_over_devices_matches_single_device(engine):
    """In-process block-row sharding (b200ic_encode_host_sharded). With one visible GPU the call degenerates to the
    single-device path; with more (gpurun --gpus N) every device encodes its range straight into the caller's buffer."""
    n = engine.device_count()
    px = synth.rgba8_gradnoise(512, 512, 3, "lefthalf")
    one = engine.encode_host(engine.BC7_AMD, px, synth.FMT_RGBA8)
    many = engine.encode_host(engine.BC7_AMD, px, synth.FMT_RGBA8, devices=0)
    assert np.array_equal(one, many), f"{n} devices"
    if n > 1:
        two = engine.encode_host(engine.BC1, px, synth.FMT_RGBA8, devices=2)
        assert np.array_equal(two, engine.encode_host(engine.BC1, px, synth.FMT_RGBA8))
    engine.init(0)


@pytest.mark.gpu
def test_misaligned_pointers_are_rejected_not_faulted(engine):
    import torch
    buf = torch.zeros(64 * 64 * 4 + 16, dtype=torch.uint8, device="cuda")
    src = buf[1:1 + 64 * 64 * 4]
    with pytest.raises(engine.B200Error):
        engine.encode_device(engine.BC1, src, synth.FMT_RGBA8, 64, 64)
    ok = buf[:64 * 64 * 4]
    out = torch.empty(16 * 16 * 8 + 8, dtype=torch.uint8, device="cuda")
    with pytest.raises(engine.B200Error):
        engine.encode_device(engine.BC1, ok, synth.FMT_RGBA8, 64, 64, out=out[4:4 + 16 * 16 * 8])
    engine.encode_device(engine.BC1, ok, synth.FMT_RGBA8, 64, 64)
    torch.cuda.synchronize()
