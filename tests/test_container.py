"""The steps either side of the encode path: device-side mip generation and the DDS writer."""
import os
import struct

import numpy as np
import pytest

from gfx_imagecompress_b200 import synth


def test_dds_writer_header_and_payload(tmp_path):
    """Host-only: runs without a GPU."""
    import gfx_imagecompress_b200 as g
    g.load_library()
    w, h = 20, 12
    levels = []
    lw, lh = w, h
    rng = np.random.default_rng(3)
    while True:
        levels.append(rng.integers(0, 256, (((lw + 3) // 4) * ((lh + 3) // 4), 16), dtype=np.uint8))
        if lw == 1 and lh == 1:
            break
        lw, lh = max(1, lw // 2), max(1, lh // 2)
    path = str(tmp_path / "t.dds")
    g.write_dds(path, g.BC7_AMD, w, h, levels, srgb=True)
    raw = open(path, "rb").read()
    hdr = struct.unpack("<32I", raw[:128])
    assert raw[:4] == b"DDS " and hdr[1] == 124 and hdr[3] == h and hdr[4] == w and hdr[7] == len(levels)
    assert hdr[5] == levels[0].nbytes and hdr[21] == struct.unpack("<I", b"DX10")[0]
    dx10 = struct.unpack("<5I", raw[128:148])
    assert dx10[0] == 99 and dx10[1] == 3 and dx10[3] == 1  # DXGI_FORMAT_BC7_UNORM_SRGB, TEXTURE2D, one array slice
    assert raw[148:] == b"".join(l.tobytes() for l in levels)
    g.write_dds(path, g.BC1, 8, 8, [np.zeros((4, 8), np.uint8)])
    assert os.path.getsize(path) == 148 + 32 and struct.unpack("<I", open(path, "rb").read()[128:132])[0] == 71
    with pytest.raises(g.B200Error):
        g.write_dds(str(tmp_path / "no" / "dir.dds"), g.BC1, 8, 8, [np.zeros((4, 8), np.uint8)])


@pytest.mark.gpu
def test_device_mip_chain_matches_box_filter(engine):
    import torch
    from test_parity_wide import mip_tail
    for w, h in ((64, 64), (40, 24), (5, 9), (1, 7)):
        top = synth.rgba8_gradnoise(w, h, 11, "ramp")
        got = engine.box_mip_chain(torch.from_numpy(top).cuda())
        want = mip_tail(top)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert np.array_equal(a.cpu().numpy(), b), (w, h, b.shape)


@pytest.mark.gpu
def test_chain_to_dds_roundtrip(engine, ref, tmp_path):
    """mip chain on the device -> batch encode -> DDS file; every level in the file is the reference's bytes."""
    import torch
    from oracle.ref import BC7_RG
    top = synth.rgba8_gradnoise(64, 32, 5, "lefthalf")
    chain = engine.box_mip_chain(torch.from_numpy(top).cuda())
    outs = engine.encode_batch_device(engine.BC7_RG, chain, synth.FMT_RGBA8)
    torch.cuda.synchronize()
    path = str(tmp_path / "chain.dds")
    engine.write_dds(path, engine.BC7_RG, 64, 32, [o.cpu().numpy() for o in outs])
    raw = open(path, "rb").read()[148:]
    pos = 0
    for lvl in chain:
        want = ref.encode(BC7_RG, lvl.cpu().numpy(), synth.FMT_RGBA8).tobytes()
        assert raw[pos:pos + len(want)] == want
        pos += len(want)
    assert pos == len(raw)
