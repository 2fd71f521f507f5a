"""GPU decoders (b200ic_decode_device, csrc/decode.cu) against the spec decoders of oracle/bcdec.c:
on encoder output of every codec, and on random bytes (every BC7 / BC6H mode incl. reserved encodings, both BC1 modes,
both BC4 ramps)."""
import numpy as np
import pytest

from gfx_imagecompress_b200 import synth
from oracle import metrics


def _dec(engine, codec, blocks, w, h, **kw):
    import torch
    return engine.decode_device(codec, torch.from_numpy(np.ascontiguousarray(blocks)).cuda(), w, h, **kw).cpu().numpy()


@pytest.mark.gpu
def test_decoders_match_oracle_on_encoder_output(engine):
    D = metrics.decoders()
    px = synth.rgba8_gradnoise(68, 36, 3, "lefthalf")  # NPOT in blocks' terms: 17 x 9 blocks
    for codec in (1, 2, 3, 7, 8):
        b = engine.encode_host(codec, px, synth.FMT_RGBA8)
        got = _dec(engine, codec, b, 68, 36)
        want = D.bc1(b, 68, 36) if codec == 1 else (D.bc23(b, 68, 36, codec == 2) if codec in (2, 3) else D.bc7(b, 68, 36))
        assert np.array_equal(got, want), codec
    rg = synth.height_rg8(70, 30, 2)
    for codec, nch in ((4, 1), (5, 2)):
        b = engine.encode_host(codec, rg, synth.FMT_RG8)
        assert np.array_equal(_dec(engine, codec, b, 70, 30), D.bc45(b, 70, 30, nch)), codec
    hdr = synth.hdr_rgba16f(32, 32, 4)
    b = engine.encode_host(6, hdr, synth.FMT_RGBA16UF)
    got = _dec(engine, 6, b, 32, 32)
    assert np.array_equal(got[..., :3].view(np.uint16), D.bc6h(b, 32, 32).view(np.uint16)) and (got[..., 3] == 1.0).all()


@pytest.mark.gpu
def test_decoders_match_oracle_on_random_bytes(engine):
    D = metrics.decoders()
    rng = np.random.default_rng(5)
    w = h = 128  # 1024 blocks
    b16 = rng.integers(0, 256, (1024, 16), dtype=np.uint8)
    b8 = rng.integers(0, 256, (1024, 8), dtype=np.uint8)
    b16[:8, 0] = 0  # reserved BC7 encodings decode to zero
    assert np.array_equal(_dec(engine, 7, b16, w, h), D.bc7(b16, w, h))
    assert np.array_equal(_dec(engine, 1, b8, w, h), D.bc1(b8, w, h))
    assert np.array_equal(_dec(engine, 3, b16, w, h), D.bc23(b16, w, h, False))
    assert np.array_equal(_dec(engine, 2, b16, w, h), D.bc23(b16, w, h, True))
    assert np.array_equal(_dec(engine, 4, b8, w, h), D.bc45(b8, w, h, 1))
    assert np.array_equal(_dec(engine, 5, b16, w, h), D.bc45(b16, w, h, 2))
    for sgn in (False, True):
        got = _dec(engine, 6, b16, w, h, is_signed=sgn)
        assert np.array_equal(got[..., :3].view(np.uint16), D.bc6h(b16, w, h, sgn).view(np.uint16)), sgn
