"""Multi-GPU path: block-row shards, no data-path collective, final gather of the blocks.

CPU: the planner's invariants, and a world_size-2 gloo run of the planning / packing / gather logic with the plain-C
BC5 restatement (oracle, test infrastructure) standing in for the CUDA call.  GPU: batch + shards through the C-ABI."""
import os
import socket
import sys

import numpy as np
import pytest

import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import sharded, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_covers_every_block_row_once():
    dims = sharded.mip_chain_dims(2048) * 3 + [(257, 131), (3, 5), (1, 1)]
    for world in (1, 2, 3, 4, 8):
        seen = {}
        loads = []
        for r in range(world):
            load = 0
            for i, r0, r1 in g.plan_shards(dims, world, r):
                assert r0 < r1 <= (dims[i][1] + 3) // 4
                for row in range(r0, r1):
                    assert (i, row) not in seen, "block-row dealt twice"
                    seen[(i, row)] = r
                load += ((dims[i][0] + 3) // 4) * (r1 - r0)
            loads.append(load)
        assert len(seen) == sum((h + 3) // 4 for _, h in dims)
        assert max(loads) - min(loads) <= 512 * 64, f"world {world}: imbalance {loads}"  # one chunk of the top level


def test_plan_is_deterministic_and_merges_adjacent_rows():
    assert g.plan_shards([(64, 64)], 1, 0) == [(0, 0, 16)]
    a = g.plan_shards([(8192, 8192)], 8, 3, 16)
    assert a == g.plan_shards([(8192, 8192)], 8, 3, 16) and a == [(0, 768, 1024)]  # one contiguous run per rank
    assert g.plan_shards([(64, 64)], 2, 5) == []  # rank outside the world


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        import oracle
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        restated = oracle.Restated()
        top = synth.height_rg8(128, 128, 5)
        imgs = [np.ascontiguousarray(top[::1 << m, ::1 << m]) for m in range(8)] + [synth.height_rg8(37, 21, 6)]
        tens = [torch.from_numpy(i) for i in imgs]

        def encode_rows(i, r0, r1):  # what the CUDA library does for a shard: rows [4*r0, 4*r1) as their own image
            sub = np.ascontiguousarray(imgs[i][4 * r0:min(4 * r1, imgs[i].shape[0])])
            return torch.from_numpy(restated.bc5(sub))

        outs = sharded.encode_batch_sharded(g.BC5, tens, synth.FMT_RG8, rank, world, chunk_rows=4, encode_rows=encode_rows)
        ok = all(np.array_equal(o.numpy(), restated.bc5(i)) for o, i in zip(outs, imgs))
        part = sharded.encode_batch_sharded(g.BC5, tens, synth.FMT_RG8, rank, world, chunk_rows=4, encode_rows=encode_rows,
                                            gather=False)
        own = sum(int((o.numpy() != 0).any(axis=1).sum()) for o in part)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, own))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_encode_matches_whole_image(oracle_built, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] is True for r in res), res
    total = sum(r[2] for r in res)
    assert total > 0 and all(r[2] < total for r in res), res  # every rank encoded only a part


@pytest.mark.gpu
def test_batch_device_matches_per_image_encode(engine):
    import torch
    dev = torch.device("cuda", 0)
    top = torch.from_numpy(synth.rgba8_gradnoise(256, 256, 7, "lefthalf")).to(dev)
    chain = sharded.box_mips(top)
    assert [tuple(t.shape[:2]) for t in chain][-1] == (1, 1) and len(chain) == 9
    for codec in (engine.BC7_RG, engine.BC1, engine.BC5):
        outs = engine.encode_batch_device(codec, chain, synth.FMT_RGBA8)
        torch.cuda.synchronize()
        for t, o in zip(chain, outs):
            want = engine.encode_host(codec, t.cpu().numpy(), synth.FMT_RGBA8)
            assert np.array_equal(o.cpu().numpy(), want)
    # shards of two ranks together == whole images
    dims = [(t.shape[1], t.shape[0]) for t in chain]
    whole = engine.encode_batch_device(engine.BC7_RG, chain, synth.FMT_RGBA8)
    parts = [torch.zeros_like(o) for o in whole]
    for r in range(2):
        engine.encode_batch_device(engine.BC7_RG, chain, synth.FMT_RGBA8, outs=parts, shards=engine.plan_shards(dims, 2, r, 8))
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(whole, parts))
