"""Multi-GPU encode of images / texture batches: one process per GPU, block-row shards, no data-path collective.

Every 4x4 block is independent, so rank r encodes the block-rows `b200ic_plan_shards` deals to it and the
only exchange is the final gather of the 8/16-byte blocks.  torch.distributed is the plumbing (NCCL over NVLink on the
GPUs, gloo in the CPU tests); the encode itself is the CUDA library.

The reference has no equivalent (single-threaded, one image per call: src/amd_bc7_compressor.cpp:25-80); a caller of the
reference that loops over textures and mip levels maps onto `encode_batch_sharded`.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from . import api


def blocks_dims(w: int, h: int):
    return (w + 3) // 4, (h + 3) // 4


def shard_bytes(dims, shards, bb: int) -> int:
    return sum(blocks_dims(*dims[i])[0] * (r1 - r0) * bb for i, r0, r1 in shards)


def mip_chain_dims(size: int):
    """(w, h) of a full mip chain down to 1x1 (BASELINE config[4]: 2048^2 + 11 mips)."""
    out = []
    s = size
    while True:
        out.append((s, s))
        if s == 1:
            break
        s = max(1, s // 2)
    return out


def encode_batch_sharded(codec: int, images: Sequence, fmt: int, rank: int, world: int, group=None, opts=None,
                         chunk_rows: int = 0, gather: bool = True,
                         encode_rows: Callable | None = None):
    """Encodes `images` (torch tensors (H, W, C) in `fmt`; every rank passes the same list, only its shards are read)
    and returns one (nblocks, blockBytes) uint8 tensor per image.  With gather=True the tensors are complete on every
    rank (all_gather of the shard bytes, padded to the largest rank); otherwise only this rank's block-rows are filled.

    `encode_rows(image_index, row0, row1) -> uint8 tensor (rows*blocksX, blockBytes)` replaces the CUDA call; it exists
    so that the CPU tests can exercise the planning / packing / gather logic over gloo.  The default is the CUDA
    library (b200ic_encode_batch_device), which raises without a GPU."""
    import torch
    import torch.distributed as dist

    bb = api.BLOCK_BYTES[codec]
    dims = [(int(t.shape[1]), int(t.shape[0])) for t in images]
    plans = [api.plan_shards(dims, world, r, chunk_rows) for r in range(world)]
    mine = plans[rank]
    dev = images[0].device if len(images) else torch.device("cpu")
    outs = [torch.zeros((blocks_dims(*d)[0] * blocks_dims(*d)[1], bb), dtype=torch.uint8, device=dev) for d in dims]
    if encode_rows is None:
        api.encode_batch_device(codec, list(images), fmt, outs=outs, shards=mine, opts=opts)
    else:
        for i, r0, r1 in mine:
            bx = blocks_dims(*dims[i])[0]
            outs[i][r0 * bx:r1 * bx] = encode_rows(i, r0, r1)
    if not gather or world == 1:
        return outs
    sizes = [shard_bytes(dims, p, bb) for p in plans]
    cap = max(sizes)
    flat = torch.zeros(cap, dtype=torch.uint8, device=dev)
    pos = 0
    for i, r0, r1 in mine:
        bx = blocks_dims(*dims[i])[0]
        n = bx * (r1 - r0) * bb
        flat[pos:pos + n] = outs[i][r0 * bx:r1 * bx].reshape(-1)
        pos += n
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(gathered, flat, group=group)
    else:
        parts = [torch.empty(cap, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, flat, group=group)
        gathered = torch.cat(parts)
    for r in range(world):
        if r == rank:
            continue
        pos = r * cap
        for i, r0, r1 in plans[r]:
            bx = blocks_dims(*dims[i])[0]
            n = bx * (r1 - r0) * bb
            outs[i][r0 * bx:r1 * bx] = gathered[pos:pos + n].reshape(-1, bb)
            pos += n
    return outs


def box_mips(top):
    """Box-filtered mip chain of an (H, W, 4) uint8 CUDA tensor down to 1x1 (b200ic_box_mip_rgba8_device).  Input preparation
    for the batch workload; the reference leaves mip generation to its callers."""
    return api.box_mip_chain(top.contiguous())
