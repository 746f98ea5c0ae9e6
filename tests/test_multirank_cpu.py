"""world_size-2 (gloo, CPU) test of the host-side sharding logic used by bench.py for
N > 1: every rank owns the interleaved tiles rt_tile_count() says, fills its packed buffer in
the device's slot order, rank 0 gathers (the NCCL gather's CPU twin) and de-interleaves with
the same index arithmetic as the k_unpack kernel; the result must be the full frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(__file__))
from conftest import load_package  # noqa: E402

TILE = 32


def pixel_value(px, py):
    return (px * 7 + py * 13) % 251


def owned_tiles(w, h, rank, world):
    tx_n, ty_n = (w + TILE - 1) // TILE, (h + TILE - 1) // TILE
    return [(tx, ty) for ty in range(ty_n) for tx in range(tx_n) if (tx + ty) % world == rank]


def pack_rank(w, h, rank, world, max_tiles):
    """Slot order of the device framebuffer: tile-major, one warp = an 8x4 pixel block."""
    buf = np.zeros((max_tiles, TILE * TILE), np.uint8)
    for lt, (tx, ty) in enumerate(owned_tiles(w, h, rank, world)):
        for j in range(TILE * TILE):
            wv, l = j >> 5, j & 31
            px, py = tx * TILE + (wv & 3) * 8 + (l & 7), ty * TILE + (wv >> 2) * 4 + (l >> 3)
            if px < w and py < h:
                buf[lt, j] = pixel_value(px, py)
    return buf


def unpack(w, h, world, max_tiles, gathered):
    frame = np.zeros((h, w), np.uint8)
    starts = {}
    for r in range(world):
        for lt, t in enumerate(owned_tiles(w, h, r, world)):
            starts[t] = (r, lt)
    for py in range(h):
        for px in range(w):
            r, lt = starts[(px // TILE, py // TILE)]
            x, y = px % TILE, py % TILE
            j = ((y // 4) * 4 + x // 8) * 32 + (y % 4) * 8 + x % 8
            frame[py, px] = gathered[r][lt, j]
    return frame


def worker(rank, world, port, w, h, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    own, max_tiles, total = pkg.tile_counts(pkg.make_params(w, h, 5, tile_rank=rank, tile_world=world))
    assert own == len(owned_tiles(w, h, rank, world))
    mine = torch.from_numpy(pack_rank(w, h, rank, world, max_tiles))
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    counts = torch.tensor([own], dtype=torch.int64)
    dist.all_reduce(counts)
    assert int(counts.item()) == total
    if rank == 0:
        frame = unpack(w, h, world, max_tiles, [g.numpy() for g in gathered])
        yy, xx = np.mgrid[0:h, 0:w]
        ok.value = int(np.array_equal(frame, pixel_value(xx, yy).astype(np.uint8)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [(100, 70), (33, 65)])
def test_two_rank_tile_gather(size):
    w, h = size
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    procs = [ctx.Process(target=worker, args=(r, 2, port, w, h, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ok.value == 1
