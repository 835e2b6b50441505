"""world_size-2 gloo test of the N > 1 host logic (sort-first block partition, tile gather layout, LIC-volume slabs).
The device kernels are not involved: each rank builds its block-major tile buffer from a known image with the same
bookkeeping the renderer uses, the buffers are all-gathered over gloo and assembled with the numpy mirror of the
un-block kernel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, q, unit=1):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from vectorvisualization_b200 import dist as vd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(7)
    full = rng.rand(h, w, 4).astype(np.float32)                      # the frame every rank would agree on
    nbx, nby = vd.block_grid(w, h)
    bpr = vd.blocks_per_rank(w, h, world, unit=unit)
    mine = np.zeros((bpr, 256, 4), np.float32)
    for lb, b in enumerate(vd.local_blocks(w, h, rank, world, unit=unit)):      # this rank "renders" only its own blocks
        bx, by = vd.block_xy(nbx, vd.block_skew(world), b, world, unit)
        tile = np.zeros((16, 16, 4), np.float32)
        y0, x0 = by * 16, bx * 16
        hh, ww = max(0, min(16, h - y0)), max(0, min(16, w - x0))    # blocks of a border unit may lie outside the image
        if hh and ww:
            tile[:hh, :ww] = full[y0:y0 + hh, x0:x0 + ww]
        mine[lb] = tile.reshape(256, 4)
    gathered = torch.empty((world * bpr * 256 * 4,), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(mine).reshape(-1))
    img = vd.assemble_host(gathered.numpy().reshape(world, bpr, 256, 4), w, h, world, unit=unit)
    ok = bool(np.array_equal(img, full))
    # LIC-volume slabs: every rank fills its z range, slabs are exchanged, result is complete
    depth = 13
    vol = torch.zeros((depth, 5))
    z0, z1 = vd.slab_range(depth, rank, world)
    vol[z0:z1] = torch.arange(z0, z1, dtype=torch.float32)[:, None]
    for r in range(world):
        a, b = vd.slab_range(depth, r, world)
        dist.broadcast(vol[a:b], src=r)
    ok = ok and bool(torch.equal(vol[:, 0], torch.arange(depth, dtype=torch.float32)))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("size,unit", [((90, 70), 1), ((64, 64), 1), ((33, 17), 1), ((90, 70), 2), ((200, 120), 4)])
def test_partition_gather_assemble_gloo(size, unit):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, size[0], size[1], q, unit)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_partition_bookkeeping():
    from vectorvisualization_b200 import dist as vd
    for (w, h) in ((1024, 1024), (90, 70), (16, 16), (1, 1)):
        nbx, nby = vd.block_grid(w, h)
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lb = vd.local_blocks(w, h, r, world)
                assert len(lb) <= vd.blocks_per_rank(w, h, world)
                seen += lb
            assert sorted(seen) == list(range(nbx * nby))
            # id <-> (bx, by) is a bijection (what unblock_kernel and ray_setup_kernel rely on)
            sk = vd.block_skew(world)
            pos = [vd.block_xy(nbx, sk, b) for b in range(nbx * nby)]
            assert sorted(pos) == sorted((x, y) for y in range(nby) for x in range(nbx))
            assert all(vd.block_id(nbx, sk, x, y) == b for b, (x, y) in enumerate(pos))
    # units of U x U blocks (VV_OPT_PARTITION_UNIT): ids stay a bijection onto the padded block grid, id % world owns whole units
    for (w, h) in ((1024, 1024), (90, 70), (200, 120), (16, 16)):
        nbx, nby = vd.block_grid(w, h)
        for unit in (2, 4):
            for world in (2, 3, 8):
                sk = vd.block_skew(world)
                nids = vd.num_block_ids(w, h, unit)
                seen = []
                for r in range(world):
                    lb = vd.local_blocks(w, h, r, world, unit=unit)
                    assert len(lb) <= vd.blocks_per_rank(w, h, world, unit=unit) and len(lb) % (unit * unit) == 0
                    seen += lb
                pos = {}
                for b in seen:
                    x, y = vd.block_xy(nbx, sk, b, world, unit)
                    assert (x, y) not in pos
                    pos[(x, y)] = b
                    assert vd.block_id(nbx, sk, x, y, world, unit) == b
                assert len(seen) == nids and all((x, y) in pos for y in range(nby) for x in range(nbx))
                # the blocks of a unit share their owner and are consecutive local blocks
                for (x, y), b in pos.items():
                    b0 = pos[(x // unit * unit, y // unit * unit)]
                    assert b % world == b0 % world and b // world - b0 // world == (y % unit) * unit + x % unit
    # the rotation scatters a rank's blocks over columns as well as rows even when nbx % world == 0
    nbx, nby = vd.block_grid(1024, 1024)
    for world in (2, 4, 8):
        for r in range(world):
            cols = {vd.block_xy(nbx, vd.block_skew(world), b)[0] for b in vd.local_blocks(1024, 1024, r, world)}
            assert len(cols) == nbx
    for depth in (1024, 13, 7):
        for world in (1, 2, 4, 8):
            z = [vd.slab_range(depth, r, world) for r in range(world)]
            assert z[0][0] == 0 and z[-1][1] == depth and all(z[i][1] == z[i + 1][0] for i in range(world - 1))
