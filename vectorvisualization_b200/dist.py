"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the gathers.

Sort-first: the image is cut into 16x16-pixel blocks; block (bx, by) has the id by*nbx + (bx + skew*by) % nbx and
belongs to rank id % world.  The per-row rotation scatters one rank's blocks over x and y, which balances the ~20-70 %
box coverage and the chord-length variation without a cost model (plain row-major ids gave whole block columns to a
rank whenever nbx % world == 0: 9 % imbalance at 8 GPUs on cfg3, 0.2 % with the rotation); the vector field, noise,
scalar volume and tables are replicated on every GPU.  Each rank renders its blocks into a compact block-major tile buffer; one
all_gather of those buffers (a few MiB per frame) and one un-block kernel give every rank the frame.
LIC-volume mode: output z-slabs per rank (input replicated, so no halo exchange is needed -- SURVEY 8(e)), one
all_gather of the slabs.
"""
import numpy as np


class _CAI:
    """minimal __cuda_array_interface__ carrier so torch can view a raw device pointer without a copy"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def device_tensor(ptr, shape, dtype):
    import torch
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CAI(ptr, shape, typestr), device="cuda")


# ---- host-side partition bookkeeping (mirrors vv_renderer.cu: ensure_frame / unblock_kernel) ----

def block_grid(width, height, block=16):
    return (width + block - 1) // block, (height + block - 1) // block


def block_skew(world):
    """per-row rotation of the block ids (vv_device.cuh: block_skew_for)"""
    return 0 if world <= 1 else 2 * ((382 * world // 1000) // 2) + 1


def block_id(nbx, skew, bx, by, world=1, unit=1):
    """id of block (bx, by): owner = id % world, local block index = id // world (vv_device.cuh: block_id / block_id_u).
    unit > 1: units of unit x unit blocks are dealt to the ranks instead of single blocks"""
    if unit <= 1:
        return by * nbx + (bx + skew * by) % nbx
    nux = (nbx + unit - 1) // unit
    uid = (by // unit) * nux + (bx // unit + skew * (by // unit)) % nux
    return ((uid // world) * unit * unit + (by % unit) * unit + bx % unit) * world + uid % world


def block_xy(nbx, skew, b, world=1, unit=1):
    if unit <= 1:
        by = b // nbx
        return (b % nbx - skew * by) % nbx, by
    nux = (nbx + unit - 1) // unit
    lb, sub = b // world, (b // world) % (unit * unit)
    uid = (lb // (unit * unit)) * world + b % world
    uy = uid // nux
    ux = (uid % nux - skew * uy) % nux
    return ux * unit + sub % unit, uy * unit + sub // unit


def num_block_ids(width, height, unit=1, block=16):
    """ids in use: whole units, so blocks of a border unit that lie outside the image are counted (they hold no pixel)"""
    nbx, nby = block_grid(width, height, block)
    return ((nbx + unit - 1) // unit) * ((nby + unit - 1) // unit) * unit * unit


def blocks_per_rank(width, height, world, block=16, unit=1):
    nunits = num_block_ids(width, height, unit, block) // (unit * unit)
    return (nunits + world - 1) // world * unit * unit


def local_blocks(width, height, rank, world, block=16, unit=1):
    """block ids owned by `rank`, in local order; block_xy() gives their position"""
    nunits = num_block_ids(width, height, unit, block) // (unit * unit)
    n_local = max(0, (nunits - rank + world - 1) // world) * unit * unit
    return [rank + lb * world for lb in range(n_local)]


def assemble_host(gathered, width, height, world, block=16, unit=1):
    """numpy reference of unblock_kernel: gathered [world][blocks_per_rank][block*block][C] -> [h][w][C]"""
    nbx, nby = block_grid(width, height, block)
    out = np.zeros((height, width, gathered.shape[-1]), dtype=gathered.dtype)
    sk = block_skew(world)
    for by in range(nby):
        for bx in range(nbx):
            b = block_id(nbx, sk, bx, by, world, unit)
            r, lb = b % world, b // world
            tile = gathered[r, lb].reshape(block, block, -1)
            y0, x0 = by * block, bx * block
            h, w = min(block, height - y0), min(block, width - x0)
            out[y0:y0 + h, x0:x0 + w] = tile[:h, :w]
    return out


def slab_range(depth, rank, world):
    """z-slab [z0,z1) of the LIC volume owned by `rank`: equal slabs, remainder to the first ranks"""
    base, rem = divmod(depth, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def render_distributed(renderer, group=None):
    """render this rank's blocks, all-gather the tile buffers over NCCL, assemble the frame on every rank.
    Returns the gathered tensor (kept alive by the caller until the frame has been read)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    st = torch.cuda.current_stream().cuda_stream   # kernels and NCCL on one stream: no cross-stream race
    if st == 0:
        raise RuntimeError("render_distributed needs a non-default torch stream (torch.cuda.set_stream(torch.cuda.Stream()))")
    renderer.setStream(st)
    renderer.render(True)
    ptr, bpr, _ = renderer.tileBuffer()
    local = device_tensor(ptr, (bpr * 256 * 4,), torch.float32)
    gathered = torch.empty((world * bpr * 256 * 4,), dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    renderer.assembleTiles(gathered.data_ptr(), world)
    return gathered


def connect_p2p(renderer, group=None):
    """peer-to-peer exchange instead of the NCCL gather: every rank allocates its gather buffer, the CUDA IPC handles go
    round once over the host channel, every rank maps every buffer.  Call after resize / setPartition; collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    st = torch.cuda.current_stream().cuda_stream
    if st == 0:
        raise RuntimeError("connect_p2p needs a non-default torch stream (torch.cuda.set_stream(torch.cuda.Stream()))")
    renderer.setStream(st)
    handle, _ = renderer.p2pExport()
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    renderer.p2pConnect(handles=handles)
    dist.barrier(group)          # every buffer is mapped everywhere before the first store


def render_distributed_p2p(renderer):
    """render this rank's blocks, store them into every rank's gather buffer over NVLink, wait for the peers' blocks,
    assemble the frame -- one call per frame on every rank, nothing but kernels on the stream"""
    renderer.p2pRender()


def disconnect_p2p(renderer, group=None):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    renderer.p2pStatus()
    dist.barrier(group)          # nobody still writes into a buffer that is about to be freed
    renderer.p2pDisconnect()


def update_lic_volume_distributed(renderer, depth, group=None):
    """each rank computes its z-slab of the LIC volume, then the slabs are all-gathered in place"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    z0, z1 = slab_range(depth, rank, world)
    st = torch.cuda.current_stream().cuda_stream
    if st == 0:
        raise RuntimeError("update_lic_volume_distributed needs a non-default torch stream")
    renderer.setStream(st)
    renderer.setLICVolumeSlab(z0, z1)
    renderer.updateLICVolume()
    ptr, dims = renderer.licVolumePtr()
    vol = device_tensor(ptr, (dims[2], dims[1] * dims[0]), torch.float32)
    if depth % world == 0:
        dist.all_gather_into_tensor(vol.view(-1), vol[z0:z1].reshape(-1).clone(), group=group)
    else:
        for r in range(world):
            a, b = slab_range(depth, r, world)
            dist.broadcast(vol[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return vol
