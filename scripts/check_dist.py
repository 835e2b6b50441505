"""torchrun --nproc-per-node N scripts/check_dist.py : distributed frame (sort-first + NCCL all_gather) == single-GPU frame,
and distributed LIC volume (z-slabs + all_gather) == single-GPU LIC volume, bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs
from vectorvisualization_b200.dist import (render_distributed, update_lic_volume_distributed, connect_p2p, render_distributed_p2p,
                                           disconnect_p2p)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_stream(torch.cuda.Stream())
ok = True
for mk in (lambda: configs.cfg3(n=64, size=200), lambda: configs.cfg1(n=64, size=250)):
    s = mk()
    s.width, s.height = s.width + 7, s.height - 9
    ref = vv.Renderer(local)
    configs.apply_scene(ref, s)
    ref.render(True)
    want = ref.readRGBA32F()
    total = ref.lastRaySamples()
    r = vv.Renderer(local)
    configs.apply_scene(r, s)
    r.setPartition(rank, world)
    keep = render_distributed(r)
    torch.cuda.synchronize()
    got = r.readRGBA32F()
    n = torch.tensor([r.lastRaySamples()], dtype=torch.int64, device="cuda")
    dist.all_reduce(n)
    same = bool(np.array_equal(got, want)) and int(n.item()) == total
    print("rank %d %s: frame identical %s, ray samples %d / %d" % (rank, s.name, same, int(n.item()), total), flush=True)
    ok = ok and same
    # the same frame through the peer-to-peer exchange (CUDA IPC + NVLink stores), three frames for both buffer parities
    connect_p2p(r)
    for it in range(3):
        render_distributed_p2p(r)
        torch.cuda.synchronize()
        r.p2pStatus()
        same = bool(np.array_equal(r.readRGBA32F(), want))
        print("rank %d %s: p2p frame %d identical %s" % (rank, s.name, it, same), flush=True)
        ok = ok and same
    disconnect_p2p(r)
# LIC volume slabs
s = configs.cfg2(n=48, size=64)
s.licvol_fp16 = 0
ref = vv.Renderer(local)
configs.apply_scene(ref, s)
ref.updateLICVolume()
want = ref.readLICVolume()
r = vv.Renderer(local)
configs.apply_scene(r, s)
vol = update_lic_volume_distributed(r, 48)
torch.cuda.synchronize()
got = r.readLICVolume()
same = bool(np.array_equal(got, want))
print("rank %d LIC volume identical %s" % (rank, same), flush=True)
ok = ok and same
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
