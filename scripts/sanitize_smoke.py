"""tiny frames of every technique for compute-sanitizer (memcheck on the smallest config):
    compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs, fields as F

n, size = 12, 24
for name, mk in (("cfg1 ray-cast", lambda: configs.cfg1(n=n, size=size)),
                 ("cfg3 gradient ray-cast, clip plane", lambda: configs.cfg3(n=n, size=size, camera=F.CAMERA_CLOSE)),
                 ("cfg2 slicing", lambda: configs.cfg2(n=n, size=size))):
    s = mk()
    if "clip" in name:
        s.clip_planes = ((0.0, 0.0, -1.0, 0.1),)
    if "slicing" in name:
        s.technique, s.tf_mode, s.gate_mode = vv.VOLIC_SLICING, vv.TF_A, vv.GATE_TF_ALPHA
        s.fbo = 1
    r = vv.Renderer(0)
    configs.apply_scene(r, s)
    r.render(True)
    r.readRGBA8()
    print("%s: %d ray samples, %d launches" % (name, r.lastRaySamples(), r.lastLaunchCount()), flush=True)
    if "cfg1" in name:
        r.setTechnique(vv.VOLIC_LICVOLUME)
        r.render(True)
        print("  LIC volume + volume ray-cast: %d ray samples" % r.lastRaySamples(), flush=True)
    r.close()

# three partitioned handles in one process exchanging tiles through peer stores + system-scope arrival counters (vv_p2p_*):
# the atomic work queue and the exchange are where racecheck / synccheck matter
s = configs.cfg3(n=n, size=size)
s.width, s.height = 40, 28
hs = []
for rank in range(3):
    r = vv.Renderer(0)
    configs.apply_scene(r, s)
    r.setPartition(rank, 3)
    r.render(True)
    r.synchronize()
    hs.append(r)
bases = [r.p2pExport()[1] for r in hs]
for r in hs:
    r.p2pConnect(local_bases=bases)
for frame in range(2):
    for r in hs:
        r.p2pRender()
    for r in hs:
        r.p2pStatus()
        r.readRGBA8()
print("p2p exchange, 3 handles x 2 frames: %d ray samples on rank 0" % hs[0].lastRaySamples(), flush=True)
for r in hs:
    r.p2pDisconnect()
    r.close()
