"""render a few frames of one config (for ncu): python scripts/profile_frame.py cfg3 [frames] [key=value options]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs, fields as F

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
opts = dict(a.split("=") for a in sys.argv[3:])
kw = {}
if "n" in opts:
    kw["n"] = int(opts["n"])
if "size" in opts:
    kw["size"] = int(opts["size"])
if opts.get("camera") == "close":
    kw["camera"] = F.CAMERA_CLOSE
s = getattr(configs, name)(**kw)
r = vv.Renderer(0)
if "mode" in opts:
    r.setOption(vv.OPT_RAYCAST_MODE, int(opts["mode"]))
if "layout" in opts:
    r.setOption(vv.OPT_FIELD_LAYOUT, int(opts["layout"]))
if "xf" in opts:
    r.setOption(vv.OPT_WALK_FAST_PATHS, int(opts["xf"]))
if "depthmajor" in opts:
    r.setOption(vv.OPT_DEPTH_MAJOR, int(opts["depthmajor"]))
if "band" in opts:
    r.setOption(vv.OPT_BAND_ROWS, int(opts["band"]))
if "noiselayout" in opts:
    r.setOption(vv.OPT_NOISE_LAYOUT, int(opts["noiselayout"]))
if "unit" in opts:
    r.setOption(vv.OPT_PARTITION_UNIT, int(opts["unit"]))
if "ctas" in opts:
    r.setOption(vv.OPT_LIC_CTAS_PER_SM, int(opts["ctas"]))
configs.apply_scene(r, s)
if "part" in opts:   # one rank's share of a sort-first partition, e.g. part=0/8
    pr, pw = (int(v) for v in opts["part"].split("/"))
    r.setPartition(pr, pw)
if "loop" in opts:   # back-to-back frames, one synchronize at the end (the bench.py `value` loop without NCCL)
    r.render(True)
    r.synchronize()
    k = int(opts["loop"])
    t = time.perf_counter()
    for i in range(k):
        r.render(True)
    r.synchronize()
    dt = (time.perf_counter() - t) / k
    print("%s loop of %d frames: %.3f ms/frame wall, last kernel %.3f ms, %d ray samples" % (name, k, dt * 1e3, r.lastKernelMs(), r.lastRaySamples()), flush=True)
for i in range(frames):
    t = time.perf_counter()
    r.render(True)
    r.synchronize()
    dt = time.perf_counter() - t
    n = r.lastRaySamples()
    print("%s frame %d: %.3f ms wall, kernel %.3f ms, %d ray samples, %.3f G samples/s (kernel), launches %d"
          % (name, i, dt * 1e3, r.lastKernelMs(), n, n / max(r.lastKernelMs(), 1e-9) / 1e6, r.lastLaunchCount()), flush=True)
