"""LIC-volume mode timing: python scripts/run_licvol.py <n> <image size> [cfg name] [field layout]  (precompute + volume ray-cast)"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs

n = int(sys.argv[1]); size = int(sys.argv[2]); name = sys.argv[3] if len(sys.argv) > 3 else "cfg5"
t = time.perf_counter()
s = getattr(configs, name)(n=n, size=size) if name in ("cfg5", "cfg4") else getattr(configs, name)(n=n, size=size)
s.technique = vv.VOLIC_LICVOLUME
print("generated %s n=%d in %.1f s" % (name, n, time.perf_counter() - t), flush=True)
r = vv.Renderer(0)
if len(sys.argv) > 4:
    r.setOption(vv.OPT_FIELD_LAYOUT, int(sys.argv[4]))      # 1 x-pair (default), 2 xy-quad
t = time.perf_counter()
configs.apply_scene(r, s)
r.synchronize()
print("upload + pack %.2f s" % (time.perf_counter() - t), flush=True)
for i in range(3):
    t = time.perf_counter()
    r.updateLICVolume(); r.synchronize()
    dt = time.perf_counter() - t
    ms = r.lastKernelMs()
    vox = n ** 3
    print("lic_volume %d^3: %.2f ms (kernel %.2f ms) = %.3f G voxels/s = %.3f G LIC taps/s" % (n, dt * 1e3, ms, vox / ms / 1e6, vox * 65 / ms / 1e6), flush=True)
print("volume sha1 %s" % hashlib.sha1(r.readLICVolume().tobytes()).hexdigest()[:16], flush=True)
for i in range(3):
    t = time.perf_counter()
    r.render(True); r.synchronize()
    dt = time.perf_counter() - t
    print("volume_raycast %dx%d: %.2f ms (kernel %.3f ms), %d ray samples, %.2f G samples/s" % (size, size, dt * 1e3, r.lastKernelMs(), r.lastRaySamples(), r.lastRaySamples() / r.lastKernelMs() / 1e6), flush=True)
