"""A/B runner for compile-time kernel variants (no torch import: a call costs seconds of GPU time).

    python scripts/ab.py [cfg=cfg3] [loop=20] [view=default|close] lib1.so lib2.so@WALK_FAST_PATHS:0 ...

(`lib.so@OPTION:VALUE[,OPTION:VALUE]` sets vv_set_option knobs of that run: A/B of run-time switches with one library.)

Every library (built with `python -m vectorvisualization_b200.build --variant <tag> -D...`) runs in its own process:
back-to-back frames of the configuration, ms/frame (wall over the loop, one synchronize at the end), the library's own
event-timed duration of the dominant kernel, the ray-sample count, and a SHA-1 of the RGBA32F frame -- variants must agree
on the hash (bit-identical frames) to be eligible.  One JSON line per library."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(cfg, loop, view):
    sys.path.insert(0, ROOT)
    import numpy as np
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    kw = {}
    if view == "close":
        kw["camera"] = F.CAMERA_CLOSE
    s = getattr(configs, cfg)(**kw)
    r = vv.Renderer(0)
    sets = os.environ.get("VV_AB_SET", "")
    for kv in filter(None, sets.split(",")):
        k, v = kv.split(":")
        r.setOption(getattr(vv, "OPT_" + k), int(v))
    configs.apply_scene(r, s)
    for _ in range(3):
        r.render(True)
    r.synchronize()
    t = time.perf_counter()
    for _ in range(loop):
        r.render(True)
    r.synchronize()
    ms = (time.perf_counter() - t) / loop * 1e3
    kms = []
    for _ in range(5):
        r.render(True)
        r.synchronize()
        kms.append(r.lastKernelMs())
    img = r.readRGBA32F()
    print(json.dumps({"lib": os.path.basename(vv.LIB_PATH) + ("@" + sets if sets else ""), "cfg": cfg, "view": view, "ms_per_frame": round(ms, 4),
                      "kernel_ms": round(float(np.median(kms)), 4), "ray_samples": int(r.lastRaySamples()),
                      "launches": int(r.lastLaunchCount()), "sha1": hashlib.sha1(np.ascontiguousarray(img).tobytes()).hexdigest()[:16]}), flush=True)


def main():
    opts = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a and ".so" not in a)
    libs = [a for a in sys.argv[1:] if ".so" in a and "=" not in a]
    cfg, loop, view = opts.get("cfg", "cfg3"), int(opts.get("loop", "20")), opts.get("view", "default")
    if os.environ.get("VV_AB_CHILD"):
        return child(cfg, loop, view)
    for lib in libs or [os.path.join(ROOT, "vectorvisualization_b200", "libvv_b200.so")]:
        lib, _, sets = lib.partition("@")
        env = dict(os.environ, VV_AB_CHILD="1", VV_B200_LIB=os.path.abspath(lib), VV_AB_SET=sets)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "cfg=" + cfg, "loop=%d" % loop, "view=" + view],
                           env=env, capture_output=True, text=True, timeout=600)
        out = [l for l in p.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else json.dumps({"lib": os.path.basename(lib), "error": (p.stderr or p.stdout)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
