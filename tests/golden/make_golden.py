#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE'S OWN SHADER CODE (oracle/_ref/libvv_ref.so, built by
oracle/build_ref.py from /root/reference/VectorVisualization/shader/*.glsl).  Run in the build container (where the
reference is mounted); the vectors are committed so the GPU box can check both the oracle and the CUDA path against
outputs of the reference without the reference being present.

A third family (y_chain_*) goes through the whole reference chain: the draw calls of Renderer::render(true) captured in the
shim, rasterised per the GL specification (oracle/softgl.py), shaded by the reference's shader code (oracle/refchain.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import refshim  # noqa: E402
from oracle import vvo  # noqa: E402
from scenes import golden_scenes, golden_extra_scenes, golden_chain_scenes  # noqa: E402


def main():
    if not refshim.available():
        raise SystemExit("oracle/_ref/libvv_ref.so missing: run python oracle/build_ref.py (needs /root/reference)")
    for name, mk in golden_scenes().items():
        s = mk()
        r = refshim.RefScene(s)
        img, cnt, tot = r.raycast()
        s.licvol_fp16 = 0
        lv = r.lic_volume((12, 12, 12))
        nz, ny, nx = s.field.shape[:3]
        lv_full = r.lic_volume((nx, ny, nz))
        vimg, vcnt, vtot = r.raycast_licvolume(lv_full)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, raycast=img, raycast_samples=cnt.astype(np.uint16), total=np.int64(tot),
                            licvol12=lv, volraycast=vimg, volraycast_samples=vcnt.astype(np.uint16))
        print("%-32s ray samples %7d  -> %s (%d bytes)" % (name, tot, os.path.basename(out), os.path.getsize(out)))
    tables = vvo.illum_tables(40.0)      # bit-identical to VV/illumination.cpp (tests/test_host_vs_ref.py)
    for name, (mk, kind) in golden_extra_scenes().items():
        s = mk()
        need_tables = "MALLO" in (s.defines or "") or "ZOECKLER" in (s.defines or "")
        r = refshim.RefScene(s, illum_tables=tables if need_tables else None)
        img, cnt, tot = r.slicing() if kind == "slicing" else r.raycast()
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, frame=img, samples=cnt.astype(np.uint16), total=np.int64(tot))
        print("%-32s %-8s samples %7d  -> %s (%d bytes)" % (name, kind, tot, os.path.basename(out), os.path.getsize(out)))
    from oracle import refchain
    for name, (mk, kind) in golden_chain_scenes().items():
        s = mk()
        img, cnt, tot, edge = refchain.reference_chain_frame(s)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, frame=img, samples=cnt.astype(np.uint16), total=np.int64(tot), edge=edge)
        print("%-36s %-8s samples %7d, %d edge pixels -> %s (%d bytes)" % (name, kind, tot, int(edge.sum()), os.path.basename(out), os.path.getsize(out)))


if __name__ == "__main__":
    main()
