"""refchain.py -- TEST INFRASTRUCTURE ONLY: one frame through the whole reference chain, as far as it can run without a GL.

    Renderer::render(true)            VV/renderer.cpp compiled unmodified, its GL calls captured   (refhost.raycast_draws)
 -> clipping, culling, rasterisation  per the OpenGL 2.1 specification                              (softgl)
 -> fragment shader                   VV/shader/*.glsl compiled as C++                             (refshim.RefScene)

Texture contents and uniform values are those the reference's host code produces (checked bit for bit in
tests/test_host_vs_ref.py).  Needs oracle/_ref/libvv_ref.so, i.e. /root/reference at build time.
"""
import os
import tempfile

import numpy as np

from . import refhost, refshim, softgl

EDGE_EPS = 1e-6      # pixels; fragment centres closer than this to an edge line are implementation-defined


def reference_chain_frame(s, illum_tables=None):
    """s: vectorvisualization_b200.configs.Scene (ray-cast or slicing technique).
    Returns (rgba float32 [h][w][4], samples uint32 [h][w], total, edge bool [h][w] -- True where coverage is implementation-defined)"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import fields as F
    slicing = s.technique == vv.VOLIC_SLICING
    with tempfile.TemporaryDirectory() as tmp:
        dat = F.write_dat(os.path.join(tmp, "vol.dat"), s.field, slice_thickness=s.slice_dist)
        with open(dat, "a") as f:
            f.write("TimeDependent: 0 0\n")
        draws = refhost.raycast_draws(dat, s.camera, s.width, s.height, lowres=s.lowres, planes=tuple(getattr(s, "clip_planes", ()) or ()),
                                      slicing=1 if slicing else 0, step_size_vol=s.lic_params().stepSizeVol)
    r = refshim.RefScene(s, illum_tables=illum_tables)
    if slicing:
        starts, frags, edge = softgl.fragment_lists(draws, s.width, s.height, program=79)
        img, cnt, tot = r.slicing(fragments=(starts, frags))
    else:
        tex, _, edge, _ = softgl.rasterize(draws, s.width, s.height, program=77)
        img, cnt, tot = r.raycast(texcoords=tex)
    return img, cnt, tot, edge < EDGE_EPS
