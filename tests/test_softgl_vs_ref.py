"""The step the reference leaves to the GL implementation -- clipping, culling, rasterisation of its proxy geometry -- restated
from the OpenGL 2.1 specification (oracle/softgl.py) and applied to the primitives that Renderer::render(true) itself emits
(VV/renderer.cpp compiled unmodified, GL calls captured: oracle/ref_host_driver.cpp vvref_raycast_draws).

Checked here:
  * the oracle's analytic ray entry points (pixel_ray) are the gl_TexCoord[0] of the rasterised fragments: same coverage, same
    coordinates, for plain, moved, anisotropic, near-plane-clipped and user-clipped views (cube faces + cap polygons);
  * the oracle's slice fragments are the rasterised fragments of the reference's slice polygons, pixel by pixel and in order;
  * the reference's shader run on the rasterised fragments gives the oracle's frame;
  * the low-res preset renders into a (w/2, h/2) viewport with the window's aspect ratio;
  * the clip-plane normal is normalised in place by the first cap draw (first frame != steady state for non-unit normals).
Runs wherever oracle/_ref/libvv_ref.so exists (built from /root/reference; travels to the GPU box)."""
import ctypes

import numpy as np
import pytest

from oracle import refhost, refshim, softgl

pytestmark = pytest.mark.skipif(not (refshim.available() and refhost.available()), reason="oracle/_ref/libvv_ref.so not built (needs /root/reference)")

EDGE_EPS = 1e-6      # pixels: a fragment centre this close to an edge line may go either way (implementation-defined)
TC_EPS = 3e-6        # texcoord units: fp32 rounding of the oracle's entry point + conditioning of the interpolation


def _cams():
    from vectorvisualization_b200 import fields as F
    return dict(
        default=None,
        close=F.CAMERA_CLOSE,
        rot=dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0.0, 0.2), dist=3.0, fovy=35.0),
        back=dict(quat=F.quat_from_axis_angle((0.0, 1.0, 0.0), 180.0), pos=(0.0, 0.0, 0.0), dist=3.0, fovy=35.0),
        # 0.037 outside the front face: all of what the frustum sees of that face lies nearer than the near plane (0.1)
        near=dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 30.0), pos=(0.0, 0.0, 0.0), dist=0.62, fovy=35.0),
        # the near plane cuts the front face: a hole in the middle of the frame
        near_partial=dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 50.0), pos=(0.0, 0.0, 0.0), dist=0.75, fovy=35.0),
        inside=dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 30.0), pos=(0.0, 0.0, 0.0), dist=0.3, fovy=35.0))


PLANES = dict(
    none=(),
    front=((0.0, 0.0, -1.0, 0.1),),
    back=((0.0, 0.0, 1.0, 0.1),),
    two_nonunit=((0.3, 0.5, -0.8, 0.05), (1.0, 0.0, 0.0, 0.12)),                        # |n| = 0.99: steady state = (n / |n|, d)
    three=((0.6, -0.8, 0.0, 0.1), (-1.0, 0.0, 0.0, -0.2), (0.0, 0.3, 1.0, 0.2)))


def _scene(cam, size=72, aniso=False):
    from vectorvisualization_b200 import configs, fields as F
    s = configs.cfg1(n=12, size=size, camera=_cams()[cam])
    if aniso:
        s.height = size - 22
        s.field = np.ascontiguousarray(F.abc_flow(24)[::2, :16, :])            # 24 x 16 x 12, anisotropic spacing
        s.slice_dist = (1.0, 1.5, 2.0)
    return s


def _dat(tmp_path, s):
    from vectorvisualization_b200 import fields as F
    dat = F.write_dat(str(tmp_path / "vol.dat"), s.field, slice_thickness=s.slice_dist)
    with open(dat, "a") as f:
        f.write("TimeDependent: 0 0\n")
    return dat


def _oracle_rays(oracle, s):
    o = oracle.OracleScene(s)
    e = np.zeros((s.height, s.width, 4), np.float32)
    oracle.lib().vvo_pixel_rays(ctypes.byref(o.c), 0, 0, s.width, s.height, oracle._p(e))
    return e


@pytest.mark.parametrize("cam,planes,aniso", [
    ("default", "none", False), ("default", "none", True), ("close", "none", False), ("rot", "none", True), ("back", "none", False),
    ("near", "none", False), ("near_partial", "none", False), ("inside", "none", False),
    ("default", "front", False), ("close", "back", False), ("rot", "two_nonunit", False), ("close", "two_nonunit", True),
    ("back", "three", False), ("rot", "three", False), ("near", "front", False), ("inside", "two_nonunit", False)])
def test_raycast_fragments_are_the_oracle_entry_points(oracle, tmp_path, cam, planes, aniso):
    s = _scene(cam, aniso=aniso)
    s.clip_planes = PLANES[planes]
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, s.width, s.height, planes=s.clip_planes)
    prog = [d for d in draws if d["program"] == 77]
    # Renderer::raycastVolume: one GL_QUADS batch of 24 vertices with back-face culling on, then one fan per active plane
    assert prog[0]["mode"] == softgl.GL_QUADS and len(prog[0]["verts"]) == 24 and prog[0]["cull"] == 1
    assert len(prog) == 1 + len(s.clip_planes) and all(d["clip_mask"] == (1 << len(s.clip_planes)) - 1 for d in prog)
    assert all(d["viewport"] == [0, 0, s.width, s.height] for d in prog)
    tex, count, edge, _ = softgl.rasterize(draws, s.width, s.height, program=77)
    e = _oracle_rays(oracle, s)
    mism = tex[..., 3] != e[..., 3]
    assert mism.sum() <= 0.005 * mism.size and (edge[mism] < EDGE_EPS).all(), (int(mism.sum()), edge[mism])
    both = (tex[..., 3] == 1) & (e[..., 3] == 1)
    if cam in ("near", "inside") and planes == "none":
        assert not tex[..., 3].any() and not e[..., 3].any()          # the near plane / the culled back faces leave nothing
    else:
        assert both.sum() > 150
        assert np.abs(tex[both][:, :3] - e[both][:, :3]).max() < TC_EPS
    if cam == "near_partial":
        assert 0.02 < both.mean() < 0.9                                # a hole, not an empty and not a full frame
    # cube faces never overlap; a cap sits 1e-4 inside the kept half-space, so it can overlap a clipped face by a sliver
    assert (count > 1).sum() <= (0 if not s.clip_planes else 0.005 * count.size)


def test_clip_normal_is_normalised_by_the_first_cap_draw(oracle, tmp_path):
    """VV/renderer.cpp:1301 hands ClipPlane::getNormal() (the object's own array) to ViewSlicing::setupSingleSlice, which
    normalises its argument in place (VV/slicing.cpp:337-348): the plane the GL clips with is (n, d) in the first frame and
    (n / |n|, d) from the second frame on"""
    s = _scene("rot")
    plane = (0.0, 0.6, 2.0, 0.2)
    ln = float(np.linalg.norm(plane[:3]))
    dat = _dat(tmp_path, s)
    eq = []
    for frames in (1, 2, 3):
        d = [d for d in refhost.raycast_draws(dat, s.camera, s.width, s.height, planes=(plane,), frames=frames) if d["program"] == 77][0]
        # back to the object space of the glClipPlane call (volume-centred): p_obj = p_eye M, M = frame model-view * T(center)
        o = oracle.OracleScene(s)
        T = np.eye(4); T[:3, 3] = list(o.c.center)
        eq.append(d["clip_eye"][0] @ (d["modelview"] @ T))
    assert np.abs(eq[0] - np.asarray(plane)).max() < 1e-9
    assert np.abs(eq[1] - np.asarray([plane[0] / ln, plane[1] / ln, plane[2] / ln, plane[3]])).max() < 1e-9
    assert np.abs(eq[2] - eq[1]).max() < 1e-12
    # the oracle (and vv_set_clip_plane) evaluate the steady state
    s.clip_planes = (plane,)
    tex, _, edge, _ = softgl.rasterize(refhost.raycast_draws(dat, s.camera, s.width, s.height, planes=(plane,), frames=2), s.width, s.height, 77)
    e = _oracle_rays(oracle, s)
    mism = tex[..., 3] != e[..., 3]
    assert (edge[mism] < EDGE_EPS).all()
    first, _, _, _ = softgl.rasterize(refhost.raycast_draws(dat, s.camera, s.width, s.height, planes=(plane,), frames=1), s.width, s.height, 77)
    assert (first[..., 3] != e[..., 3]).sum() > 20                      # the first frame really is different


@pytest.mark.parametrize("cam,planes", [("default", "none"), ("rot", "none"), ("near_partial", "none"), ("inside", "none"),
                                        ("close", "front"), ("rot", "three_unit")])
def test_slicing_fragments_are_the_oracle_slice_fragments(oracle, tmp_path, cam, planes):
    import vectorvisualization_b200 as vv
    s = _scene(cam, size=40)
    s.params.update(stepSizeVol=1 / 32)
    s.technique = vv.VOLIC_SLICING
    # unit normals: Renderer::sliceVolume never draws caps, so nothing normalises a non-unit normal in a slicing-only session
    s.clip_planes = ((0.6, 0.0, -0.8, 0.05), (1.0, 0.0, 0.0, 0.12), (0.0, -0.8, 0.6, 0.2)) if planes == "three_unit" else PLANES[planes]
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, s.width, s.height, planes=s.clip_planes, slicing=True,
                                  step_size_vol=s.lic_params().stepSizeVol)
    o = oracle.OracleScene(s)
    _, _, nslices = o.slicing_setup()
    prog = [d for d in draws if d["program"] == 79]
    assert len(prog) == nslices and all(d["mode"] == softgl.GL_TRIANGLE_FAN and d["cull"] == 0 for d in prog)
    starts, frags, edge = softgl.fragment_lists(draws, s.width, s.height, program=79)
    buf = np.zeros((nslices, 4), np.float32)
    total = differ = 0
    worst = 0.0
    for y in range(s.height):
        for x in range(s.width):
            n = oracle.lib().vvo_slice_fragments(ctypes.byref(o.c), x, y, oracle._p(buf), nslices)
            g = frags[starts[y * s.width + x]:starts[y * s.width + x + 1]]
            total += n
            if len(g) != n:
                differ += 1
                assert edge[y, x] < EDGE_EPS, (x, y, len(g), n)
            elif n:
                worst = max(worst, float(np.abs(g[:, :3] - buf[:n, :3]).max()))       # same fragments in the same (front-to-back) order
                assert np.array_equal(g[:, 3], buf[:n, 3])                            # ... of the same slices (-> the same ping-pong target)
    assert total > 2000 and differ <= 4 and worst < TC_EPS


def test_rasterised_fragments_through_the_reference_shaders(oracle, tmp_path):
    """reference geometry (Renderer::render) -> GL-spec rasteriser -> reference shader code  ==  the oracle's frame"""
    import vectorvisualization_b200 as vv
    from util import psnr8
    # ray-cast program, rotated view with two clip planes
    s = _scene("rot", size=40)
    s.clip_planes = ((0.6, 0.0, -0.8, 0.05), (1.0, 0.0, 0.0, 0.12))
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, s.width, s.height, planes=s.clip_planes)
    tex, _, edge, _ = softgl.rasterize(draws, s.width, s.height, program=77)
    ref_img, _, ref_n = refshim.RefScene(s).raycast(texcoords=tex)
    img, cnt, n = oracle.OracleScene(s).raycast()
    interior = edge > EDGE_EPS
    a, b = oracle.quantize_rgba8(ref_img)[interior], oracle.quantize_rgba8(img)[interior]
    assert n > 20000 and abs(ref_n - n) <= 0.002 * n
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1 and psnr8(a, b) >= 55.0
    # slicing program
    s = _scene("close", size=36)
    s.params.update(stepSizeVol=1 / 32)
    s.technique, s.tf_mode, s.gate_mode = vv.VOLIC_SLICING, vv.TF_A, vv.GATE_TF_ALPHA
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, s.width, s.height, slicing=True, step_size_vol=1 / 32)
    starts, frags, edge = softgl.fragment_lists(draws, s.width, s.height, program=79)
    ref_img, _, ref_n = refshim.RefScene(s).slicing(fragments=(starts, frags))
    img, cnt, n = oracle.OracleScene(s).slicing()
    interior = edge > EDGE_EPS
    a, b = oracle.quantize_rgba8(ref_img)[interior], oracle.quantize_rgba8(img)[interior]
    assert n > 1000 and abs(ref_n - n) <= 0.002 * n
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1 and psnr8(a, b) >= 55.0


def test_lowres_preset_viewport(oracle, tmp_path):
    """VV/renderer.cpp:111-119,159-160: the low-res preset rasterises into a (w/2, h/2) viewport while gluPerspective keeps the
    window's aspect ratio (Camera::setWindow): the frame a caller gets from vv_enable_lowres + vv_resize(w/2, h/2)"""
    s = _scene("close", size=64)
    s.height = 48
    s.lowres = 1
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, s.width, s.height, lowres=1)
    prog = [d for d in draws if d["program"] == 77]
    assert prog[0]["viewport"] == [0, 0, 32, 24]
    assert abs(prog[0]["projection"][1, 1] / prog[0]["projection"][0, 0] - 64 / 48) < 1e-6      # aspect is a float in Camera
    tex, _, edge, _ = softgl.rasterize(draws, 32, 24, program=77)
    s.width, s.height = 32, 24
    e = _oracle_rays(oracle, s)
    mism = tex[..., 3] != e[..., 3]
    assert (edge[mism] < EDGE_EPS).all()
    both = (tex[..., 3] == 1) & (e[..., 3] == 1)
    assert both.sum() > 150 and np.abs(tex[both][:, :3] - e[both][:, :3]).max() < TC_EPS
    # odd window: 51 x 41 -> a 25 x 20 frame whose projection keeps 51 / 41 (vv_set_window + vv_resize)
    s = _scene("close", size=51)
    s.height = 41
    s.lowres = 1
    draws = refhost.raycast_draws(_dat(tmp_path, s), s.camera, 51, 41, lowres=1)
    assert [d for d in draws if d["program"] == 77][0]["viewport"] == [0, 0, 25, 20]
    tex, _, edge, _ = softgl.rasterize(draws, 25, 20, program=77)
    s.width, s.height, s.window = 25, 20, (51, 41)
    e = _oracle_rays(oracle, s)
    mism = tex[..., 3] != e[..., 3]
    assert (edge[mism] < EDGE_EPS).all()
    both = (tex[..., 3] == 1) & (e[..., 3] == 1)
    assert both.sum() > 100 and np.abs(tex[both][:, :3] - e[both][:, :3]).max() < TC_EPS
    s.window = None                                  # without the window's aspect the frame would be a different one
    e2 = _oracle_rays(oracle, s)
    assert np.abs(e2[both][:, :3] - e[both][:, :3]).max() > 1e-3


def test_lic_volume_layers_address_voxel_centres():
    """Renderer::updateLICVolume (VV/renderer.cpp:1311-1374, VV/VolumeBuffer.cpp:100-108): one screen-filling quad per layer of
    the w x h x d target; rasterised, fragment (x, y) of layer z carries texcoord0 = ((x + .5) / w, (y + .5) / h, (z + .5) / d)
    -- the positions the oracle's vvo_lic_volume and the CUDA lic_volume_kernel evaluate"""
    w, h, d = 7, 5, 3
    draws = refhost.licvolume_draws(w, h, d)
    prog = [x for x in draws if x["program"] == 80]
    assert len(prog) == d and all(x["viewport"] == [0, 0, w, h] and x["mode"] == softgl.GL_QUADS for x in prog)
    for z, x in enumerate(prog):
        assert np.array_equal(x["modelview"], np.eye(4)) and np.array_equal(x["projection"], np.eye(4))
        tex, count, _, _ = softgl.rasterize([x], w, h)
        assert (count == 1).all()
        gy, gx = np.mgrid[0:h, 0:w]
        want = np.stack([(gx + 0.5) / w, (gy + 0.5) / h, np.full(gx.shape, (np.float32(z) + np.float32(0.5)) / np.float32(d), np.float64)], axis=-1)
        assert np.abs(tex[..., :3] - want).max() < 1e-12


def test_draw_state_of_the_three_techniques(tmp_path):
    """What state the reference issues its proxy geometry under (Renderer::raycastVolume / sliceVolume / raycastLICVolume,
    VV/renderer.cpp:1093-1120, 1123-1267, 1376-1405).  LIC ray-cast and FBO slicing draw with blending off: the frame is the
    shader's gl_FragColor (the parity point).  The LIC-volume ray-cast leaves GL_BLEND ENABLED (:1389) with whatever blend function
    was set last (renderLICVolume :1337, the HUD and the TF editor set SRC_ALPHA / ONE_MINUS_SRC_ALPHA): what reaches the frame
    buffer there depends on that state and on the clear colour, so -- like Q19 -- the parity point of volume_raycast is
    gl_FragColor itself."""
    s = _scene("default", size=24)
    dat = _dat(tmp_path, s)
    ray = [d for d in refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=0) if d["program"] == 77]
    sli = [d for d in refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=1) if d["program"] == 79]
    vol = [d for d in refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=2) if d["program"] == 81 and d["mode"] == softgl.GL_QUADS]
    assert ray and all(d["blend"] == 0 and d["cull"] == 1 for d in ray)
    assert sli and all(d["blend"] == 0 and d["cull"] == 0 for d in sli)
    assert len(vol) == 1 and vol[0]["blend"] == 1 and vol[0]["cull"] == 1
    # same proxy cube, same matrices for the two ray-cast techniques: same fragments
    a, _, _, _ = softgl.rasterize(ray, s.width, s.height)
    b, _, _, _ = softgl.rasterize(vol, s.width, s.height)
    assert np.array_equal(a, b) and a[..., 3].sum() > 50


def test_fbo_slicing_ping_pong_targets(tmp_path):
    """Renderer::sliceVolume with the FBO (VV/renderer.cpp:1201-1225), read from the reference's own GL calls: the two image textures
    are swapped before EVERY slice -- slice i is drawn into one of them with the other (the target of slice i - 1) bound as
    imageFBOSampler -- so a pixel that slice i does not cover keeps, in slice i's target, what slice i - 2 left there; the frame that
    is displayed / stored is the target of the last slice.  (What the oracle's fbo_pingpong switch and the CUDA path implement.)"""
    s = _scene("default", size=24)
    dat = _dat(tmp_path, s)
    draws = refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=1)
    sli = [d for d in draws if d["program"] == 79]
    assert len(sli) > 8
    tex = sorted(set(d["fbo_tex"] for d in sli))
    assert len(tex) == 2 and 0 not in tex                               # exactly two colour targets
    for i, d in enumerate(sli):
        assert d["fbo_tex"] == (sli[0]["fbo_tex"] if i % 2 == 0 else sli[1]["fbo_tex"])    # alternate, starting at slice 0
        assert d["bound_tex"] == (set(tex) - {d["fbo_tex"]}).pop()       # the source is the OTHER texture = the previous slice's target
        assert d["blend"] == 0
    # the display pass (background program) samples the target of the last slice
    bg = [d for d in draws if d["program"] == 78]
    assert bg and bg[-1]["bound_tex"] == sli[-1]["fbo_tex"]


def test_slicing_without_fbo_draw_state(tmp_path):
    """Renderer::sliceVolume without the FBO (the reference's start-up state for F3, VV/renderer.cpp:1151-1160, 1236-1262): every slice
    is drawn with lic3d_slicingblend and fixed-function blending (ONE_MINUS_DST_ALPHA, ONE) into the back buffer, then a
    screen-filling white quad is blended in with the same function and no program."""
    s = _scene("default", size=24)
    dat = _dat(tmp_path, s)
    draws = refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=3)
    sli = [d for d in draws if d["program"] == 82]
    fbo = [d for d in refhost.raycast_draws(dat, s.camera, s.width, s.height, slicing=1) if d["program"] == 79]
    assert len(sli) == len(fbo) > 8
    GL_ONE_MINUS_DST_ALPHA, GL_ONE = 0x0305, 1
    assert all(d["blend"] == 1 and d["blend_func"] == (GL_ONE_MINUS_DST_ALPHA, GL_ONE) and d["fbo_tex"] == 0 for d in sli)
    assert all(np.array_equal(a["verts"], b["verts"]) and np.array_equal(a["tex"], b["tex"]) for a, b in zip(sli, fbo))   # same polygons
    k = next(i for i, d in enumerate(draws) if d is sli[-1])
    white = draws[k + 1]
    assert white["program"] == 0 and white["mode"] == softgl.GL_QUADS and white["blend"] == 1
    assert white["blend_func"] == (GL_ONE_MINUS_DST_ALPHA, GL_ONE) and len(white["verts"]) == 4
