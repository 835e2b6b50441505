"""The oracle's restatement of the reference's HOST code (packing, noise gradients, filter kernel, transfer function,
DAT reader, argument parser, quaternion helpers) against that host code itself, compiled unmodified from
/root/reference against a capturing GL stub (oracle/_ref, oracle/ref_host_driver.cpp).  Also checks the product's
host-only loaders (vv_parse_dat / vv_parse_args / vv_png_read) against the reference's."""
import ctypes
import os

import numpy as np
import pytest

from oracle import refhost

pytestmark = pytest.mark.skipif(not refhost.available(), reason="oracle/_ref host objects not built (needs /root/reference)")


def test_vector_texture_pack(oracle, tmp_path):
    from vectorvisualization_b200 import fields as F
    f0, f1 = F.tornado(20), F.abc_flow(20)
    f0[3, 4, 5] = 0.0                                      # zero vector -> rgb 0.5 (VV/dataset.cpp:596-602)
    dat = F.write_dat(str(tmp_path / "vec.dat"), None, time_steps=[f0, f1], slice_thickness=(1, 1.5, 2))
    for idx in (0, 3):
        ref, geom, ifmt, wrap = refhost.vector_texture(dat, (20, 20, 20), interp=(idx, 10))
        assert ifmt == refhost.GL_RGBA16F_ARB and wrap == refhost.GL_CLAMP_TO_EDGE
        mine = oracle.pack_vector_field(f0, f1, (idx, 10), fp16=False)
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
        # RGBA16F upload: round to nearest even half
        mine16 = oracle.pack_vector_field(f0, f1, (idx, 10), fp16=True)
        assert np.array_equal(mine16, ref.astype(np.float16).astype(np.float32))
    ext, sc, sci, cen = (np.zeros(3, np.float32) for _ in range(4))
    oracle.lib().vvo_volume_geometry((ctypes.c_int * 3)(20, 20, 20), (ctypes.c_float * 3)(1, 1.5, 2), oracle._p(ext), oracle._p(sc),
                                     oracle._p(sci), oracle._p(cen))
    assert np.array_equal(ext, geom["extent"]) and np.array_equal(cen, geom["center"])
    assert np.array_equal(sc, geom["scale"][:3]) and np.array_equal(sci, geom["scale_inv"][:3])
    assert geom["scale"][3] == 0.0 and geom["scale_inv"][3] == 1.0     # VV/dataset.cpp:173-174


def test_single_time_step(oracle, tmp_path):
    from vectorvisualization_b200 import fields as F
    f0 = F.rankine_vortex(16)
    dat = F.write_dat(str(tmp_path / "one.dat"), f0)
    # DatFile::_timestep is only initialised by a TimeDependent line (VV/reader.cpp:60-68,187-203): without one the
    # reference reads an indeterminate time step.  Give it the line; the product accepts both forms.
    with open(dat, "a") as f:
        f.write("TimeDependent: 0 0\n")
    ref, _, _, _ = refhost.vector_texture(dat, (16, 16, 16))
    assert np.array_equal(oracle.pack_vector_field(f0, None, (0, 10), fp16=False).view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("shape", [(16, 16, 16), (9, 14, 11)])
def test_noise_gradients(oracle, tmp_path, shape):
    from vectorvisualization_b200 import fields as F
    rng = np.random.RandomState(3)
    noise = (rng.rand(*shape) < 0.3).astype(np.uint8) * 255 if shape[0] == 16 else rng.randint(0, 256, size=shape).astype(np.uint8)
    path = F.write_noise(str(tmp_path / "noise"), noise)
    ref, ifmt, wrap = refhost.noise_texture(path, shape, True)
    assert ifmt == refhost.GL_RGBA and wrap == refhost.GL_REPEAT
    assert np.array_equal(oracle.pack_noise_rgba(noise, oracle.noise_gradients(noise)), ref)
    ref1, ifmt1, _ = refhost.noise_texture(path, shape, False)
    assert ifmt1 == refhost.GL_LUMINANCE and np.array_equal(ref1, noise)
    # intermediate stages, bit for bit
    g, f, q = refhost.noise_gradients(noise)
    og = np.zeros_like(g)
    dims = (ctypes.c_int * 3)(*shape[::-1]); sd = (ctypes.c_float * 3)(1, 1, 1)
    oracle.lib().vvo_compute_gradients_f(oracle._p(noise), dims, sd, oracle._p(og))
    assert np.array_equal(og.view(np.uint32), g.view(np.uint32))
    oracle.lib().vvo_filter_gradients_f(dims, oracle._p(og))
    assert np.array_equal(og.view(np.uint32), f.view(np.uint32))
    assert np.array_equal(oracle.noise_gradients(noise), q)


def test_gradient_cache_file(oracle, tmp_path):
    """.grd cache (VV/gradient.cpp:93-187, VV/dataset.cpp:1238-1267): the product reads what the reference's saveGradients
    writes, and the reference's loadGradients reads what the product writes"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import fields as F
    shape = (9, 14, 11)
    noise = np.random.RandomState(5).randint(0, 256, size=shape).astype(np.uint8)
    path = F.write_noise(str(tmp_path / "noise"), noise)
    dims = shape[::-1]
    with pytest.raises(RuntimeError):
        vv.grd_read(path, dims)                                   # no cache yet
    ref = refhost.noise_texture_cached(path, shape)               # the reference computes and saves <path>.grd
    assert os.path.getsize(path + ".grd") == 3 * noise.size
    got = vv.grd_read(path, dims)
    assert np.array_equal(got, oracle.noise_gradients(noise)) and np.array_equal(got, ref[..., :3])
    # the other direction: a cache written by the product is what the reference loads (sentinel values prove it is used)
    sentinel = np.random.RandomState(6).randint(0, 256, size=shape + (3,)).astype(np.uint8)
    vv.grd_write(path, sentinel)
    ref2 = refhost.noise_texture_cached(path, shape)
    assert np.array_equal(ref2[..., :3], sentinel) and np.array_equal(ref2[..., 3], noise)
    with open(path + ".grd", "r+b") as f:
        f.truncate(10)
    with pytest.raises(RuntimeError):
        vv.grd_read(path, dims)                                   # short file: "Reading gradients ... failed"


def test_filter_kernels(oracle, tmp_path):
    from vectorvisualization_b200 import fields as F
    data, inv, wrap = refhost.filter_texture(None)
    mine, minv = oracle.box_filter(256)
    assert wrap == refhost.GL_CLAMP and inv == minv and np.array_equal(data, mine)
    for name, width in (("gaussian", 256), ("cos2", 256), ("triangle", 200), ("box", 37)):
        row = F.filter_kernel(name, width)
        png = F.write_png(str(tmp_path / (name + ".png")), row[None, :])
        data, inv, _ = refhost.filter_texture(png)
        mine, minv = oracle.filter_from_row(row)
        assert np.array_equal(data, mine) and inv == minv, name
    # RGB kernel image: first channel is used (VV/dataset.cpp:1439-1461)
    rgb = np.stack([F.filter_kernel("cos2", 64), np.zeros(64, np.uint8), np.full(64, 9, np.uint8)], axis=-1)
    png = F.write_png(str(tmp_path / "rgb.png"), rgb[None])
    data, inv, _ = refhost.filter_texture(png)
    mine, minv = oracle.filter_from_row(rgb.reshape(-1), channels=3)
    assert np.array_equal(data, mine) and inv == minv


def test_transfer_function(oracle, tmp_path):
    from vectorvisualization_b200 import fields as F
    rgba, la, loaded = refhost.tf_textures(None)
    tf = oracle.default_tf()
    assert not loaded and np.array_equal(rgba, tf[:, :4]) and np.array_equal(la, tf[:, 3:5])
    assert np.array_equal(tf, F.default_tf())
    preset = F.tf_preset("tf-length")
    name = F.write_tf(str(tmp_path / "tf.png"), preset)
    rgba, la, loaded = refhost.tf_textures(name)
    assert loaded
    assert np.array_equal(rgba, preset[:, :4])
    # loadTF short-circuit (VV/transferEdit.cpp:239): with the RGBA file present the alpha/opacity file is not read
    assert np.array_equal(la[:, 0], preset[:, 3]) and np.array_equal(la[:, 1], tf[:, 4])


def test_dat_reader_matches(vv, tmp_path):
    from vectorvisualization_b200 import fields as F
    cases = []
    cases.append(F.write_dat(str(tmp_path / "a.dat"), F.abc_flow(8), slice_thickness=(1, 2, 0.5)))
    cases.append(F.write_dat(str(tmp_path / "b.dat"), np.zeros((4, 5, 6), np.uint8)))
    cases.append(F.write_dat(str(tmp_path / "c.dat"), None, time_steps=[F.abc_flow(4)] * 3))
    p = tmp_path / "d.dat"
    p.write_text("# comment\nObjectFileName: a.raw\nResolution: 8 8 8\nFormat: USHORT\nFoo: bar\n")
    cases.append(str(p))
    for c in cases:
        ref = refhost.parse_dat(c)
        mine = vv.parse_dat(c)
        assert ref is not None
        assert tuple(mine.resolution) == ref["resolution"]
        assert tuple(mine.slice_thickness) == ref["slice_thickness"]
        assert mine.data_type == ref["data_type"] and mine.data_dim == ref["data_dim"]
        assert (mine.time_begin, mine.time_end) == ref["time"]
    bad = tmp_path / "bad.dat"
    bad.write_text("ObjectFileName: missing.raw\nResolution: 2 2 2\nFormat: UCHAR\n")
    assert refhost.parse_dat(str(bad)) is None
    with pytest.raises(vv.VVError):
        vv.parse_dat(str(bad))


ARGV = [
    ["volic", "vol.dat"],
    ["volic", "vol.dat", "-g", "-n", "noise_256", "-f", "k.png", "-t", "tf.png"],
    ["volic", "--gradient", "--noise=n", "--filter=f.png", "--transfer=t.png", "vol.dat"],
    ["volic", "vol.dat", "-l", "-r", "out.txt", "-s", "halton.txt"],
    ["volic", "vol.dat", "other.dat"],
    ["volic", "vol.dat", "-f"],
    ["volic", "vol.dat", "-n", "-g"],
    ["volic", "vol.dat", "-x"],
    ["volic", "vol.dat", "--bogus"],
    ["volic", "vol.dat", "--filter"],
    ["volic", "--lambda2", "--redirect=r", "--halton=h", "v.dat"],
    ["volic"],
]


@pytest.mark.parametrize("argv", ARGV, ids=[" ".join(a[1:]) or "none" for a in ARGV])
def test_parse_args_matches(vv, argv):
    ok, ref = refhost.parse_args(argv)
    try:
        a = vv.parse_args(argv)
        mine_ok = True
    except vv.VVError:
        mine_ok = False
    assert mine_ok == ok
    if ok:
        assert a.vol_file.decode() == ref["vol"] and a.noise_file.decode() == ref["noise"]
        assert a.tf_file.decode() == ref["tf"] and a.filter_file.decode() == ref["filter"]
        assert a.redirect_file.decode() == ref["redirect"] and a.halton_file.decode() == ref["halton"]
        assert bool(a.use_gradients) == ref["gradients"] and bool(a.use_lambda2) == ref["lambda2"]


def test_quaternion_helpers(oracle):
    from vectorvisualization_b200 import configs, fields as F
    for axis, ang in (((1, 1, 0), 35.0), ((0.3, -1.0, 0.2), 110.0), ((0, 0, 1), 0.0), ((0, 1, 0), 179.0)):
        q = F.quat_from_axis_angle(axis, ang)
        a, ax = refhost.quat_angle_axis(q)
        s = configs.cfg1(n=8, size=16, camera=dict(quat=q, pos=(0, 0, 0), dist=4.0, fovy=35.0))
        o = oracle.OracleScene(s)
        cam, rot = np.zeros(3, np.float32), np.zeros(9, np.float32)
        oracle.lib().vvo_view(ctypes.byref(o.c), oracle._p(cam), oracle._p(rot))
        R = rot.reshape(3, 3)
        v = np.array([0.3, -0.7, 0.64], np.float32)
        # glRotatef(angle, axis) built from Quaternion_getAngleAxis must rotate like the quaternion itself
        assert np.allclose(R @ v, refhost.quat_mult_vec(q, v), atol=2e-6)
        if ang:
            assert np.allclose(a, np.deg2rad(ang), atol=1e-5)
            assert np.allclose(ax, np.asarray(axis, np.float32) / np.linalg.norm(axis), atol=1e-6)


def test_illumination_tables(vv, oracle):
    """Zoeckler / Mallo look-up tables: reference generator (captured glTexImage2D data, 8-bit internal formats) ==
    oracle restatement == the tables the product library builds"""
    z, d, s, ifmt, spec_exp = refhost.illum_tables()
    assert ifmt == (0x190A, 0x1908, 0x1908) and spec_exp == 40.0      # GL_LUMINANCE_ALPHA, GL_RGBA: floatTex is not forwarded
    q = lambda a: np.floor(np.clip(a, 0, 1).astype(np.float32) * np.float32(255) + np.float32(0.5)) / np.float32(255)
    oz, od, os_ = oracle.illum_tables(spec_exp)
    assert np.array_equal(oz, q(z)) and np.array_equal(od, q(d[..., 0])) and np.array_equal(os_, q(s[..., 0]))
    for c in (1, 2, 3):
        assert np.array_equal(d[..., 0], d[..., c]) and np.array_equal(s[..., 0], s[..., c])
    pz, pd, ps = vv.make_illum_tables(spec_exp)
    assert np.array_equal(pz, oz) and np.array_equal(pd, od) and np.array_equal(ps, os_)


# ---- slicing geometry and clip-plane caps: VV/slicing.cpp compiled unmodified, polygons captured from its gl* calls ----

def _view(oracle, o):
    cam = np.zeros(3, np.float32); rot = np.zeros(9, np.float32)
    oracle.lib().vvo_view(ctypes.byref(o.c), oracle._p(cam), oracle._p(rot))
    return cam.astype(np.float64), rot


def _modelview(o, rot):
    """column-major GL_MODELVIEW as Renderer::updateSlices reads it back (VV/renderer.cpp:1270-1292): rows of R, and the
    z translation (cam_pos.z - dist) - R3 . center; setupSlicing only uses m[2], m[6], m[10], m[3], m[7], m[11], m[14], m[15]"""
    c = o.c
    m = np.zeros(16, np.float32)
    for col in range(3):
        for row in range(3):
            m[4 * col + row] = rot[3 * row + col]
    tz = (float(c.cam_pos[2]) - float(c.cam_dist)) - sum(float(rot[6 + k]) * float(c.center[k]) for k in range(3))
    m[14] = np.float32(tz)
    m[15] = 1.0
    return m


def _inside_convex(poly, p, n, eps=1e-5):
    """p (on the polygon's plane) inside the convex polygon (any winding)"""
    sgn = 0
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        s = float(np.dot(np.cross(b - a, p - a), n))
        if abs(s) <= eps:
            continue
        if sgn == 0:
            sgn = 1 if s > 0 else -1
        elif (s > 0) != (sgn > 0):
            return False
    return True


def _scenes_for_geometry():
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    out = []
    for cam, slice_dist, step in ((None, (1, 1, 1), 1 / 64), (F.CAMERA_CLOSE, (1, 1, 1), 1 / 128),
                                  (dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0.0, 0.2), dist=3.0, fovy=35.0), (1.0, 1.5, 2.0), 1 / 32)):
        s = configs.cfg1(n=12, size=41, camera=cam)
        s.slice_dist = slice_dist
        s.params.update(stepSizeVol=step)
        s.technique = vv.VOLIC_SLICING
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        out.append(s)
    return out


def test_slicing_setup_and_polygons(oracle):
    """ViewSlicing::setupSlicing / drawSlice (VV/slicing.cpp:42-263) against the oracle's slicing geometry: view vector, depth
    range and slice count bit for bit; the oracle's fragments are exactly the pixel-ray hits inside the reference's polygons"""
    for s in _scenes_for_geometry():
        o = oracle.OracleScene(s)
        cam, rot = _view(oracle, o)
        ext = np.array(list(o.c.extent), np.float32)
        mv = _modelview(o, rot)
        step = s.lic_params().stepSizeVol
        v_ref, d_ref, n_ref = refhost.slicing_setup(mv, step, ext)
        v, d, n = o.slicing_setup()
        assert n == n_ref and np.float32(d) == np.float32(d_ref)
        assert np.array_equal(np.asarray(v, np.float32).view(np.uint32), v_ref.view(np.uint32))
        polys = [refhost.slice_polygon(mv, step, ext, i) for i in range(n_ref)]
        for verts, tex in polys:
            assert np.array_equal(verts, tex)                      # texcoord0 = vertex position (volume coordinates)
        # Renderer::sliceVolume draws drawSlice(0), drawSlice(1), ... (VV/renderer.cpp:1176-1225): increasing index must be
        # increasing depth along the view vector, which points away from the camera -> front to back
        depth = [float(np.dot(np.asarray(v_ref, np.float64), vt[0].astype(np.float64) - ext / 2)) for vt, _ in polys if len(vt) >= 3]
        assert all(b > a for a, b in zip(depth, depth[1:]))
        assert float(np.dot(np.asarray(v_ref, np.float64), ext / 2 - cam)) > 0
        buf = np.zeros((n_ref, 4), np.float32)
        ent = (ctypes.c_float * 3)(); dr = (ctypes.c_float * 3)()
        rng = np.random.RandomState(2)
        checked = 0
        for _ in range(400):
            if checked >= 25:
                break
            x, y = int(rng.randint(0, s.width)), int(rng.randint(0, s.height))
            k = oracle.lib().vvo_slice_fragments(ctypes.byref(o.c), x, y, oracle._p(buf), n_ref)
            got = buf[:k, :3].astype(np.float64)
            # pixel ray: through the camera and any fragment / entry point of that pixel
            if k == 0:
                continue
            d_ray = got[0] - cam
            d_ray /= np.linalg.norm(d_ray)
            want = []
            vn = np.asarray(v_ref, np.float64)
            for verts, _ in polys:
                if len(verts) < 3:
                    continue
                P = verts.astype(np.float64)
                dist = float(np.dot(vn, P[0] - ext / 2))
                den = float(np.dot(vn, d_ray))
                t = (dist - float(np.dot(vn, cam - ext / 2))) / den
                hit = cam + t * d_ray
                if t > 0 and _inside_convex(P, hit, vn, eps=1e-6):
                    want.append(hit)
            # polygons touching the ray at their very edge may go either way: compare the interior hits
            assert abs(len(want) - k) <= 2, (x, y, len(want), k)
            for g in got:
                assert min(np.abs(np.asarray(want) - g).max(axis=1)) < 2e-5
            checked += 1
        assert checked >= 10


@pytest.mark.parametrize("plane", [(0.0, 0.0, -1.0, 0.1), (0.3, 0.5, -0.8, 0.05), (0.6, -0.8, 0.0, 0.1), (-1.0, 0.0, 0.0, -0.2)])
def test_clip_cap_polygon(oracle, plane):
    """ClipPlane::drawSlice (VV/transform.cpp:432-444 -> VV/slicing.cpp:313-560): the cap polygon lies in n^.q = -(d - 0.0001),
    is wound to face viewers looking along +n, and contains exactly the oracle's cap entry points"""
    from vectorvisualization_b200 import configs, fields as F
    n = np.asarray(plane[:3], np.float64)
    n /= np.linalg.norm(n)
    cams = (None, F.CAMERA_CLOSE, dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0.0, 0.2), dist=3.0, fovy=35.0),
            dict(quat=F.quat_from_axis_angle((0.0, 1.0, 0.0), 180.0), pos=(0.0, 0.0, 0.0), dist=3.0, fovy=35.0))
    seen_caps = 0
    for cam_cfg in cams:
        s = configs.cfg1(n=8, size=48, camera=cam_cfg)
        o = oracle.OracleScene(s)
        ext = np.array(list(o.c.extent), np.float64)
        verts, tex = refhost.clip_cap_polygon(plane, ext.astype(np.float32))
        assert len(verts) >= 3 and np.array_equal(verts, tex)
        P = verts.astype(np.float64)
        assert np.abs((P - ext / 2) @ n + (plane[3] - 1e-4)).max() < 1e-6            # on the offset plane
        newell = sum(np.cross(P[i] - ext / 2, P[(i + 1) % len(P)] - ext / 2) for i in range(len(P)))
        assert np.dot(newell, n) < 0          # counter-clockwise seen from the -n side: front-facing for rays along +n
        cam, _ = _view(oracle, o)
        base = np.zeros((s.height, s.width, 4), np.float32)
        oracle.lib().vvo_pixel_rays(ctypes.byref(o.c), 0, 0, s.width, s.height, oracle._p(base))
        s.clip_planes = (plane,)
        o2 = oracle.OracleScene(s)
        e = np.zeros((s.height, s.width, 4), np.float32)
        oracle.lib().vvo_pixel_rays(ctypes.byref(o2.c), 0, 0, s.width, s.height, oracle._p(e))
        for y in range(0, s.height, 3):
            for x in range(0, s.width, 3):
                # the pixel ray, from the unclipped frame (any pixel that hits the box) -- else skip
                if base[y, x, 3] == 0:
                    continue
                d_ray = base[y, x, :3].astype(np.float64) - cam
                d_ray /= np.linalg.norm(d_ray)
                dn = float(np.dot(n, d_ray))
                t = (-(plane[3] - 1e-4) - float(np.dot(n, cam - ext / 2))) / dn if dn != 0 else -1
                hit = cam + t * d_ray
                expect_cap = dn > 0 and t > 0 and _inside_convex(P, hit, n, eps=1e-7) and \
                    np.all(hit > 1e-6) and np.all(hit < ext - 1e-6)
                on_cap = e[y, x, 3] > 0 and abs(float(np.dot(n, e[y, x, :3].astype(np.float64) - ext / 2)) + plane[3] - 1e-4) < 2e-6
                if expect_cap:
                    assert on_cap and np.abs(e[y, x, :3] - hit).max() < 2e-6, (x, y)
                    seen_caps += 1
                elif on_cap:
                    # only possible at the polygon's very edge
                    assert _inside_convex(P, e[y, x, :3].astype(np.float64), n, eps=1e-4)
    assert seen_caps > 20


@pytest.mark.parametrize("case", ["default", "moved_lowres", "anisotropic_illum"])
def test_renderer_camera_light_uniforms(oracle, tmp_path, case):
    """VV/renderer.cpp (setRenderVolParams, updateLightPos, updateSlices), camera.cpp and transform.cpp compiled unmodified; their
    GL matrix / uniform / light calls captured.  Checks the oracle's view (camera position, rotation), light position,
    slicing set-up and every uniform of the parameter block, including the low-res preset and Q1 (scaleVolInv -> scaleVol)."""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    s = configs.cfg1(n=12, size=48)
    if case == "moved_lowres":
        s = configs.cfg2(n=12, size=50, camera=dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.15, -0.1, 0.3), dist=3.0, fovy=35.0))
        s.height = 40
        s.lowres = 1
        s.light = dict(quat=F.quat_from_axis_angle((1, 0.2, 0), 40.0), dist=1.5)
        s.params.update(stepsForward=25, stepsBackward=40, stepSizeLIC=0.02, gradientScale=8.0, freqScale=3.0)
    if case == "anisotropic_illum":
        s = configs.cfg3(n=12, size=44, camera=F.CAMERA_CLOSE)
        s.field = np.ascontiguousarray(F.abc_flow(24)[::2, :16, :])        # 24 x 16 x 12
        s.slice_dist = (1.0, 1.5, 2.0)
        s.light = dict(quat=F.quat_from_axis_angle((0.2, 1, 0), 70.0), dist=1.0)
    dat = F.write_dat(str(tmp_path / "vol.dat"), s.field, slice_thickness=s.slice_dist)
    with open(dat, "a") as f:
        f.write("TimeDependent: 0 0\n")
    kpng = F.write_png(str(tmp_path / "kernel.png"), s.filter_row[None, :]) if s.filter_row is not None else None
    illum = "ILLUM_" in (s.defines or "")
    ref = refhost.renderer_state(dat, kpng, s.camera, s.light, s.lic_params(), s.lowres, illum, s.width, s.height)
    o = oracle.OracleScene(s)
    cam, rot = _view(oracle, o)
    # view: rotation block and camera position (= gl_ModelViewMatrixInverse[3]) of the frame's model-view
    mv = ref["modelview"].astype(np.float64).reshape(4, 4).T            # column-major -> [row][col]
    assert np.abs(mv[:3, :3] - rot.reshape(3, 3)).max() < 2e-6
    cam_ref = -mv[:3, :3].T @ mv[:3, 3]
    assert np.abs(cam_ref - cam).max() < 5e-6
    # light position
    lp = np.zeros(4, np.float32)
    oracle.lib().vvo_light_position(ctypes.byref(o.c), oracle._p(lp))
    assert np.abs(ref["light_position"][:3] - lp[:3]).max() < 5e-6 and ref["light_position"][3] == 1.0
    # slicing set-up through Renderer::updateSlices (model-view built by the reference's own GL calls)
    v_ref, d_ref, n_ref = ref["slicing"]
    v, d, n = o.slicing_setup()
    assert n == n_ref and abs(d - d_ref) < 1e-6 and np.abs(np.asarray(v) - v_ref).max() < 2e-6
    # parameter block
    un = o.uniforms()
    u = ref["uniforms"]
    assert u["stepSize"][0] == np.float32(un[0])
    assert np.array_equal(u["gradient"][:3], np.asarray(un[1:4], np.float32))
    assert np.array_equal(u["licParams"][:3], np.asarray(un[4:7], np.float32))
    assert np.array_equal(u["licKernel"][:3], np.asarray(un[7:10], np.float32))
    assert u["alphaCorrection"][0] == np.float32(un[10]) and int(u["numIterations"][0]) == int(un[11])
    sc = np.zeros(9, np.float32)
    oracle.lib().vvo_scale_uniforms(ctypes.byref(o.c), oracle._p(sc))
    assert np.array_equal(u["scaleVol"][:3], sc[0:3])                    # Q1: holds scaleInv in ILLUM_* builds
    assert np.array_equal(u["texMax"][:3], sc[6:9])
    assert not u["scaleVolInv"].any()                                    # the scaleVolInv slot itself is never written (Q1)
    rw, rh = (max(s.width // 2, 1), max(s.height // 2, 1)) if s.lowres else (s.width, s.height)
    assert list(u["viewport"]) == [0, 0, rw, rh]


def test_proxy_cube_faces(oracle, tmp_path):
    """Renderer::drawCubeFaces (VV/renderer.cpp:682-736): the proxy geometry is the box [0, extent]^3, every vertex carries its
    own position as texcoord0 (so a fragment's gl_TexCoord[0] is the point where the pixel ray enters the box -- the oracle's
    pixel_ray), and every quad is wound counter-clockwise seen from outside (front faces survive glEnable(GL_CULL_FACE))"""
    from vectorvisualization_b200 import configs, fields as F
    s = configs.cfg1(n=12, size=16)
    s.field = np.ascontiguousarray(F.abc_flow(24)[::2, :16, :])            # 24 x 16 x 12, anisotropic spacing
    s.slice_dist = (1.0, 1.5, 2.0)
    dat = F.write_dat(str(tmp_path / "vol.dat"), s.field, slice_thickness=s.slice_dist)
    with open(dat, "a") as f:
        f.write("TimeDependent: 0 0\n")
    verts, tex = refhost.cube_faces(dat)
    assert np.array_equal(verts, tex)
    o = oracle.OracleScene(s)
    ext = np.array(list(o.c.extent), np.float32)
    assert set(np.unique(verts[:, 0])) == {0.0, ext[0]} and set(np.unique(verts[:, 1])) == {0.0, ext[1]} and set(np.unique(verts[:, 2])) == {0.0, ext[2]}
    centre = ext.astype(np.float64) / 2
    for q in range(6):
        P = verts[4 * q: 4 * q + 4].astype(np.float64)
        newell = sum(np.cross(P[i], P[(i + 1) % 4]) for i in range(4))
        outward = P.mean(axis=0) - centre
        assert np.dot(newell, outward) > 0
        # planar, on a face of the box
        axis = int(np.argmax(np.abs(outward)))
        assert len(set(P[:, axis])) == 1


@pytest.mark.skipif(not refhost.has_app(), reason="oracle/_ref built without VV/3DLIC.cpp")
def test_key_map_matches_the_application(vv, tmp_path):
    """vv_key_apply against keyboard() / keyboardSpecial() of VV/3DLIC.cpp:243-488, compiled unmodified and called directly
    (oracle/ref_app_driver.cpp): every key on its own, every key repeated until its clamp bites, and random key sequences;
    LICParams compared bit for bit, technique / flags / clip-plane selection / shader defines exactly"""
    from vectorvisualization_b200 import configs, fields as F
    s = configs.cfg1(n=8, size=16)
    dat = F.write_dat(str(tmp_path / "vol.dat"), s.field, slice_thickness=s.slice_dist)
    with open(dat, "a") as f:
        f.write("TimeDependent: 0 0\n")
    if not os.path.isdir("/root/reference/VectorVisualization/shader"):
        pytest.skip("the shader sources the reference's key handler re-loads are not on this box")

    def ours(keys):
        st = vv.AppState()
        acts = [vv.keyApply(st, k, sp) for k, sp in keys]
        return st, acts

    def check(keys):
        ref = refhost.app_keyboard(dat, keys)
        st, acts = ours(keys)
        lic = st.lic
        got = (np.float32(lic.stepSizeVol), np.float32(lic.gradientScale), np.float32(lic.illumScale), np.float32(lic.freqScale),
               lic.numIterations, lic.stepsForward, lic.stepsBackward, np.float32(lic.stepSizeLIC))
        assert all(np.asarray(a).tobytes() == np.asarray(b).tobytes() for a, b in zip(got, ref["lic"])), (keys, got, ref["lic"])
        assert st.technique == ref["technique"], keys
        assert (st.lowres, st.float_target, st.recording, st.animation, st.screenshot, st.continuous) == \
            (ref["lowres"], ref["fbo"], ref["recording"], ref["animation"], ref["screenshot"], ref["continuous"]), keys
        assert ref["store_frame"] == 1 - st.continuous
        assert list(st.clip_active) == ref["clip_active"] and st.selected_clip == ref["selected_clip"], keys
        reloads = sum(1 for a in acts if a & vv.KEY_RELOAD_SHADER)
        assert ref["shader_loads"] == 16 * reloads, keys          # 8 programs x (vertex + fragment) per Renderer::loadGLSLShader
        if reloads:
            assert st.defines.decode() == ref["defines"], keys
        return st, acts, ref

    plain = list("0RHrwtFpLI[]sxSXazhnjmgbu1234567899. ") + ["y", "Z"]
    for k in plain:
        check([(k, 0)])
    for f in range(1, 7):
        check([(f, 1)])
    for k in "[]xXznmb":                                   # down (or up) to the clamp and beyond
        check([(k, 0)] * 40)
    for k in "sSahjg":
        check([(k, 0)] * 12)
    # what the application does next (VV/3DLIC.cpp:450-455, 457-488)
    st, acts, ref = check([(4, 1), ("s", 0), (3, 1), ("h", 0), (2, 1), ("h", 0), ("H", 0)])
    assert acts[0] == vv.KEY_SET_TECHNIQUE | vv.KEY_UPDATE_SCENE | vv.KEY_UPDATE_LICVOLUME
    assert acts[1] == vv.KEY_UPDATE_SCENE | vv.KEY_UPDATE_LICVOLUME
    assert acts[2] == vv.KEY_SET_TECHNIQUE | vv.KEY_UPDATE_SCENE | vv.KEY_UPDATE_SLICES
    assert acts[3] == vv.KEY_UPDATE_SCENE | vv.KEY_UPDATE_SLICES
    assert acts[5] == vv.KEY_UPDATE_SCENE and acts[6] == 0
    assert vv.keyApply(st, "q") == vv.KEY_QUIT and vv.keyApply(st, 27) == vv.KEY_QUIT      # exit(1) there (not fed to the reference)
    assert "Raycast" in ref["hud"] and "Freqency Scale: 1.4" in ref["hud"]
    rng = np.random.RandomState(11)
    alphabet = [(k, 0) for k in "0RrFpL[]sxSXazhnjmgbu12346789. "] + [(f, 1) for f in range(1, 6)]
    for _ in range(25):
        check([alphabet[i] for i in rng.randint(0, len(alphabet), size=30)])


@pytest.mark.skipif(not refhost.has_app(), reason="oracle/_ref built without VV/3DLIC.cpp")
def test_animation_cursor_matches_the_application(vv, tmp_path):
    """VVTimeCursor against the animation as VV/3DLIC.cpp drives it: init() (createTextureIterp + checkInterpolateStage once) and
    then the same pair once per idle() tick, on a time-dependent DAT file -- current time step, fraction index of every tick's
    texture, and the ticks at which the pair of time steps moves on (wrapping from the last step to the first)"""
    from vectorvisualization_b200 import fields as F
    steps = [F.abc_flow(6) * np.float32(1 + 0.3 * t) for t in range(3)]
    dat = F.write_dat(str(tmp_path / "anim.dat"), None, time_steps=steps)
    ref, _ = refhost.animation_ticks(dat, 47)
    c = vv.TimeCursor(0, 2, 10)
    c.tick()                                           # what init() has done before the first idle tick
    ours = []
    for _ in range(47):
        cur = c.current
        used, adv = c.tick()
        ours.append((cur, used, int(adv)))
    assert ours == ref
    assert [t for t, (_, _, adv) in enumerate(ref) if adv] == [8, 18, 28, 38]
    assert ref[0] == (0, 1, 0) and ref[9] == (1, 0, 0) and ref[29] == (0, 0, 0)        # first tick: 1/10; wraps 2 -> 0


@pytest.mark.parametrize("seed", range(6))
def test_random_host_preprocessing(oracle, tmp_path, seed):
    """seeded random inputs through the reference's host code (compiled unmodified) and the oracle: vector packing with time
    interpolation on non-cubic, anisotropically spaced fields with huge / tiny / zero vectors; noise gradients (Sobel, 5^3
    smoothing, quantisation) on non-cubic random noise; filter kernels of random width -- bit for bit"""
    from vectorvisualization_b200 import fields as F
    rng = np.random.RandomState(40 + seed)
    nx, ny, nz = (int(v) for v in rng.randint(3, 14, size=3))
    scale = np.float32(10.0 ** rng.uniform(-6, 6))
    f0 = (rng.standard_normal((nz, ny, nx, 3)) * scale).astype(np.float32)
    f1 = (rng.standard_normal((nz, ny, nx, 3)) * scale).astype(np.float32)
    f0[rng.rand(nz, ny, nx) < 0.1] = 0.0
    f1[rng.rand(nz, ny, nx) < 0.1] = 0.0
    f0[0, 0, 0] = (1e-6 * scale, 0.0, 0.0)                                       # around the |v| < 1e-5 zero-vector test
    sd = tuple(float(v) for v in rng.choice([0.5, 1.0, 1.5, 3.0], size=3))
    dat = F.write_dat(str(tmp_path / "vec.dat"), None, time_steps=[f0, f1], slice_thickness=sd)
    idx = int(rng.randint(0, 10))
    ref, geom, _, _ = refhost.vector_texture(dat, (nz, ny, nx), interp=(idx, 10))
    mine = oracle.pack_vector_field(f0, f1, (idx, 10), fp16=False)
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
    ext, sc, sci, cen = (np.zeros(3, np.float32) for _ in range(4))
    oracle.lib().vvo_volume_geometry((ctypes.c_int * 3)(nx, ny, nz), (ctypes.c_float * 3)(*sd), oracle._p(ext), oracle._p(sc),
                                     oracle._p(sci), oracle._p(cen))
    assert np.array_equal(ext, geom["extent"]) and np.array_equal(cen, geom["center"])
    assert np.array_equal(sc, geom["scale"][:3]) and np.array_equal(sci, geom["scale_inv"][:3])
    # noise gradients
    shape = tuple(int(v) for v in rng.randint(5, 13, size=3))
    noise = rng.randint(0, 256, size=shape).astype(np.uint8) if seed % 2 else (rng.rand(*shape) < 0.2).astype(np.uint8) * 255
    path = F.write_noise(str(tmp_path / "noise"), noise)
    ref, _, _ = refhost.noise_texture(path, shape, True)
    assert np.array_equal(oracle.pack_noise_rgba(noise, oracle.noise_gradients(noise)), ref)
    # filter kernel of random width and content
    width = int(rng.randint(2, 300))
    row = rng.randint(0, 256, size=width).astype(np.uint8)
    row[rng.randint(0, width)] = 200
    png = F.write_png(str(tmp_path / "k.png"), row[None, :])
    data, inv, _ = refhost.filter_texture(png)
    mine, minv = oracle.filter_from_row(row)
    assert np.array_equal(data, mine) and inv == minv


def test_scalar_volume_loader(vv, oracle, tmp_path):
    """VolumeDataSet::loadData + createTexture (VV/dataset.cpp:840-1050): the scalar volume goes to the GL as it is on disk -- UCHAR
    or FLOAT source, GL_LUMINANCE, LINEAR, CLAMP_TO_EDGE -- and the product's DAT / RAW reader hands over the same bytes; a FLOAT
    source is clamped to [0, 1] and stored as UNORM8 by the GL (2.1 spec 3.6.4 / 2.14.9), which is what vv_set_scalar's conversion
    kernel does on the GPU (test_preprocess_bit_exact)."""
    from vectorvisualization_b200 import fields as F
    rng = np.random.RandomState(9)
    u8 = rng.randint(0, 256, size=(5, 7, 9)).astype(np.uint8)
    f32 = rng.uniform(-0.2, 1.2, size=(6, 4, 8)).astype(np.float32)
    for vol, gl_type in ((u8, 0x1401), (f32, 0x1406)):                # GL_UNSIGNED_BYTE, GL_FLOAT
        dat = F.write_dat(str(tmp_path / ("s%d.dat" % vol.itemsize)), vol)
        with open(dat, "a") as f:
            f.write("TimeDependent: 0 0\n")
        ref, ifmt, wrap, typ, linear = refhost.scalar_texture(dat, vol.shape, vol.dtype)
        assert ifmt == 0x1909 and wrap == refhost.GL_CLAMP_TO_EDGE and typ == gl_type and linear      # GL_LUMINANCE
        assert np.array_equal(ref.view(np.uint8), vol.view(np.uint8))
        info = vv.parse_dat(dat)
        assert info.data_dim == 1 and tuple(info.resolution) == vol.shape[::-1]
        buf = np.zeros_like(vol)
        assert vv.load_library().vv_read_raw(ctypes.byref(info), 0, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes) == 0
        assert np.array_equal(buf.view(np.uint8), ref.view(np.uint8))


def test_sampler_state_of_every_texture(oracle, tmp_path):
    """filter and wrap mode of every texture on the hot path, read back from the reference's own glTexParameteri calls
    (loaders and Renderer compiled unmodified) and compared with what oracle, shim and CUDA samplers implement:
      vector field      RGBA16F        LINEAR   CLAMP_TO_EDGE          (fetch_field: edge-replicated padding)
      scalar volume     LUMINANCE      LINEAR   CLAMP_TO_EDGE          (fetch_scalar)
      noise             LUMINANCE/RGBA LINEAR   REPEAT                 (fetch_noise_*: wrapped border)
      filter kernel     LUMINANCE      LINEAR   GL_CLAMP (border 0)    (kernel_lookup, Q10)
      TF rgba / alpha-opacity          LINEAR   CLAMP_TO_EDGE          (tf_lookup / opac_lookup)
      Zoeckler / Mallo tables (2D)     LINEAR   CLAMP_TO_EDGE
      LIC volume buffer RGBA16F        LINEAR   REPEAT                 (fetch_licvol, Q14)
      FBO colour targets RGBA16F rect  NEAREST  CLAMP_TO_EDGE          (texture2DRect: imageFBOSampler)
      MC offsets        LUMINANCE16F rect NEAREST CLAMP_TO_EDGE        (mc_offset)"""
    from vectorvisualization_b200 import fields as F
    GL = dict(NEAREST=0x2600, LINEAR=0x2601, CLAMP=0x2900, REPEAT=0x2901, CLAMP_TO_EDGE=0x812F, TEX1D=0x0DE0, TEX2D=0x0DE1, TEX3D=0x806F,
              RECT=0x84F5, RGBA16F=0x881A, LUMINANCE=0x1909, RGBA=0x1908, LA=0x190A, L16F=0x881E)
    f = F.abc_flow(6)
    noise = np.random.RandomState(1).randint(0, 256, size=(5, 5, 5)).astype(np.uint8)
    row = F.filter_kernel("cos2", 64)

    def one(fn):
        refhost.upload_log_reset()
        fn()
        return refhost.uploaded_textures()

    dat = F.write_dat(str(tmp_path / "v.dat"), f)
    with open(dat, "a") as fh:
        fh.write("TimeDependent: 0 0\n")
    t = one(lambda: refhost.vector_texture(dat, (6, 6, 6)))[-1]
    assert (t["target"], t["internal_format"], t["min_filter"], t["mag_filter"]) == (GL["TEX3D"], GL["RGBA16F"], GL["LINEAR"], GL["LINEAR"])
    assert (t["wrap_s"], t["wrap_t"], t["wrap_r"]) == (GL["CLAMP_TO_EDGE"],) * 3
    sdat = F.write_dat(str(tmp_path / "s.dat"), noise)
    with open(sdat, "a") as fh:
        fh.write("TimeDependent: 0 0\n")
    t = one(lambda: refhost.scalar_texture(sdat, noise.shape, np.uint8))[-1]
    assert (t["target"], t["internal_format"], t["min_filter"], t["mag_filter"]) == (GL["TEX3D"], GL["LUMINANCE"], GL["LINEAR"], GL["LINEAR"])
    assert (t["wrap_s"], t["wrap_t"], t["wrap_r"]) == (GL["CLAMP_TO_EDGE"],) * 3
    npath = F.write_noise(str(tmp_path / "noise"), noise)
    for grad, ifmt in ((False, GL["LUMINANCE"]), (True, GL["RGBA"])):
        t = one(lambda: refhost.noise_texture(npath, noise.shape, grad))[-1]
        assert (t["target"], t["internal_format"], t["min_filter"], t["mag_filter"]) == (GL["TEX3D"], ifmt, GL["LINEAR"], GL["LINEAR"])
        assert (t["wrap_s"], t["wrap_t"], t["wrap_r"]) == (GL["REPEAT"],) * 3
    png = F.write_png(str(tmp_path / "k.png"), row[None, :])
    t = one(lambda: refhost.filter_texture(png))[-1]
    assert (t["target"], t["min_filter"], t["mag_filter"], t["wrap_s"]) == (GL["TEX1D"], GL["LINEAR"], GL["LINEAR"], GL["CLAMP"])
    ts = one(lambda: refhost.tf_textures(None))
    assert len(ts) >= 2
    for t, ifmt in zip(ts[-2:], (GL["RGBA"], GL["LA"])):
        assert (t["target"], t["internal_format"], t["min_filter"], t["mag_filter"], t["wrap_s"]) == \
            (GL["TEX1D"], ifmt, GL["LINEAR"], GL["LINEAR"], GL["CLAMP_TO_EDGE"]) and t["dims"][0] == 256
    ts = one(lambda: refhost.illum_tables())
    assert len(ts) == 3
    for t in ts:
        assert (t["target"], t["min_filter"], t["mag_filter"], t["wrap_s"], t["wrap_t"]) == \
            (GL["TEX2D"], GL["LINEAR"], GL["LINEAR"], GL["CLAMP_TO_EDGE"], GL["CLAMP_TO_EDGE"]) and t["dims"][:2] == (256, 256)
    ts = one(lambda: refhost.renderer_textures(40, 30, (8, 6, 4)))
    fbo = [t for t in ts if t["target"] == GL["RECT"] and t["internal_format"] == GL["RGBA16F"]]
    mc = [t for t in ts if t["target"] == GL["RECT"] and t["internal_format"] == GL["L16F"]]
    lic = [t for t in ts if t["target"] == GL["TEX3D"]]
    assert len(fbo) == 2 and len(mc) == 1 and len(lic) == 2
    for t in fbo:
        assert (t["min_filter"], t["mag_filter"], t["wrap_s"], t["wrap_t"], t["dims"][:2]) == (GL["NEAREST"], GL["NEAREST"], GL["CLAMP_TO_EDGE"], GL["CLAMP_TO_EDGE"], (40, 30))
    assert (mc[0]["min_filter"], mc[0]["mag_filter"], mc[0]["wrap_s"], mc[0]["dims"][:2]) == (GL["NEAREST"], GL["NEAREST"], GL["CLAMP_TO_EDGE"], (40, 30))
    for t in lic:
        assert (t["internal_format"], t["min_filter"], t["mag_filter"], t["dims"]) == (GL["RGBA16F"], GL["LINEAR"], GL["LINEAR"], (8, 6, 4))
        assert (t["wrap_s"], t["wrap_t"], t["wrap_r"]) == (GL["REPEAT"],) * 3
