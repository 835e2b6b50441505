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
