"""The oracle (fp32 C++ restatement) against the REFERENCE'S OWN SHADER SOURCE compiled as C++ (oracle/_ref, built by
oracle/build_ref.py from /root/reference/VectorVisualization/shader/*.glsl).  Runs wherever oracle/_ref/libvv_ref.so
exists (the build container; the .so also travels to the GPU box).  Expectation: bit-identical frames and ray-sample
counts -- both sides are single-operation IEEE fp32 in the order the shader source writes the arithmetic."""
import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libvv_ref.so not built (needs /root/reference)")


def _scenes():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from scenes import golden_scenes
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    S = dict(golden_scenes())

    def lum():
        s = configs.cfg1(n=16, size=32)
        s.quirk_luminance_alpha = 1          # GL_LUMINANCE noise: .a == 1 (Q7)
        s.params.update(gradientScale=1.0)
        return s
    S["quirk_luminance_alpha"] = lum

    def lowres():
        s = configs.cfg2(n=20, size=32)
        s.lowres = 1
        return s
    S["lowres"] = lowres

    def q1():
        s = configs.cfg3(n=16, size=32, camera=F.CAMERA_CLOSE)
        s.field = np.ascontiguousarray(F.tornado(24)[:12, :20, :])
        s.slice_dist = (1.0, 1.5, 2.0)        # anisotropic: scale != scaleInv, Q1 visible
        s.tf_mode = vv.TF_B
        return s
    S["q1_anisotropic_gradient"] = q1
    return S


@pytest.mark.parametrize("name", sorted(_scenes().keys()))
def test_raycast_bit_exact(oracle, name):
    s = _scenes()[name]()
    a, ca, ta = oracle.OracleScene(s).raycast()
    b, cb, tb = refshim.RefScene(s).raycast()
    assert ta == tb and np.array_equal(ca, cb)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff %g" % np.abs(a - b).max()


@pytest.mark.parametrize("define", ["", "ILLUM_GRADIENT", "ILLUM_MALLO"])
def test_slicing_bit_exact(oracle, define):
    """VOLIC_SLICING: lic3d_slicing_fragment.glsl run slice by slice on the frame buffer (VV/renderer.cpp:1123-1267)"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    tables = oracle.illum_tables(40.0)
    for mk in (lambda: configs.cfg3(n=20, size=32, camera=F.CAMERA_CLOSE), lambda: configs.cfg2(n=20, size=36)):
        s = mk()
        s.defines = ("#define " + define) if define else ""
        s.with_gradients = True
        s.technique = vv.VOLIC_SLICING
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA         # what the slicing shader hard-codes
        s.tf = F.default_tf()
        s.params.update(gradientScale=4.0)
        a, ca, ta = oracle.OracleScene(s, illum_tables=tables).slicing()
        b, cb, tb = refshim.RefScene(s, illum_tables=tables).slicing()
        assert ta == tb and ta > 0 and np.array_equal(ca, cb)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff %g" % np.abs(a - b).max()


@pytest.mark.parametrize("define", ["ILLUM_MALLO", "ILLUM_ZOECKLER"])
def test_illuminated_streamlines_bit_exact(oracle, define):
    """Mallo / Zoeckler builds (keys 8 / 7, VV/3DLIC.cpp:416-427) with the tables of VV/illumination.cpp"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    tables = oracle.illum_tables(40.0)
    for mk in (lambda: configs.cfg2(n=20, size=36, camera=F.CAMERA_CLOSE), lambda: configs.cfg1(n=16, size=32)):
        s = mk()
        s.defines = "#define " + define
        s.params.update(gradientScale=4.0, illumScale=1.3)
        s.light = dict(quat=F.quat_from_axis_angle((0.2, 1, 0), 70.0), dist=1.0)
        a, ca, ta = oracle.OracleScene(s, illum_tables=tables).raycast()
        b, cb, tb = refshim.RefScene(s, illum_tables=tables).raycast()
        assert ta == tb and ta > 0 and np.array_equal(ca, cb)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff %g" % np.abs(a - b).max()


@pytest.mark.parametrize("name", ["cfg2_close_gs6", "cfg3_gradient_length", "anisotropic_tf_scalar_band", "q1_anisotropic_gradient"])
def test_lic_volume_and_volume_raycast_bit_exact(oracle, name):
    s = _scenes()[name]()
    s.licvol_fp16 = 0
    o, r = oracle.OracleScene(s), refshim.RefScene(s)
    lo, lr = o.lic_volume((10, 12, 14)), r.lic_volume((10, 12, 14))
    assert np.array_equal(lo.view(np.uint32), lr.view(np.uint32))
    a, ca, ta = o.raycast_licvolume(lo)
    b, cb, tb = r.raycast_licvolume(lo)
    assert ta == tb and np.array_equal(ca, cb)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_mc_offset_builds_bit_exact(oracle):
    """USE_MC_OFFSET programs (inc_header.glsl:14; lic3d_fragment.glsl:31-33, lic3d_slicing_fragment.glsl:31-33): the ray /
    fragment start is jittered by dir * stepSize * mcOffset[pixel]"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs
    def plain():
        return configs.cfg1(n=24, size=48)
    def grad():
        s = configs.cfg3(n=24, size=40)
        s.tf_mode = vv.TF_B
        return s
    def slicing():
        s = configs.cfg1(n=24, size=40)
        s.technique = vv.VOLIC_SLICING
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        return s
    for i, mk in enumerate((plain, grad, slicing)):
        s = mk()
        base = oracle.OracleScene(s)
        base = base.slicing() if s.technique == vv.VOLIC_SLICING else base.raycast()
        s.defines = (s.defines or "") + "\n#define USE_MC_OFFSET"
        s.mc_offsets = np.random.RandomState(3 + i).rand(s.height, s.width).astype(np.float32)
        o, r = oracle.OracleScene(s), refshim.RefScene(s)
        a, ca, ta = o.slicing() if s.technique == vv.VOLIC_SLICING else o.raycast()
        b, cb, tb = r.slicing() if s.technique == vv.VOLIC_SLICING else r.raycast()
        assert ta == tb and ta > 0 and np.array_equal(ca, cb)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "max abs diff %g" % np.abs(a - b).max()
        assert not np.array_equal(a, base[0])                      # the offsets do move the samples


def test_display_pass_bit_exact(oracle):
    """background_fragment.glsl (VV/shader, compiled as C++) against the oracle's display pass: frame-sized window, and the low-res
    preset's window of twice (or 2n+1 times) the frame, where texture2DRect(.., gl_FragCoord.xy * viewport.xy) up-scales NEAREST"""
    rng = np.random.RandomState(3)
    for (rw, rh), (ww, wh) in (((13, 9), (13, 9)), ((32, 24), (64, 48)), ((25, 20), (51, 41)), ((1, 1), (3, 2)), ((40, 30), (81, 61))):
        img = rng.rand(rh, rw, 4).astype(np.float32)
        img[..., :3] *= img[..., 3:4]                       # premultiplied, as the ray-cast leaves it
        img[0, 0] = (0, 0, 0, 0); img[-1, -1] = (1, 1, 1, 1)
        ref = refshim.background(img, ww, wh)
        got = oracle.display_window(img, ww, wh)
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
        if (rw, rh) == (ww, wh):
            assert np.array_equal(oracle.background(img).view(np.uint32), ref.view(np.uint32))


def _random_scene(seed):
    """a random small scene: everything the parameter block and the textures can vary in"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    rng = np.random.RandomState(seed)
    nx, ny, nz = (int(v) for v in rng.randint(6, 15, size=3))
    field = rng.standard_normal((nz, ny, nx, 3)).astype(np.float32)
    field[rng.rand(nz, ny, nx) < 0.05] = 0.0                                     # zero vectors (rgb = 0.5, a = 0)
    nn = int(rng.randint(5, 12))
    noise = (rng.rand(nn, nn, nn) < rng.choice([1 / 6, 0.5])).astype(np.uint8) * 255
    if rng.rand() < 0.3:
        noise = rng.randint(0, 256, size=(nn, nn, nn)).astype(np.uint8)
    sn = int(rng.randint(3, 9))
    scalar = rng.randint(20, 90, size=(sn, sn, sn)).astype(np.uint8)             # straddles the (0.1, 0.3) band
    kw = int(rng.choice([7, 16, 100, 256]))
    filt = None if rng.rand() < 0.2 else rng.randint(0, 256, size=kw).astype(np.uint8)
    if filt is not None:
        filt[rng.randint(0, kw)] = 255                                           # never an all-zero kernel
    tf = rng.randint(0, 256, size=(256, 5)).astype(np.uint8)
    tf[:, 3] = (tf[:, 3] * rng.choice([0.1, 0.4, 1.0])).astype(np.uint8)         # semi-transparent ... opaque (early termination)
    size = int(rng.randint(14, 26))
    illum = rng.choice(["", "ILLUM_GRADIENT", "SPEED_OF_FLOW", "ILLUM_MALLO", "ILLUM_ZOECKLER"], p=[0.3, 0.3, 0.2, 0.1, 0.1])
    axis = rng.standard_normal(3)
    cam = dict(quat=F.quat_from_axis_angle(tuple(axis), float(rng.uniform(0, 360))), pos=tuple(float(v) for v in rng.uniform(-0.2, 0.2, 3)),
               dist=float(rng.uniform(1.2, 4.0)), fovy=35.0)
    s = configs.Scene("random%d" % seed, field, noise, scalar, filt, tf, size, int(size * rng.choice([1.0, 0.75])),
                      with_gradients=bool(illum == "ILLUM_GRADIENT" or rng.rand() < 0.5), defines=("#define " + illum) if illum else "",
                      params=dict(stepSizeVol=float(rng.choice([1 / 16, 1 / 32, 1 / 64])), stepSizeLIC=float(rng.choice([0.005, 0.01, 0.04])),
                                  stepsForward=int(rng.randint(1, 12)), stepsBackward=int(rng.randint(1, 12)),
                                  freqScale=float(rng.choice([0.5, 1.0, 2.2, 4.0])), gradientScale=float(rng.uniform(1, 30)),
                                  illumScale=float(rng.uniform(0.5, 1.5))),
                      camera=cam, light=dict(quat=F.quat_from_axis_angle(tuple(rng.standard_normal(3)), float(rng.uniform(0, 360))), dist=float(rng.uniform(0.5, 2))))
    s.slice_dist = tuple(float(v) for v in rng.choice([1.0, 1.5, 2.0], size=3))
    s.lowres = int(rng.rand() < 0.15)
    if illum in ("", "ILLUM_GRADIENT", "SPEED_OF_FLOW"):
        s.tf_mode = int(rng.choice([vv.TF_B, vv.TF_A, vv.TF_R, vv.TF_LENGTH, vv.TF_SCALAR]))
    else:
        s.tf_mode = int(rng.choice([vv.TF_B, vv.TF_A, vv.TF_R, vv.TF_LENGTH, vv.TF_SCALAR]))
    if illum == "" and s.tf_mode in (vv.TF_B, vv.TF_A) and rng.rand() < 0.3:
        s.gate_mode = vv.GATE_TF_ALPHA                                           # built for raycast_none (.b / .a)
    if rng.rand() < 0.3:
        n = rng.standard_normal(3)
        s.clip_planes = ((float(n[0]), float(n[1]), float(n[2]), float(rng.uniform(-0.1, 0.3))),)   # any length: the steady state applies
    if rng.rand() < 0.2:
        s.camera = dict(cam, dist=float(rng.uniform(0.55, 0.9)))                  # near the box: the near plane clips
    return s


@pytest.mark.parametrize("block", range(4))
def test_random_scenes_bit_exact(oracle, block):
    """seeded random scenes -- random field (incl. zero vectors, anisotropic spacing, non-cubic), noise, scalar volume, filter
    kernel, transfer function, LIC parameters, camera, light, build, TF index, gate, clip plane, near-plane positions: the oracle
    against the reference's shader code, frames and ray-sample counts bit for bit"""
    tables = oracle.illum_tables(40.0)
    hits = 0
    for seed in range(100 + 8 * block, 108 + 8 * block):
        s = _random_scene(seed)
        need = "MALLO" in s.defines or "ZOECKLER" in s.defines
        a, ca, ta = oracle.OracleScene(s, illum_tables=tables if need else None).raycast()
        b, cb, tb = refshim.RefScene(s, illum_tables=tables if need else None).raycast()
        assert ta == tb and np.array_equal(ca, cb), seed
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "seed %d: max abs diff %g" % (seed, np.abs(a - b).max())
        hits += int(ta > 0)
    assert hits >= 5


def test_random_scenes_slicing_and_lic_volume_bit_exact(oracle):
    """the same random scenes through the slicing program (TF index .a, gate tfData.a > 0.05 as the shader hard-codes them) and
    through the LIC-volume program + the ray-cast of the LIC volume"""
    import vectorvisualization_b200 as vv
    tables = oracle.illum_tables(40.0)
    done = 0
    for seed in range(200, 216):
        s = _random_scene(seed)
        if "SPEED_OF_FLOW" in s.defines:
            s.defines = ""                                                        # no slicing build with that define in the shim
        need = "MALLO" in s.defines or "ZOECKLER" in s.defines
        s.technique, s.tf_mode, s.gate_mode = vv.VOLIC_SLICING, vv.TF_A, vv.GATE_TF_ALPHA
        s.with_gradients = True
        a, ca, ta = oracle.OracleScene(s, illum_tables=tables if need else None).slicing()
        b, cb, tb = refshim.RefScene(s, illum_tables=tables if need else None).slicing()
        assert ta == tb and np.array_equal(ca, cb), seed
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "seed %d: max abs diff %g" % (seed, np.abs(a - b).max())
        done += int(ta > 0)
        if need:
            continue
        s.technique = vv.VOLIC_RAYCAST
        s.licvol_fp16 = 0
        o, r = oracle.OracleScene(s), refshim.RefScene(s)
        dims = (5 + seed % 4, 6, 4 + seed % 3)
        lo, lr = o.lic_volume(dims), r.lic_volume(dims)
        assert np.array_equal(lo.view(np.uint32), lr.view(np.uint32)), seed
        a, ca, ta = o.raycast_licvolume(lo)
        b, cb, tb = r.raycast_licvolume(lo)
        assert ta == tb and np.array_equal(ca, cb) and np.array_equal(a.view(np.uint32), b.view(np.uint32)), seed
    assert done >= 8


@pytest.mark.parametrize("define", ["", "ILLUM_GRADIENT", "ILLUM_MALLO"])
def test_slicing_without_fbo_fragments_bit_exact(oracle, define):
    """The slicing variant the reference runs until key 'F' switches the FBO on (Renderer::sliceVolume, VV/renderer.cpp:1150-1160):
    lic3d_slicingblend_fragment.glsl returns every fragment's premultiplied sample and the GL blends it into the RGBA8 back
    buffer.  The oracle's fragment colours against that shader's code, bit for bit (after the GL's clamp to [0, 1]); the 8-bit
    blend of vvo_slicing_blend8 against the same arithmetic done here from the reference shader's outputs."""
    import ctypes
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    tables = oracle.illum_tables(40.0)
    s = configs.cfg3(n=16, size=20, camera=F.CAMERA_CLOSE) if define else configs.cfg2(n=16, size=20)
    s.defines = ("#define " + define) if define else ""
    s.with_gradients = True
    s.technique, s.tf_mode, s.gate_mode = vv.VOLIC_SLICING, vv.TF_A, vv.GATE_TF_ALPHA
    s.tf = F.default_tf()
    s.params.update(gradientScale=4.0, stepSizeVol=1 / 32)
    need = "MALLO" in define
    o = oracle.OracleScene(s, illum_tables=tables if need else None)
    r = refshim.RefScene(s, illum_tables=tables if need else None)
    _, _, nslices = o.slicing_setup()
    frame = np.zeros((s.height, s.width, 4), np.uint8)
    buf = np.zeros((nslices, 4), np.float32)
    shaded = 0
    for y in range(s.height):
        for x in range(s.width):
            n = oracle.lib().vvo_slice_fragments(ctypes.byref(o.c), x, y, oracle._p(buf), nslices)
            mine = o.slice_fragment_colors(x, y)
            assert len(mine) == n
            dst = np.zeros(4, np.uint8)
            if n:
                ref = np.clip(r.slicing_blend_fragments(buf[:n, :3].copy()), 0.0, 1.0)
                assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), (x, y)
                shaded += int((ref != 0).any(axis=1).sum())
                for src in ref:                                          # glBlendFunc(GL_ONE_MINUS_DST_ALPHA, GL_ONE) on RGBA8
                    da = np.float32(dst[3]) / np.float32(255.0)
                    v = src * (np.float32(1.0) - da) + dst.astype(np.float32) / np.float32(255.0)
                    dst = np.floor(np.clip(v, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
            da = np.float32(dst[3]) / np.float32(255.0)                  # the white plane, VV/renderer.cpp:1238-1255
            v = np.float32(1.0) * (np.float32(1.0) - da) + dst.astype(np.float32) / np.float32(255.0)
            frame[y, x] = np.floor(np.clip(v, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    got, cnt, tot = o.slicing_blend8()
    assert np.array_equal(got, frame) and tot == shaded and tot > 200
    assert (got[..., 3] == 255).all()                                    # the white plane completes the alpha
    # against the FBO variant (one fp32 accumulator) composited over white: same picture up to the 8-bit accumulation and the 0.95 skip
    fbo, _, _ = oracle.OracleScene(s, illum_tables=tables if need else None, fbo_fp16=0, fbo_pingpong=0).slicing()
    disp = oracle.quantize_rgba8(oracle.background(fbo))
    assert np.abs(disp[..., :3].astype(np.int32) - got[..., :3].astype(np.int32)).max() <= 24
