"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances are the north star's: <= 2/255 max per-channel RGBA and PSNR >= 45 dB on the stored RGBA8 frame,
ray-sample counts exactly equal, LIC volume <= 1e-4 relative, pre-processing bit-exact.
"""
import numpy as np
import pytest

from util import MAX_DIFF_8BIT, MIN_PSNR_DB, assert_image_parity, render_cuda, LICVOL_REL

pytestmark = pytest.mark.gpu


def _scenes():
    from vectorvisualization_b200 import configs, fields as F
    import vectorvisualization_b200 as vv
    S = {}
    S["cfg1_small"] = lambda: configs.cfg1(n=64, size=160)
    S["cfg1_close"] = lambda: configs.cfg1(n=32, size=96, camera=F.CAMERA_CLOSE)
    S["cfg2_small"] = lambda: configs.cfg2(n=64, size=128)
    S["cfg3_small"] = lambda: configs.cfg3(n=64, size=128)
    S["cfg3_close"] = lambda: configs.cfg3(n=48, size=96, camera=F.CAMERA_CLOSE)
    S["cfg4_small"] = lambda: configs.cfg4(n=64, size=96, noise_n=32)

    def lowcontrast():
        s = configs.cfg1(n=48, size=128)
        s.params.update(gradientScale=3.0)
        return s
    S["cfg1_gradscale3"] = lowcontrast

    def lowcontrast3():
        s = configs.cfg3(n=48, size=112)
        s.params.update(gradientScale=4.0, illumScale=1.2)
        return s
    S["cfg3_gradscale4"] = lowcontrast3

    def gate_tf():
        s = configs.cfg2(n=48, size=96)
        s.gate_mode = vv.GATE_TF_ALPHA
        s.tf_mode = vv.TF_A
        s.params.update(gradientScale=5.0)
        return s
    S["gate_tf_alpha_tf_a"] = gate_tf

    def tf_r():
        s = configs.cfg1(n=32, size=80)
        s.tf_mode = vv.TF_R
        s.params.update(gradientScale=6.0)
        return s
    S["tf_r"] = tf_r

    def tf_scalar():
        s = configs.cfg1(n=32, size=80)
        s.tf_mode = vv.TF_SCALAR
        rng = np.random.RandomState(5)
        s.scalar = rng.randint(30, 90, size=(16, 16, 16)).astype(np.uint8)   # band (0.1,0.3) = 26..76: partly gated
        s.params.update(gradientScale=6.0)
        return s
    S["tf_scalar_band_gate"] = tf_scalar

    def lowres():
        s = configs.cfg2(n=48, size=96)
        s.lowres = 1
        return s
    S["lowres_preset"] = lowres

    def sof():
        s = configs.cfg2(n=48, size=96)
        s.defines = "#define SPEED_OF_FLOW"
        s.params.update(gradientScale=6.0)
        return s
    S["speed_of_flow"] = sof

    def nogate():
        s = configs.cfg1(n=32, size=80)
        s.noise_gate = 0
        s.scalar = None
        return s
    S["noise_gate_off"] = nogate

    def lum():
        s = configs.cfg1(n=32, size=80)
        s.quirk_luminance_alpha = 1
        s.params.update(gradientScale=1.0)
        return s
    S["quirk_luminance_alpha"] = lum

    def aniso():
        s = configs.cfg1(n=32, size=112, camera=F.CAMERA_CLOSE)
        s.field = np.ascontiguousarray(F.abc_flow(64)[::2, :48, ::1][:24])   # 64 x 48 x 24 voxels (x,y,z)
        s.slice_dist = (1.0, 1.0, 2.0)
        s.params.update(gradientScale=6.0)
        return s
    S["anisotropic_noncubic"] = aniso

    def aniso_grad():
        s = configs.cfg3(n=32, size=96, camera=F.CAMERA_CLOSE)
        s.field = np.ascontiguousarray(F.tornado(48)[:24, :40, :])          # 48 x 40 x 24
        s.slice_dist = (1.0, 1.5, 2.0)
        return s
    S["anisotropic_gradient_q1"] = aniso_grad

    def moved_cam():
        cam = dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.15, -0.1, 0.3), dist=3.0, fovy=35.0)
        s = configs.cfg3(n=40, size=104, camera=cam)
        s.light = dict(quat=F.quat_from_axis_angle((1, 0, 0), 40.0), dist=1.0)
        return s
    S["rotated_translated_camera_light"] = moved_cam

    def steps():
        s = configs.cfg2(n=40, size=96)
        s.params.update(stepsForward=25, stepsBackward=40, stepSizeLIC=0.02, gradientScale=8.0)
        return s
    S["steps_25_40"] = steps

    def onestep():
        s = configs.cfg1(n=32, size=64)
        s.params.update(stepsForward=1, stepsBackward=1)
        return s
    S["steps_1_1"] = onestep

    def timeinterp():
        s = configs.cfg1(n=32, size=80)
        s.next_field = F.rankine_vortex(32)
        s.interp = (3, 10)
        return s
    S["time_interpolation"] = timeinterp

    def odd_size():
        s = configs.cfg1(n=32, size=64, camera=F.CAMERA_CLOSE)
        s.width, s.height = 75, 41
        return s
    S["odd_image_size"] = odd_size

    def fine_step():
        s = configs.cfg1(n=32, size=72)
        s.params.update(stepSizeVol=1.0 / 256.0)
        return s
    S["raycast_step_256"] = fine_step
    return S


SCENES = list(_scenes().keys()) if True else []


@pytest.mark.parametrize("name", SCENES)
def test_raycast_parity(vv, oracle, name):
    scene = _scenes()[name]()
    o = oracle.OracleScene(scene)
    ref, ref_cnt, ref_tot = o.raycast()
    r, img, img8, cnt, tot = render_cuda(vv, scene)
    # ray-sample counts must match exactly unless an early-termination decision (src.a > 0.95) flipped
    mism = int((cnt != ref_cnt).sum())
    assert mism <= max(1, cnt.size // 20000), "%s: %d pixels with different ray-sample counts" % (name, mism)
    if mism == 0:
        assert tot == ref_tot
    md, ps, mf = assert_image_parity(oracle, img, ref, name)
    assert np.array_equal(img8, oracle.quantize_rgba8(img))     # the library's RGBA8 store == GL conversion
    assert ref_tot > 0 or "inside" in name
    print("%s: samples %d, max8 %d, psnr %.1f dB, float max diff %.3g" % (name, tot, md, ps, mf))


@pytest.mark.parametrize("define", ["ILLUM_MALLO", "ILLUM_ZOECKLER"])
def test_illuminated_streamlines_parity(vv, oracle, define):
    """Mallo / Zoeckler illuminated-streamline builds"""
    from vectorvisualization_b200 import configs, fields as F
    tables = oracle.illum_tables(40.0)
    for mk in (lambda: configs.cfg2(n=48, size=112, camera=F.CAMERA_CLOSE), lambda: configs.cfg1(n=32, size=80)):
        s = mk()
        s.defines = "#define " + define
        s.params.update(gradientScale=4.0, illumScale=1.3)
        s.light = dict(quat=F.quat_from_axis_angle((0.2, 1, 0), 70.0), dist=1.0)
        ref, ref_cnt, ref_tot = oracle.OracleScene(s, illum_tables=tables).raycast()
        _, img, _, cnt, tot = render_cuda(vv, s)
        assert int((cnt != ref_cnt).sum()) <= 1
        assert_image_parity(oracle, img, ref, define)


@pytest.mark.parametrize("define", ["", "ILLUM_GRADIENT"])
def test_slicing_parity(vv, oracle, define):
    """VOLIC_SLICING (F3): view-aligned slices, TF index .a, gate tfData.a > 0.05, stop on dest.a >= 0.95"""
    from vectorvisualization_b200 import configs, fields as F
    for mk in (lambda: configs.cfg3(n=48, size=112, camera=F.CAMERA_CLOSE), lambda: configs.cfg2(n=40, size=96),
               lambda: configs.cfg1(n=32, size=75, camera=dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0, 0.2), dist=3.0, fovy=35.0))):
        s = mk()
        s.defines = ("#define " + define) if define else ""
        s.with_gradients = True
        s.technique = vv.VOLIC_SLICING
        s.fbo = 1
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        s.tf = F.default_tf()
        s.params.update(gradientScale=4.0)
        ref, ref_cnt, ref_tot = oracle.OracleScene(s).slicing()
        _, img, _, cnt, tot = render_cuda(vv, s)
        assert ref_tot > 0
        assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000), "%d pixels differ in fragment count" % int((cnt != ref_cnt).sum())
        assert_image_parity(oracle, img, ref, "slicing " + define)


@pytest.mark.parametrize("define", ["", "ILLUM_GRADIENT"])
def test_slicing_without_fbo_parity(vv, oracle, define):
    """VOLIC_SLICING as the reference starts up (no FBO until key 'F'): lic3d_slicingblend_fragment.glsl, every fragment blended with
    (ONE_MINUS_DST_ALPHA, ONE) into the RGBA8 back buffer -- rounded to 8 bits per slice, no dest.a skip -- and the white plane
    at the end (VV/renderer.cpp:1151-1160, 1236-1262).  The oracle's fragment colours are bit-identical to that shader."""
    from util import psnr8
    from vectorvisualization_b200 import configs, fields as F
    for mk in (lambda: configs.cfg3(n=48, size=112, camera=F.CAMERA_CLOSE), lambda: configs.cfg2(n=40, size=96),
               lambda: configs.cfg1(n=32, size=75, camera=dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0, 0.2), dist=3.0, fovy=35.0))):
        s = mk()
        s.defines = ("#define " + define) if define else ""
        s.with_gradients = True
        s.technique = vv.VOLIC_SLICING
        s.fbo = 0
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        s.tf = F.default_tf()
        s.params.update(gradientScale=4.0)
        ref8, ref_cnt, ref_tot = oracle.OracleScene(s).slicing_blend8()
        r, img, img8, cnt, tot = render_cuda(vv, s)
        assert ref_tot > 0 and (img8[..., 3] == 255).all()                      # the white plane completes the alpha
        assert np.array_equal(img8, oracle.quantize_rgba8(img))                  # the float read-back is the byte frame / 255
        d = np.abs(img8.astype(np.int32) - ref8.astype(np.int32))
        print("slicing without FBO %s %s: %d fragments (oracle %d), max 8-bit diff %d, %d bytes differ, PSNR %.1f"
              % (s.name, define or "plain", tot, ref_tot, d.max(), int((d > 0).sum()), psnr8(img8, ref8)))
        assert d.max() <= MAX_DIFF_8BIT and psnr8(img8, ref8) >= MIN_PSNR_DB
        assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000) and abs(tot - ref_tot) <= max(1, ref_tot // 20000)
        # the displayed frame: background_fragment.glsl over a frame whose alpha is 1 changes nothing
        assert np.array_equal(r.readDisplayRGBA8(), img8)


def test_mc_offset_parity(vv, oracle):
    """USE_MC_OFFSET builds: jittered ray starts for the ray-cast programs and the slicing program"""
    from vectorvisualization_b200 import configs, fields as F
    def plain():
        return configs.cfg1(n=32, size=80)
    def grad():
        return configs.cfg3(n=40, size=96, camera=F.CAMERA_CLOSE)
    def slicing():
        s = configs.cfg2(n=40, size=75)
        s.technique = vv.VOLIC_SLICING
        s.fbo = 1
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        s.tf = F.default_tf()
        return s
    for i, mk in enumerate((plain, grad, slicing)):
        s = mk()
        _, base, _, _, _ = render_cuda(vv, s)
        s.defines = (s.defines or "") + "\n#define USE_MC_OFFSET"
        s.mc_offsets = np.random.RandomState(11 + i).rand(s.height, s.width).astype(np.float32)
        o = oracle.OracleScene(s)
        ref, ref_cnt, ref_tot = o.slicing() if s.technique == vv.VOLIC_SLICING else o.raycast()
        r, img, _, cnt, tot = render_cuda(vv, s)
        assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000)
        assert_image_parity(oracle, img, ref, "mc offset %d" % i)
        assert not np.array_equal(img, base)
        # per-ray kernel (VV_OPT_RAYCAST_MODE 0) applies the same offsets
        if s.technique != vv.VOLIC_SLICING:
            r.setOption(vv.OPT_RAYCAST_MODE, 0)
            r.render(True)
            # (the two kernels inline the shading code separately: fused-multiply-add contraction may differ by an ulp)
            assert np.abs(r.readRGBA32F() - img).max() <= 2e-5
        # a define without a texture is the plain program (an unbound sampler reads 0)
        r.setMCOffsets(None)
        r.setOption(vv.OPT_RAYCAST_MODE, 1)
        r.render(True)
        assert np.array_equal(r.readRGBA32F(), base)
    # the generator fills a frame-sized texture with values in [0,1]; a wrong size is refused at render time
    r = vv.Renderer(0)
    s = plain()
    s.defines = "#define USE_MC_OFFSET"
    configs.apply_scene(r, s)
    r.updateMCOffsetTex(s.width, s.height, 5)
    r.render(True)
    assert not np.array_equal(r.readRGBA32F(), render_cuda(vv, plain())[1])
    r.updateMCOffsetTex(s.width + 1, s.height, 5)
    with pytest.raises(vv.VVError):
        r.render(True)


@pytest.mark.parametrize("planes", [((0.0, 0.0, -1.0, 0.1),), ((0.0, 0.0, 1.0, 0.1),), ((0.3, 0.5, -0.8, 0.05), (1.0, 0.0, 0.0, 0.12)),
                                    ((0.0, 0.0, -1.0, 0.2), (0.6, -0.8, 0.0, 0.1), (0.0, 1.0, 0.0, 0.3))])
def test_clip_planes_parity(vv, oracle, planes):
    """user clip planes: clipped front faces + cap polygons for the ray-cast techniques, clipped slices"""
    from vectorvisualization_b200 import configs, fields as F
    from vectorvisualization_b200.configs import apply_scene
    cam = dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 40.0), pos=(0.05, 0.0, 0.1), dist=3.0, fovy=35.0)
    for mk in (lambda: configs.cfg1(n=32, size=96), lambda: configs.cfg3(n=40, size=88, camera=cam)):
        s = mk()
        _, base, _, base_cnt, _ = render_cuda(vv, s)
        s.clip_planes = planes
        ref, ref_cnt, ref_tot = oracle.OracleScene(s).raycast()
        r, img, _, cnt, tot = render_cuda(vv, s)
        assert np.array_equal(cnt > 0, ref_cnt > 0)                          # same set of covered pixels
        assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000)
        assert_image_parity(oracle, img, ref, "clip raycast")
        # (a plane whose kept side contains the whole front of the box changes nothing: rays still leave through the box)
        assert planes[0][2] > 0 or not np.array_equal(cnt, base_cnt)
        r.setOption(vv.OPT_RAYCAST_MODE, 0)
        r.render(True)
        assert np.abs(r.readRGBA32F() - img).max() <= 2e-5
        # deactivating the planes restores the unclipped frame
        for i in range(3):
            r.setClipPlane(i, None, False)
        r.setOption(vv.OPT_RAYCAST_MODE, 1)
        r.render(True)
        assert np.array_equal(r.readRGBA32F(), base)
    # slicing: the slice polygons are clipped
    s = configs.cfg2(n=40, size=75)
    s.technique = vv.VOLIC_SLICING
    s.fbo = 1
    s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
    s.tf = F.default_tf()
    s.clip_planes = planes
    ref, ref_cnt, ref_tot = oracle.OracleScene(s).slicing()
    _, img, _, cnt, tot = render_cuda(vv, s)
    assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000)
    assert_image_parity(oracle, img, ref, "clip slicing")
    # LIC-volume technique: the plain ray-cast over the precomputed volume starts at the same entry points
    s = configs.cfg1(n=32, size=96)
    s.technique = vv.VOLIC_LICVOLUME
    s.params.update(gradientScale=4.0)
    s.clip_planes = planes
    o = oracle.OracleScene(s)
    r = vv.Renderer(0)
    apply_scene(r, s)
    r.setOption(vv.OPT_SAMPLE_MAP, 1)
    r.render(True)
    ref, ref_cnt, _ = o.raycast_licvolume(r.readLICVolume())
    assert int((r.readSampleMap() != ref_cnt).sum()) <= 1
    assert_image_parity(oracle, r.readRGBA32F(), ref, "clip volume raycast")


def test_slicing_default_modes(vv):
    """selecting VOLIC_SLICING switches to the slicing program's own TF index / gate (Q5, Q6) without touching the ray-cast ones"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg2(n=24, size=48)
    s.params.update(gradientScale=4.0)
    r = vv.Renderer(0)
    apply_scene(r, s)                      # sets ray-cast options (B, always)
    r.render(True)
    a = r.readRGBA32F().copy()
    r.setTechnique(vv.VOLIC_SLICING)       # slicing defaults: TF .a, gate tf.a > 0.05
    r.render(True)
    b = r.readRGBA32F().copy()
    r.setTechnique(vv.VOLIC_RAYCAST)
    r.render(True)
    assert np.array_equal(r.readRGBA32F(), a) and not np.array_equal(a, b) and b.any()


def test_layouts_bit_identical(vv):
    """float4, x-pair fp16 and xy-quad fp16 layouts hold the same RGBA16F values -> identical frames, field read-backs and LIC volumes"""
    from vectorvisualization_b200 import configs
    for s in (configs.cfg3(n=48, size=96), configs.cfg1(n=40, size=80), configs.cfg4(n=48, size=96, noise_n=32)):
        ra, a, _, ca, ta = render_cuda(vv, s, layout=vv.LAYOUT_PAIR)
        _, b, _, cb, tb = render_cuda(vv, s, layout=vv.LAYOUT_F4)
        rc, c, _, cc, tc = render_cuda(vv, s, layout=vv.LAYOUT_QUAD)
        assert ta == tb == tc and np.array_equal(ca, cb) and np.array_equal(ca, cc)
        assert np.array_equal(a, b) and np.array_equal(a, c)
        assert np.array_equal(ra.readFieldTexture(s.field.shape[:3]), rc.readFieldTexture(s.field.shape[:3]))
        assert (ra.fieldLayout(), rc.fieldLayout()) == (vv.LAYOUT_PAIR, vv.LAYOUT_QUAD)
        # the default (VV_OPT_FIELD_LAYOUT = 3) resolves to the xy-quad layout at these sizes; one-thread-per-ray kernel on it too
        rd, d, _, cd, td = render_cuda(vv, s)
        assert rd.fieldLayout() == vv.LAYOUT_QUAD and td == ta and np.array_equal(d, a) and np.array_equal(cd, ca)
        rd.setOption(vv.OPT_RAYCAST_MODE, 0)
        rd.render(True)
        assert np.array_equal(rd.readRGBA32F(), a)
        ra.setOption(vv.OPT_LICVOL_FP16, 0); rc.setOption(vv.OPT_LICVOL_FP16, 0)
        ra.updateLICVolume(); rc.updateLICVolume()
        assert np.array_equal(ra.readLICVolume(), rc.readLICVolume())


@pytest.mark.parametrize("name", ["cfg1_small", "cfg3_small", "gate_tf_alpha_tf_a", "odd_image_size"])
def test_raycast_modes_bit_identical(vv, name):
    """sample-parallel pipeline (default) == one-thread-per-ray kernel, bit for bit, incl. early termination"""
    from vectorvisualization_b200.configs import apply_scene
    scene = _scenes()[name]()
    out = []
    for mode in (1, 0):
        r = vv.Renderer(0)
        r.setOption(vv.OPT_RAYCAST_MODE, mode)
        apply_scene(r, scene)
        r.setOption(vv.OPT_SAMPLE_MAP, 1)
        r.render(True)
        out.append((r.readRGBA32F(), r.readSampleMap(), r.lastRaySamples()))
    assert out[0][2] == out[1][2] and out[0][2] > 0
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], out[1][0])


def test_camera_inside_box_draws_nothing(vv, oracle):
    from vectorvisualization_b200 import configs
    s = configs.cfg1(n=32, size=64)
    s.camera = dict(quat=(0, 0, 0, 1), pos=(0, 0, 0), dist=0.2, fovy=35.0)   # back faces are culled (VV/renderer.cpp:1099)
    ref, _, ref_tot = oracle.OracleScene(s).raycast()
    _, img, _, _, tot = render_cuda(vv, s)
    assert ref_tot == 0 and tot == 0
    assert not img.any() and not ref.any()


def test_render_without_update_keeps_frame(vv):
    from vectorvisualization_b200 import configs
    s = configs.cfg1(n=32, size=64)
    r, img, _, _, _ = render_cuda(vv, s)
    r.render(False)    # Renderer::render(false) re-presents the stored frame (VV/renderer.cpp:150-152,228)
    assert r.lastLaunchCount() > 0
    assert np.array_equal(r.readRGBA32F(), img)
    disp = r.readDisplayRGBA8()
    assert disp.shape == (64, 64, 4)


def test_display_background(vv, oracle):
    from vectorvisualization_b200 import configs
    s = configs.cfg1(n=32, size=64)
    r, img, _, _, _ = render_cuda(vv, s)
    want = oracle.quantize_rgba8(oracle.background(img))      # background_fragment.glsl:9-16
    assert np.array_equal(r.readDisplayRGBA8(), want)


def test_preprocess_bit_exact(vv, oracle):
    """GPU packing / noise-gradient kernels reproduce the reference's host loops bit for bit"""
    from vectorvisualization_b200 import fields as F
    field = F.tornado(40)
    nxt = F.abc_flow(40)
    noise = F.white_noise(40, 11, F.SPARSE_P)
    r = vv.Renderer(0)
    r.setVectorField(field, nxt)
    r.setTimeInterp(4, 10)
    got = r.readFieldTexture((40, 40, 40))
    want = oracle.pack_vector_field(field, nxt, (4, 10), fp16=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    r.setOption(vv.OPT_FIELD_LAYOUT, vv.LAYOUT_F4)
    got = r.readFieldTexture((40, 40, 40))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # zero vectors -> rgb 0.5 (VV/dataset.cpp:596-602)
    z = np.zeros((8, 8, 8, 3), np.float32)
    z[1:] = field[:7, :8, :8]
    r.setVectorField(z)
    assert np.array_equal(r.readFieldTexture((8, 8, 8)).view(np.uint32), oracle.pack_vector_field(z).view(np.uint32))
    # UCHAR3
    u = np.random.RandomState(3).randint(0, 256, size=(12, 10, 14, 3)).astype(np.uint8)
    r.setVectorField(u)
    assert np.array_equal(r.readFieldTexture((12, 10, 14)).view(np.uint32), oracle.pack_vector_field(u).view(np.uint32))
    # noise gradients (-g): RGBA8 = (quantised smoothed Sobel gradient, noise)
    r.setNoise(noise, True)
    got = r.readNoiseTexture((40, 40, 40), 4)
    want = oracle.pack_noise_rgba(noise, oracle.noise_gradients(noise))
    assert np.array_equal(got, want)
    # ragged, non-cubic noise
    n2 = np.random.RandomState(4).randint(0, 256, size=(9, 17, 12)).astype(np.uint8)
    r.setNoise(n2, True)
    assert np.array_equal(r.readNoiseTexture((9, 17, 12), 4), oracle.pack_noise_rgba(n2, oracle.noise_gradients(n2)))
    # built-in white noise == mt19937 definition
    r.generateWhiteNoise(24, seed=9, p=F.SPARSE_P)
    assert np.array_equal(r.readNoiseTexture((24, 24, 24), 1), F.white_noise(24, 9, F.SPARSE_P))


@pytest.mark.parametrize("grad", [False, True])
def test_lic_volume_parity(vv, oracle, grad):
    """LIC-volume mode vs the fp32 transcription of the shader integrator: <= 1e-4 relative"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg3(n=32, size=64) if grad else configs.cfg2(n=32, size=64)
    s.licvol_fp16 = 0
    s.technique = vv.VOLIC_LICVOLUME
    want = oracle.OracleScene(s).lic_volume()
    r = vv.Renderer(0)
    apply_scene(r, s)
    r.updateLICVolume()
    got = r.readLICVolume()
    assert got.shape == want.shape
    scale = np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    rel = np.abs(got - want) / scale
    print("LIC volume max rel err %.3g (grad=%s)" % (rel.max(), grad))
    assert rel.max() <= LICVOL_REL
    # fp16 target (Q14): identical after the same rounding, up to one fp16 ulp where the fp32 values straddle a tie
    r.setOption(vv.OPT_LICVOL_FP16, 1)
    r.updateLICVolume()
    got16 = r.readLICVolume()
    want16 = want.astype(np.float16).astype(np.float32)
    assert np.abs(got16 - want16).max() <= np.abs(want).max() * 2.0 ** -10


def test_lic_volume_resolution_and_slabs(vv, oracle):
    """target resolution != field resolution, computed in two z-slabs == computed at once"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg2(n=24, size=64)
    s.licvol_fp16 = 0
    s.licvol_size = 40
    r = vv.Renderer(0)
    apply_scene(r, s)
    r.updateLICVolume()
    full = r.readLICVolume().copy()
    assert full.shape == (40, 40, 40)
    want = oracle.OracleScene(s).lic_volume((40, 40, 40))
    assert (np.abs(full - want) / np.maximum(np.abs(want), 1e-3 * want.max())).max() <= LICVOL_REL
    r2 = vv.Renderer(0)
    apply_scene(r2, s)
    r2.setLICVolumeSlab(0, 17)
    r2.updateLICVolume()
    a = r2.readLICVolume().copy()
    r2.setLICVolumeSlab(17, 40)
    r2.updateLICVolume()
    b = r2.readLICVolume()
    assert np.array_equal(b[17:], full[17:]) and np.array_equal(a[:17], full[:17])


def test_volume_raycast_parity(vv, oracle):
    """raycast_lic3d_fragment.glsl: plain ray-cast over the precomputed LIC volume"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    for mk in (lambda: configs.cfg2(n=32, size=128), lambda: configs.cfg1(n=32, size=96)):
        s = mk()
        s.technique = vv.VOLIC_LICVOLUME
        s.params.update(gradientScale=4.0)
        o = oracle.OracleScene(s)
        lv = o.lic_volume()
        r = vv.Renderer(0)
        apply_scene(r, s)
        r.setOption(vv.OPT_SAMPLE_MAP, 1)
        r.render(True)
        got_lv = r.readLICVolume()
        # ray-cast the oracle against the CUDA LIC volume so that the comparison isolates the ray-cast stage
        ref, ref_cnt, ref_tot = o.raycast_licvolume(got_lv)
        img = r.readRGBA32F()
        cnt = r.readSampleMap()
        assert int((cnt != ref_cnt).sum()) <= 1
        assert_image_parity(oracle, img, ref, "volume_raycast")
        # and end to end (oracle LIC volume)
        ref2, _, _ = o.raycast_licvolume(lv)
        assert_image_parity(oracle, img, ref2, "volume_raycast_e2e")


@pytest.mark.parametrize("unit", [1, 2, 4])
def test_partition_invariance_single_gpu(vv, unit):
    """sort-first block partition: 1 handle == 3 partitioned handles assembled (bit-identical), for single blocks and for units of
    2 x 2 / 4 x 4 blocks (VV_OPT_PARTITION_UNIT; the 90 x 70 frame has 6 x 5 blocks, so border units stick out of the image)"""
    import torch
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    from vectorvisualization_b200.dist import device_tensor
    s = configs.cfg1(n=32, size=64, camera=None)
    s.width, s.height = 90, 70
    _, full, _, _, tot = render_cuda(vv, s, sample_map=False)
    world = 3
    hs = []
    total = 0
    for rank in range(world):
        r = vv.Renderer(0)
        r.setOption(vv.OPT_PARTITION_UNIT, unit)
        apply_scene(r, s)
        r.setPartition(rank, world)
        r.render(True)
        r.synchronize()
        total += r.lastRaySamples()
        hs.append(r)
    ptr, bpr, nb = hs[0].tileBuffer()
    gathered = torch.empty((world, bpr, 256, 4), dtype=torch.float32, device="cuda")
    for rank, r in enumerate(hs):
        p, b, _ = r.tileBuffer()
        gathered[rank].copy_(device_tensor(p, (b, 256, 4), torch.float32))
    torch.cuda.synchronize()
    hs[0].assembleTiles(gathered.data_ptr(), world)
    img = hs[0].readRGBA32F()
    assert total == tot
    assert np.array_equal(img, full)


@pytest.mark.parametrize("unit", [1, 2])
def test_p2p_exchange_single_process(vv, unit):
    """vv_p2p_*: three partitioned handles in one process exchange their tiles through peer stores + arrival counters
    (the multi-GPU path without IPC): every handle ends up with the unpartitioned frame, over several frames (buffer
    parity / arrival targets) and a camera change"""
    from vectorvisualization_b200 import configs, fields as F
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg3(n=32, size=64)
    s.width, s.height = 90, 70
    world = 3
    hs = []
    for rank in range(world):
        r = vv.Renderer(0)
        r.setOption(vv.OPT_PARTITION_UNIT, unit)
        apply_scene(r, s)
        r.setPartition(rank, world)
        # warm-up with both views: every buffer reaches its final size.  (Ranks normally live in separate processes; here
        # they share one, where a cudaFree / cudaMalloc of one handle would wait for the other handle's spinning stream.)
        for cam in (dict(F.CAMERA_CLOSE), s.camera):
            r.setCamera(**cam)
            r.render(True)
            r.synchronize()
        hs.append(r)
    bases = [r.p2pExport()[1] for r in hs]
    for r in hs:
        r.p2pConnect(local_bases=bases)
    cams = [s.camera, dict(F.CAMERA_CLOSE), s.camera, dict(F.CAMERA_CLOSE), s.camera]
    for cam in cams:
        s.camera = cam
        _, want, _, _, _ = render_cuda(vv, s, sample_map=False)
        for r in hs:
            r.setCamera(**cam)
            r.updateLightPos()             # the light is placed relative to the camera (VV/renderer.cpp:431-466)
        for r in hs:
            r.p2pRender()                  # asynchronous: rank 0 waits on the device for ranks 1 and 2
        for r in hs:
            r.p2pStatus()
            assert np.array_equal(r.readRGBA32F(), want)
    for r in hs:
        r.p2pDisconnect()


def test_full_size_cfg3(vv, oracle):
    """BASELINE.json configs[1] at full size (256^3 field, 256^3 noise + gradients, 1024^2): the oracle shades sampled pixel
    patches of the full frame (direct parity), and the whole frame is checked through size-independent properties:
    ray-sample total == sum of the per-pixel map == the analytic chord count the bench reports, idempotence, layout
    invariance, and invariance under an 8-way sort-first partition (bit-identical assembly)"""
    import torch
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    from vectorvisualization_b200.dist import device_tensor
    s = configs.cfg3()
    assert s.field.shape[:3] == (256, 256, 256) and (s.width, s.height) == (1024, 1024)
    r, img, img8, cnt, tot = render_cuda(vv, s)
    assert tot == int(cnt.sum()) == 21660568
    assert np.array_equal(img8, oracle.quantize_rgba8(img))
    o = oracle.OracleScene(s)
    rng = np.random.RandomState(0)
    ys, xs = np.nonzero(cnt > 0)
    worst = 0
    for k in rng.choice(len(ys), 24, replace=False):
        x0, y0 = max(0, int(xs[k]) - 2), max(0, int(ys[k]) - 2)
        rect = (x0, y0, min(s.width, x0 + 5), min(s.height, y0 + 5))
        ref, ref_cnt, _ = o.raycast(rect=rect)
        sl = (slice(rect[1], rect[3]), slice(rect[0], rect[2]))
        assert np.array_equal(cnt[sl], ref_cnt[sl])
        a, b = oracle.quantize_rgba8(img[sl]), oracle.quantize_rgba8(ref[sl])
        worst = max(worst, int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max()))
    assert worst <= MAX_DIFF_8BIT
    # idempotence
    r.render(True)
    assert np.array_equal(r.readRGBA32F(), img)
    # float4 check layout
    r.setOption(vv.OPT_FIELD_LAYOUT, vv.LAYOUT_F4)
    r.render(True)
    assert np.array_equal(r.readRGBA32F(), img)
    del r
    # 8-way sort-first partition, assembled: bit-identical frame, same ray-sample total, balanced shares
    world = 8
    h = vv.Renderer(0)
    apply_scene(h, s)
    gathered, shares = None, []
    for rank in range(world):
        h.setPartition(rank, world)
        h.render(True)
        h.synchronize()
        shares.append(h.lastRaySamples())
        ptr, bpr, _ = h.tileBuffer()
        if gathered is None:
            gathered = torch.empty((world, bpr, 256, 4), dtype=torch.float32, device="cuda")
        gathered[rank].copy_(device_tensor(ptr, (bpr, 256, 4), torch.float32))
    torch.cuda.synchronize()
    h.assembleTiles(gathered.data_ptr(), world)
    assert np.array_equal(h.readRGBA32F(), img)
    assert sum(shares) == tot
    assert max(shares) <= 1.02 * tot / world, shares            # the rotated block ids keep the ranks within 2 %
    print("full-size cfg3: worst 8-bit diff on sampled patches %d, shares %s" % (worst, shares))


def test_full_size_cfg4(vv, oracle):
    """BASELINE.json configs[3] at full size (512^3 field, 2048^2 view, step 1/256): sampled-patch parity against the oracle,
    ray-sample accounting, idempotence"""
    from vectorvisualization_b200 import configs
    s = configs.cfg4()
    assert s.field.shape[:3] == (512, 512, 512) and (s.width, s.height) == (2048, 2048)
    r, img, img8, cnt, tot = render_cuda(vv, s)
    assert tot == int(cnt.sum()) == 172805416
    o = oracle.OracleScene(s)
    rng = np.random.RandomState(1)
    ys, xs = np.nonzero(cnt > 0)
    worst = 0
    for k in rng.choice(len(ys), 12, replace=False):
        x0, y0 = max(0, int(xs[k]) - 1), max(0, int(ys[k]) - 1)
        rect = (x0, y0, min(s.width, x0 + 3), min(s.height, y0 + 3))
        ref, ref_cnt, _ = o.raycast(rect=rect)
        sl = (slice(rect[1], rect[3]), slice(rect[0], rect[2]))
        assert np.array_equal(cnt[sl], ref_cnt[sl])
        a, b = oracle.quantize_rgba8(img[sl]), oracle.quantize_rgba8(ref[sl])
        worst = max(worst, int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max()))
    assert worst <= MAX_DIFF_8BIT
    r.render(True)
    assert np.array_equal(r.readRGBA32F(), img)
    print("full-size cfg4: worst 8-bit diff on sampled patches %d" % worst)


def test_file_loaders_roundtrip(vv, oracle, tmp_path):
    """DAT/RAW + noise file + PNG kernel + PNG TF through the reference's formats == in-memory path"""
    from vectorvisualization_b200 import configs, fields as F
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg2(n=32, size=64)
    s.tf = F.tf_preset("tf-length")
    _, want, _, _, _ = render_cuda(vv, s, sample_map=False)
    dat = F.write_dat(str(tmp_path / "vol.dat"), s.field)
    sdat = F.write_dat(str(tmp_path / "scalar.dat"), s.scalar)
    nz = F.write_noise(str(tmp_path / "noise_32"), s.noise)
    kpng = F.write_png(str(tmp_path / "kernel.png"), s.filter_row[None, :])
    tfn = F.write_tf(str(tmp_path / "tf.png"), s.tf)
    args = vv.parse_args(["volic", dat, "-n", nz, "-f", kpng, "--transfer=" + tfn])
    r = vv.Renderer(0)
    r.init(None)
    r.loadDat(args.vol_file.decode())
    r.loadNoise(args.noise_file.decode(), bool(args.use_gradients))
    r.loadScalarDat(sdat)
    r.loadFilterPNG(args.filter_file.decode())
    r.loadTF(args.tf_file.decode())
    r.setLICParams(s.lic_params())
    r.setCamera(**s.camera)
    r.resize(s.width, s.height)
    r.render(True)
    got = r.readRGBA32F()
    # loadTF reads only <name>_rgba.png when it exists (VV/transferEdit.cpp:239 short-circuit): LIC opacity stays default
    s2 = configs.cfg2(n=32, size=64)
    s2.tf = F.default_tf()
    s2.tf[:, :4] = s.tf[:, :4]
    _, want2, _, _, _ = render_cuda(vv, s2, sample_map=False)
    assert np.array_equal(got, want2)
    out = str(tmp_path / "frame.png")
    r.savePNG(out)
    back = vv.png_read(out)
    assert np.array_equal(back[::-1], r.readRGBA8())


def test_noise_gradient_cache(vv, oracle, tmp_path):
    """vv_load_noise(-g): computes the gradients on the GPU and stores <noise>.grd like NoiseDataSet::createTexture
    (VV/dataset.cpp:1238-1267); a second load uses the stored file (sentinel values prove it)"""
    import os
    from vectorvisualization_b200 import fields as F
    shape = (12, 10, 14)
    noise = np.random.RandomState(8).randint(0, 256, size=shape).astype(np.uint8)
    path = F.write_noise(str(tmp_path / "noise"), noise)
    r = vv.Renderer(0)
    r.loadNoise(path, True)
    want = oracle.noise_gradients(noise)
    assert os.path.getsize(path + ".grd") == 3 * noise.size
    assert np.array_equal(vv.grd_read(path, shape[::-1]), want)
    assert np.array_equal(r.readNoiseTexture(shape, 4), oracle.pack_noise_rgba(noise, want))
    sentinel = np.random.RandomState(9).randint(0, 256, size=shape + (3,)).astype(np.uint8)
    vv.grd_write(path, sentinel)
    r.loadNoise(path, True)
    tex = r.readNoiseTexture(shape, 4)
    assert np.array_equal(tex[..., :3], sentinel) and np.array_equal(tex[..., 3], noise)
    assert np.array_equal(r.readNoiseGradients(shape), sentinel)
    # the in-memory entry point gives the same texture
    r2 = vv.Renderer(0)
    r2.setNoiseWithGradients(noise, sentinel)
    assert np.array_equal(r2.readNoiseTexture(shape, 4), tex)


def test_screenshot_and_recording(vv, tmp_path):
    """Renderer::screenshot / switchRecording + the file naming of renderFBO (VV/renderer.cpp:1478-1513)"""
    import os, re
    from vectorvisualization_b200 import configs
    s = configs.cfg1(n=16, size=40)
    r = vv.Renderer(0)
    configs.apply_scene(r, s)
    out = str(tmp_path / "snapshotOut")
    os.makedirs(out)
    r.setSnapshot(out, "shot.png")
    r.render(True)
    assert os.listdir(out) == []                                   # nothing is written unless asked
    r.screenshot()
    r.render(False)                                                # a redisplay of the stored frame is enough
    files = os.listdir(out)
    assert len(files) == 1 and re.fullmatch(r"\d\d-\d\d-\d{4} \d\d-\d\d-\d\d shot\.png", files[0])
    assert r.lastSnapshotPath() == os.path.join(out, files[0])
    assert np.array_equal(vv.png_read(r.lastSnapshotPath())[::-1], r.readRGBA8())
    r.render(False)
    assert len(os.listdir(out)) == 1                               # one-shot
    assert r.switchRecording() is True
    for i in range(3):
        r.render(i % 2 == 0)
        assert r.lastSnapshotPath() == os.path.join(out, "%d_shot.png" % i)
    assert r.switchRecording() is False
    r.render(True)
    assert sorted(f for f in os.listdir(out) if f[0].isdigit() and "_" in f) == ["0_shot.png", "1_shot.png", "2_shot.png"]


def _golden_names():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from scenes import golden_scenes
    return sorted(golden_scenes().keys())


@pytest.mark.parametrize("name", _golden_names())
def test_cuda_matches_golden_vectors(vv, oracle, name):
    """CUDA path against tests/golden/*.npz -- outputs of the reference's own shader code (tests/golden/make_golden.py)"""
    import os
    from scenes import golden_scenes
    from vectorvisualization_b200.configs import apply_scene
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    s = golden_scenes()[name]()
    r, img, _, cnt, tot = render_cuda(vv, s)
    assert tot == int(g["total"]) and np.array_equal(cnt, g["raycast_samples"].astype(np.uint32))
    assert_image_parity(oracle, img, g["raycast"], name)
    # LIC volume (12^3 target) and the volume ray-cast over the full-resolution LIC volume
    s.licvol_fp16 = 0
    s.licvol_size = 12
    s.technique = vv.VOLIC_LICVOLUME
    r2 = vv.Renderer(0)
    apply_scene(r2, s)
    r2.updateLICVolume()
    lv = r2.readLICVolume()
    want = g["licvol12"]
    assert (np.abs(lv - want) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())).max() <= LICVOL_REL
    s.licvol_size = 0
    r3 = vv.Renderer(0)
    apply_scene(r3, s)
    r3.setOption(vv.OPT_SAMPLE_MAP, 1)
    r3.render(True)
    assert int((r3.readSampleMap() != g["volraycast_samples"].astype(np.uint32)).sum()) <= 1
    assert_image_parity(oracle, r3.readRGBA32F(), g["volraycast"], name + " volume ray-cast")


def _golden_extra_names():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from scenes import golden_extra_scenes
    return sorted(golden_extra_scenes().keys())


@pytest.mark.parametrize("name", _golden_extra_names())
def test_cuda_matches_golden_extra(vv, oracle, name):
    """CUDA path against the second golden family (slicing, Mallo / Zoeckler, USE_MC_OFFSET): reference shader outputs"""
    import os
    from scenes import golden_extra_scenes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    s = golden_extra_scenes()[name][0]()
    r, img, _, cnt, tot = render_cuda(vv, s)
    assert int((cnt != g["samples"].astype(np.uint32)).sum()) <= 1
    assert_image_parity(oracle, img, g["frame"], name)


def test_item_order_does_not_change_the_frame(vv):
    """depth-major (band, depth-chunk) work-item order vs tile-major order: bit-identical frames and counts"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    for mk in (lambda: configs.cfg3(n=48, size=200), lambda: configs.cfg1(n=32, size=150)):
        s = mk()
        out = []
        for dm, band in ((1, 4), (0, 4), (1, 1), (1, 64)):
            r = vv.Renderer(0)
            r.setOption(vv.OPT_DEPTH_MAJOR, dm)
            r.setOption(vv.OPT_BAND_ROWS, band)
            apply_scene(r, s)
            r.render(True)
            out.append((r.readRGBA32F(), r.lastRaySamples()))
        for img, n in out[1:]:
            assert n == out[0][1] and np.array_equal(img, out[0][0])


def test_noise_layouts_bit_identical(vv):
    """RGBA (-g) noise as bf16 {t0, t1 - t0} (PRMT widening, the hot layout), as fp16 x-pairs (FHADD lerps) and as u8 xy-quads
    (PRMT decode): same values, same operations on them, same frames"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg3(n=48, size=128)
    out = []
    for layout in (2, 1, 0):
        r = vv.Renderer(0)
        r.setOption(vv.OPT_NOISE_LAYOUT, layout)
        apply_scene(r, s)
        r.render(True)
        out.append(r.readRGBA32F())
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


def _golden_chain_names():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from scenes import golden_chain_scenes
    return sorted(golden_chain_scenes().keys())


@pytest.mark.gpu
@pytest.mark.parametrize("name", _golden_chain_names())
def test_cuda_matches_reference_chain(vv, oracle, name):
    """CUDA path against frames of the whole reference chain (tests/golden/y_chain_*.npz: the draw calls of Renderer::render(true)
    -> GL-spec rasteriser -> reference shader code, oracle/refchain.py), at the north-star tolerance; pixels whose coverage is
    implementation-defined (fragment centre on a polygon edge) are masked"""
    import os
    from scenes import golden_chain_scenes
    from util import psnr8, MIN_PSNR_DB
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    s = golden_chain_scenes()[name][0]()
    r, img, _, cnt, tot = render_cuda(vv, s)
    keep = ~g["edge"]
    assert (cnt[keep] != g["samples"].astype(np.uint32)[keep]).mean() <= 0.005
    a, b = oracle.quantize_rgba8(img)[keep], oracle.quantize_rgba8(g["frame"])[keep]
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= MAX_DIFF_8BIT and psnr8(a, b) >= MIN_PSNR_DB


@pytest.mark.gpu
@pytest.mark.parametrize("technique", ["raycast", "slicing"])
def test_near_plane_clips_the_proxy_geometry(vv, oracle, technique):
    """GL clips the cube faces / slice polygons to the view volume: where the entry point lies nearer than gluPerspective's
    near plane (0.1, VV/camera.cpp:42-47) the reference has no fragment -- a hole in the frame -- and slices nearer than it are
    dropped (tests/test_softgl_vs_ref.py pins this against the reference's own draw calls)"""
    from vectorvisualization_b200 import configs, fields as F
    cams = (dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 50.0), pos=(0.0, 0.0, 0.0), dist=0.75, fovy=35.0),      # near plane cuts the front face
            dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 30.0), pos=(0.0, 0.0, 0.0), dist=0.62, fovy=35.0))      # everything visible is nearer than 0.1
    for i, cam in enumerate(cams):
        s = configs.cfg1(n=32, size=96, camera=cam)
        if technique == "slicing":
            s.technique = vv.VOLIC_SLICING
            s.fbo = 1
            s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
            ref, ref_cnt, ref_tot = oracle.OracleScene(s).slicing()
        else:
            ref, ref_cnt, ref_tot = oracle.OracleScene(s).raycast()
        _, img, _, cnt, tot = render_cuda(vv, s)
        if technique == "raycast":
            assert tot == ref_tot and np.array_equal(cnt, ref_cnt)
            if i == 0:
                hit = ref_cnt > 0
                assert 0.02 < hit.mean() < 0.9
            else:
                assert ref_tot == 0 and not img.any()
        else:
            assert ref_tot > 0
            assert int((cnt != ref_cnt).sum()) <= max(1, cnt.size // 20000)
        assert_image_parity(oracle, img, ref, "near plane %s %d" % (technique, i))
        # a wider view volume brings the clipped part back
        s.camera = dict(cam, near=0.001)
        full_tot = oracle.OracleScene(s).raycast()[2] if technique == "raycast" else oracle.OracleScene(s).slicing()[2]
        _, _, _, _, tot2 = render_cuda(vv, s)
        assert abs(tot2 - full_tot) <= (0 if technique == "raycast" else max(2, full_tot // 10000)) and full_tot > ref_tot


def test_lowres_window_aspect(vv, oracle):
    """low-res preset as the reference renders it: a (w/2, h/2) frame with the projection of the w x h window
    (VV/renderer.cpp:111-119, VV/transform.h:79-80) -- vv_enable_lowres + vv_set_window + vv_resize"""
    from vectorvisualization_b200 import configs, fields as F
    s = configs.cfg3(n=32, size=25, camera=F.CAMERA_CLOSE)
    s.height = 20
    s.window = (51, 41)
    s.lowres = 1
    ref, ref_cnt, ref_tot = oracle.OracleScene(s).raycast()
    _, img, _, cnt, tot = render_cuda(vv, s)
    assert tot == ref_tot and tot > 0 and np.array_equal(cnt, ref_cnt)
    assert_image_parity(oracle, img, ref, "lowres window aspect")
    # the displayed window: background pass with the NEAREST up-scaling of the half-size frame (VV/renderer.cpp:1436-1441)
    r, img, _, _, _ = render_cuda(vv, s)
    want = oracle.quantize_rgba8(oracle.display_window(img, 51, 41))
    got = r.readDisplayWindowRGBA8(51, 41)
    assert got.shape == (41, 51, 4) and np.array_equal(r.readDisplayWindowRGBA8(25, 20), r.readDisplayRGBA8())
    assert np.array_equal(got, want)
    s.window = None
    _, img2, _, _, _ = render_cuda(vv, s)
    assert not np.array_equal(img, img2)


def test_keyboard_drives_the_renderer(vv):
    """vv_keyboard = the key map of VV/3DLIC.cpp (pinned on the CPU in test_key_map_matches_the_application) carried out on a
    handle: a frame after the keys equals the frame of a scene configured directly with the resulting parameters"""
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.configs import apply_scene
    s = configs.cfg3(n=32, size=64)
    r = vv.Renderer(0)
    apply_scene(r, s)
    st = vv.AppState()
    st.lic = s.lic_params()
    st.technique = vv.VOLIC_RAYCAST
    for key in "[sXL1":                       # sample distance / 2, one more forward step, one fewer backward step, low-res, clip plane 1
        act = r.keyboard(st, key)
        assert act & vv.KEY_UPDATE_SCENE
    r.render(True)
    a, na = r.readRGBA32F(), r.lastRaySamples()
    p = s.lic_params()
    s2 = configs.cfg3(n=32, size=64)
    s2.params.update(stepSizeVol=p.stepSizeVol / 2, stepsForward=p.stepsForward + 1, stepsBackward=p.stepsBackward - 1)
    s2.lowres = 1
    s2.clip_planes = ((0.0, 0.0, -1.0, 0.0),)      # ClipPlane ctor, VV/transform.cpp:240-254
    _, b, _, _, nb = render_cuda(vv, s2)
    assert na == nb and na > 0 and np.array_equal(a, b)
    assert r.keyboard(st, "q") == vv.KEY_QUIT and r.keyboard(st, "H") == 0


def test_idle_tick_animation(vv, tmp_path):
    """vv_idle = one idle() tick of the reference's animation (VV/3DLIC.cpp:129-142): the RGBA16F vector texture after k ticks
    equals, bit for bit, the texture VectorDataSet::createTextureIterp uploads at tick k (VV/dataset.cpp compiled unmodified,
    oracle/ref_app_driver.cpp) -- across the point where the pair of time steps moves on"""
    from oracle import refhost, vvo
    from vectorvisualization_b200 import fields as F
    if not refhost.has_app():
        pytest.skip("oracle/_ref built without VV/3DLIC.cpp")
    steps = [np.ascontiguousarray(F.abc_flow(10)[:, :8, :6] * np.float32(1 + 0.3 * t)) for t in range(3)]
    dat = F.write_dat(str(tmp_path / "anim.dat"), None, time_steps=steps)
    shape = steps[0].shape[:3]
    r = vv.Renderer(0)
    r.loadDat(dat)
    seq, _ = refhost.animation_ticks(dat, 24)
    for k in range(24):
        c = r.timeCursor()
        assert (c.current, c.interp_index) == seq[k][:2]
        r.idle()
        if k in (0, 5, 8, 9, 10, 19, 23):
            _, tex = refhost.animation_ticks(dat, k + 1, tex_tick=k, tex_shape=shape)
            got = r.readFieldTexture(shape)
            assert np.array_equal(got.view(np.uint32), vvo.half_round(tex).view(np.uint32)), k


def test_float_target_outputs(vv, oracle, tmp_path):
    """enableFBO (key 'F'): the stored frame is an RGBA16F texture (VV/renderer.cpp:562-606) -- read back rounded to fp16 -- and
    saveTexture writes a float texture as (int)(255 * texel), truncated (VV/renderer.cpp:386-403)"""
    import os
    from vectorvisualization_b200 import configs
    s = configs.cfg1(n=16, size=40)
    r = vv.Renderer(0)
    configs.apply_scene(r, s)
    out = str(tmp_path / "snapshotOut")
    os.makedirs(out)
    r.setSnapshot(out, "f.png")
    r.render(True)
    a = r.readRGBA32F()
    a8 = r.readRGBA8()
    r.enableFBO(True)
    b = r.readRGBA32F()
    want = oracle.half_round(a)
    assert np.array_equal(b.view(np.uint32), want.view(np.uint32)) and not np.array_equal(a, b)
    assert np.array_equal(r.readRGBA8(), a8)                       # the RGBA8 back-buffer frame is what it was
    r.screenshot()
    r.render(False)
    png = vv.png_read(r.lastSnapshotPath())[::-1]
    trunc = np.clip((np.float32(255.0) * want).astype(np.int32), 0, 255).astype(np.uint8)
    assert np.array_equal(png, trunc)
    assert int(np.abs(png.astype(np.int32) - a8.astype(np.int32)).max()) <= 1 and not np.array_equal(png, a8)
    r.enableFBO(False)
    assert np.array_equal(r.readRGBA32F(), a)
