"""Small scenes shared by the golden-vector generator and the tests that consume the vectors."""
import numpy as np


def golden_scenes():
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    S = {}

    def g1():
        return configs.cfg1(n=24, size=40)
    S["cfg1_default"] = g1

    def g2():
        s = configs.cfg2(n=24, size=40, camera=F.CAMERA_CLOSE)
        s.params.update(gradientScale=6.0)
        return s
    S["cfg2_close_gs6"] = g2

    def g3():
        s = configs.cfg3(n=24, size=40, camera=F.CAMERA_CLOSE)
        return s
    S["cfg3_gradient_length"] = g3

    def g4():
        s = configs.cfg3(n=20, size=36)
        s.tf_mode = vv.TF_B
        s.params.update(gradientScale=4.0, stepsForward=25, stepsBackward=40)
        s.light = dict(quat=F.quat_from_axis_angle((1, 0.3, 0), 50.0), dist=1.0)
        return s
    S["cfg3_tfb_steps_25_40"] = g4

    def g5():
        s = configs.cfg4(n=24, size=36, noise_n=16)
        return s
    S["cfg4_triangle_step256"] = g5

    def g6():
        s = configs.cfg2(n=20, size=36)
        s.gate_mode = vv.GATE_TF_ALPHA
        s.tf_mode = vv.TF_A
        s.params.update(gradientScale=5.0)
        return s
    S["gate_tf_alpha"] = g6

    def g7():
        s = configs.cfg1(n=20, size=36, camera=F.CAMERA_CLOSE)
        s.defines = "#define SPEED_OF_FLOW"
        s.tf_mode = vv.TF_R
        return s
    S["speed_of_flow_tf_r"] = g7

    def g8():
        s = configs.cfg1(n=16, size=32)
        s.field = np.ascontiguousarray(F.abc_flow(32)[::2, :24, :][:12])   # 32 x 24 x 12
        s.slice_dist = (1.0, 1.0, 2.0)
        s.tf_mode = vv.TF_SCALAR
        s.scalar = np.random.RandomState(5).randint(30, 90, size=(8, 8, 8)).astype(np.uint8)
        s.camera = dict(F.CAMERA_CLOSE)
        return s
    S["anisotropic_tf_scalar_band"] = g8
    return S


def golden_extra_scenes():
    """second family: the slicing program, the illuminated-streamline builds and the USE_MC_OFFSET builds.
    name -> (maker, kind) with kind "raycast" or "slicing"; scenes with ILLUM_MALLO / ILLUM_ZOECKLER need the
    illumination tables (oracle.illum_tables(40.0))"""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    S = {}

    def slicing(s):
        s.technique = vv.VOLIC_SLICING
        s.fbo = 1
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA          # what lic3d_slicing_fragment.glsl hard-codes
        s.tf = F.default_tf()
        s.params.update(gradientScale=4.0)
        return s

    def e1():
        s = configs.cfg3(n=20, size=32, camera=F.CAMERA_CLOSE)
        return slicing(s)
    S["x_slicing_gradient"] = (e1, "slicing")

    def e2():
        s = configs.cfg2(n=20, size=36)
        s.with_gradients = True
        return slicing(s)
    S["x_slicing_plain"] = (e2, "slicing")

    def e3():
        s = configs.cfg2(n=24, size=40, camera=F.CAMERA_CLOSE)
        s.defines = "#define ILLUM_MALLO"
        s.params.update(gradientScale=4.0, illumScale=1.3)
        s.light = dict(quat=F.quat_from_axis_angle((0.2, 1, 0), 70.0), dist=1.0)
        return s
    S["x_mallo_raycast"] = (e3, "raycast")

    def e4():
        s = configs.cfg1(n=20, size=36)
        s.defines = "#define ILLUM_ZOECKLER"
        s.params.update(gradientScale=4.0)
        s.light = dict(quat=F.quat_from_axis_angle((1, 0.2, 0), 40.0), dist=1.0)
        return s
    S["x_zoeckler_raycast"] = (e4, "raycast")

    def e5():
        s = configs.cfg3(n=24, size=40)
        s.tf_mode = vv.TF_B
        s.defines += "\n#define USE_MC_OFFSET"
        s.mc_offsets = np.random.RandomState(21).rand(s.height, s.width).astype(np.float32)
        return s
    S["x_mc_raycast_gradient"] = (e5, "raycast")

    def e6():
        s = configs.cfg1(n=24, size=40)
        s = slicing(s)
        s.defines = "#define USE_MC_OFFSET"
        s.mc_offsets = np.random.RandomState(22).rand(s.height, s.width).astype(np.float32)
        return s
    S["x_mc_slicing"] = (e6, "slicing")
    return S


def golden_chain_scenes():
    """third family: frames of the WHOLE reference chain -- the draw calls of Renderer::render(true) (VV/renderer.cpp compiled
    unmodified, GL calls captured), rasterised per the OpenGL 2.1 specification (oracle/softgl.py), shaded by the reference's
    shader code.  name -> (maker, kind); the fragments' texcoords are the rasteriser's, not the oracle's analytic entry points,
    so consumers compare within the north-star tolerance instead of bit for bit."""
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs, fields as F
    S = {}

    def c1():
        s = configs.cfg3(n=20, size=40, camera=dict(quat=F.quat_from_axis_angle((0.3, -1.0, 0.2), 110.0), pos=(0.1, 0.0, 0.2), dist=3.0, fovy=35.0))
        s.clip_planes = ((0.6, 0.0, -0.8, 0.05), (1.0, 0.0, 0.0, 0.12))
        return s
    S["y_chain_raycast_gradient_clip2"] = (c1, "raycast")

    def c2():
        # the near plane (0.1) cuts the front face: a hole where the entry point is nearer than it
        s = configs.cfg1(n=20, size=48, camera=dict(quat=F.quat_from_axis_angle((0.2, 1.0, 0.1), 50.0), pos=(0.0, 0.0, 0.0), dist=0.75, fovy=35.0))
        return s
    S["y_chain_raycast_near_plane"] = (c2, "raycast")

    def c3():
        s = configs.cfg2(n=20, size=40, camera=F.CAMERA_CLOSE)
        s.height = 30
        s.field = np.ascontiguousarray(F.rankine_vortex(24)[::2, :16, :])     # 24 x 16 x 12, anisotropic spacing
        s.slice_dist = (1.0, 1.5, 2.0)
        s.clip_planes = ((0.3, 0.5, -0.8, 0.05),)                            # |n| = 0.99: the steady state (n / |n|, d)
        return s
    S["y_chain_raycast_aniso_nonunit_clip"] = (c3, "raycast")

    def c4():
        s = configs.cfg2(n=20, size=36, camera=F.CAMERA_CLOSE)
        s.with_gradients = True
        s.technique = vv.VOLIC_SLICING
        s.fbo = 1
        s.tf_mode, s.gate_mode = vv.TF_A, vv.GATE_TF_ALPHA
        s.tf = F.default_tf()
        s.params.update(gradientScale=4.0, stepSizeVol=1 / 64)
        s.clip_planes = ((0.0, 0.0, -1.0, 0.2),)
        return s
    S["y_chain_slicing_clip"] = (c4, "slicing")
    return S
