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
