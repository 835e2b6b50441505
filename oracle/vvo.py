"""ctypes wrapper of oracle/libvv_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product package (vectorvisualization_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libvv_oracle.so")


def build(force=False):
    src = [os.path.join(HERE, "vv_oracle.cpp"), os.path.join(HERE, "vv_oracle.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in src):
        return LIB
    subprocess.check_call(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                           "-o", LIB, src[0]])
    return LIB


class Scene(ctypes.Structure):
    _fields_ = [
        ("vec", ctypes.c_void_p), ("vdim", ctypes.c_int * 3),
        ("scalar", ctypes.c_void_p), ("sdim", ctypes.c_int * 3),
        ("noise", ctypes.c_void_p), ("ndim", ctypes.c_int * 3), ("noise_channels", ctypes.c_int),
        ("kernel", ctypes.c_void_p), ("kwidth", ctypes.c_int), ("inv_filter_area", ctypes.c_float),
        ("tf", ctypes.c_void_p),
        ("licvol", ctypes.c_void_p), ("ldim", ctypes.c_int * 3),
        ("illum_tex", ctypes.c_void_p * 3), ("illum_dim", ctypes.c_int * 2),
        ("extent", ctypes.c_float * 3), ("scale", ctypes.c_float * 3), ("scale_inv", ctypes.c_float * 3), ("center", ctypes.c_float * 3),
        ("step_size_vol", ctypes.c_float), ("gradient_scale", ctypes.c_float), ("illum_scale", ctypes.c_float), ("freq_scale", ctypes.c_float),
        ("num_iterations", ctypes.c_int), ("steps_fwd", ctypes.c_int), ("steps_bwd", ctypes.c_int), ("step_size_lic", ctypes.c_float),
        ("cam_quat", ctypes.c_float * 4), ("cam_pos", ctypes.c_float * 3), ("cam_dist", ctypes.c_float), ("fovy", ctypes.c_float),
        ("light_quat", ctypes.c_float * 4), ("light_dist", ctypes.c_float), ("spec_exp", ctypes.c_float),
        ("width", ctypes.c_int), ("height", ctypes.c_int),
        ("illum_mode", ctypes.c_int), ("tf_mode", ctypes.c_int), ("gate_mode", ctypes.c_int), ("noise_gate", ctypes.c_int),
        ("lowres", ctypes.c_int), ("quirk_scalevolinv", ctypes.c_int), ("quirk_luminance_alpha", ctypes.c_int),
        ("speed_of_flow", ctypes.c_int), ("licvol_fp16", ctypes.c_int), ("weight_bits", ctypes.c_int),
        ("mc_offsets", ctypes.c_void_p), ("num_clip_planes", ctypes.c_int), ("clip_planes", (ctypes.c_double * 4) * 3),
        ("near_clip", ctypes.c_float), ("far_clip", ctypes.c_float), ("window_aspect", ctypes.c_float), ("fbo_fp16", ctypes.c_int),
        ("fbo_pingpong", ctypes.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        P, I, F, U64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint64
        S = ctypes.POINTER(Scene)
        L.vvo_raycast_lic.argtypes = [S, P, P]; L.vvo_raycast_lic.restype = U64
        L.vvo_raycast_lic_rect.argtypes = [S, I, I, I, I, P, P]; L.vvo_raycast_lic_rect.restype = U64
        L.vvo_lic_volume.argtypes = [S, I, I, I, I, I, P]; L.vvo_lic_volume.restype = None
        L.vvo_raycast_licvolume.argtypes = [S, P, P]; L.vvo_raycast_licvolume.restype = U64
        L.vvo_compute_lic.argtypes = [S, P, P]; L.vvo_compute_lic.restype = None
        L.vvo_debug_walk.argtypes = [S, P, ctypes.c_int, ctypes.c_int, P]; L.vvo_debug_walk.restype = None
        L.vvo_slicing_lic.argtypes = [S, P, P]; L.vvo_slicing_lic.restype = U64
        L.vvo_slicing_setup.argtypes = [S, P]
        L.vvo_slice_fragments.argtypes = [S, I, I, P, I]; L.vvo_slice_fragments.restype = I
        L.vvo_slice_fragment_colors.argtypes = [S, I, I, P, I]; L.vvo_slice_fragment_colors.restype = I
        L.vvo_slicing_blend8.argtypes = [S, P, P]; L.vvo_slicing_blend8.restype = U64
        L.vvo_background.argtypes = [P, I, P]; L.vvo_quantize_rgba8.argtypes = [P, I, P]
        L.vvo_display_window.argtypes = [P, I, I, I, I, P]
        for n in ("vvo_sample_vec", "vvo_sample_noise", "vvo_sample_scalar"):
            getattr(L, n).argtypes = [S, P, P]
        L.vvo_sample_kernel.argtypes = [S, F]; L.vvo_sample_kernel.restype = F
        L.vvo_sample_tf.argtypes = [S, F, P, P]
        L.vvo_derive_uniforms.argtypes = [S, P]
        L.vvo_view.argtypes = [S, P, P]
        L.vvo_light_position.argtypes = [S, P]
        L.vvo_pixel_ray.argtypes = [S, I, I, P, P]; L.vvo_pixel_ray.restype = I
        L.vvo_volume_geometry.argtypes = [P, P, P, P, P, P]
        L.vvo_pixel_rays.argtypes = [S, I, I, I, I, P]
        L.vvo_scale_uniforms.argtypes = [S, P]
        L.vvo_pack_vector_field.argtypes = [P, P, P, I, I, I, P]
        L.vvo_pack_vector_field_u8.argtypes = [P, P, I, P]
        L.vvo_noise_gradients.argtypes = [P, P, P, P]
        L.vvo_compute_gradients_f.argtypes = [P, P, P, P]
        L.vvo_filter_gradients_f.argtypes = [P, P]
        L.vvo_pack_noise_rgba.argtypes = [P, P, I, P]
        L.vvo_white_noise.argtypes = [I, ctypes.c_uint32, F, P]
        L.vvo_filter_from_row.argtypes = [P, I, I, P, P]; L.vvo_filter_from_row.restype = I
        L.vvo_box_filter.argtypes = [I, P, P]; L.vvo_box_filter.restype = I
        L.vvo_default_tf.argtypes = [P]
        L.vvo_half_round.argtypes = [F]; L.vvo_half_round.restype = F
        L.vvo_illum_tables.argtypes = [F, I, I, P, P, P]
        L.vvo_num_threads.restype = I
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _i3(x):
    return (ctypes.c_int * 3)(*x)


def pack_vector_field(v0, v1=None, interp=(0, 10), fp16=True):
    v0 = np.ascontiguousarray(v0)
    nz, ny, nx = v0.shape[:3]
    out = np.empty((nz, ny, nx, 4), dtype=np.float32)
    if v0.dtype == np.uint8:
        lib().vvo_pack_vector_field_u8(_p(v0), _i3((nx, ny, nz)), int(fp16), _p(out))
    else:
        v0 = np.ascontiguousarray(v0, dtype=np.float32)
        v1c = np.ascontiguousarray(v1, dtype=np.float32) if v1 is not None else None
        lib().vvo_pack_vector_field(_p(v0), _p(v1c), _i3((nx, ny, nz)), interp[0], interp[1], int(fp16), _p(out))
    return out


def noise_gradients(noise, slice_dist=(1.0, 1.0, 1.0)):
    noise = np.ascontiguousarray(noise, dtype=np.uint8)
    nz, ny, nx = noise.shape
    out = np.empty((nz, ny, nx, 3), dtype=np.uint8)
    lib().vvo_noise_gradients(_p(noise), _i3((nx, ny, nz)), (ctypes.c_float * 3)(*slice_dist), _p(out))
    return out


def pack_noise_rgba(noise, grad):
    out = np.empty(noise.shape + (4,), dtype=np.uint8)
    lib().vvo_pack_noise_rgba(_p(np.ascontiguousarray(noise)), _p(np.ascontiguousarray(grad)), noise.size, _p(out))
    return out


def white_noise(n, seed, p):
    out = np.empty((n, n, n), dtype=np.uint8)
    lib().vvo_white_noise(n ** 3, seed, p, _p(out))
    return out


def filter_from_row(row, channels=1):
    row = np.ascontiguousarray(row, dtype=np.uint8)
    width = row.size // channels
    out = np.zeros(1 << max(0, (width - 1)).bit_length(), dtype=np.uint8)
    inv = ctypes.c_float()
    fw = lib().vvo_filter_from_row(_p(row), width, channels, _p(out), ctypes.byref(inv))
    return out[:fw].copy(), inv.value


def box_filter(width=256):
    out = np.zeros(1 << max(0, (width - 1)).bit_length(), dtype=np.uint8)
    inv = ctypes.c_float()
    fw = lib().vvo_box_filter(width, _p(out), ctypes.byref(inv))
    return out[:fw].copy(), inv.value


def illum_tables(spec_exp=40.0, w=256, h=256):
    z = np.zeros((h, w, 2), np.float32); d = np.zeros((h, w), np.float32); s = np.zeros((h, w), np.float32)
    lib().vvo_illum_tables(spec_exp, w, h, _p(z), _p(d), _p(s))
    return z, d, s


def half_round(a):
    """float32 -> fp16 -> float32 (round to nearest even), the GL_*16F upload rounding"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32))


def default_tf():
    out = np.empty((256, 5), dtype=np.uint8)
    lib().vvo_default_tf(_p(out))
    return out


ILLUM = {"": 0, "ILLUM_GRADIENT": 1, "ILLUM_MALLO": 2, "ILLUM_ZOECKLER": 3}


class OracleScene:
    """Builds the VVOScene for a vectorvisualization_b200.configs.Scene using the ORACLE's own pre-processing
    (pack / gradients / filter), and keeps the numpy arrays alive."""

    def __init__(self, s, weight_bits=0, illum_tables=None, fbo_fp16=1, fbo_pingpong=1):
        """fbo_fp16 / fbo_pingpong: the FBO slicing path as Renderer::sliceVolume runs it (RGBA16F targets, two textures swapped per
        slice); 0 / 0 = one fp32 accumulator per pixel (the idealised model)"""
        L = lib()
        self.s = s
        self.keep = []
        c = Scene()
        nz, ny, nx = s.field.shape[:3]
        self.vec = pack_vector_field(s.field, s.next_field, s.interp, fp16=True)
        c.vec = _p(self.vec); c.vdim = _i3((nx, ny, nz))
        if s.scalar is not None:
            self.scalar = np.ascontiguousarray(s.scalar, dtype=np.uint8)
            c.scalar = _p(self.scalar); c.sdim = _i3(self.scalar.shape[::-1])
        noise = np.ascontiguousarray(s.noise, dtype=np.uint8)
        if s.with_gradients:
            self.grad = noise_gradients(noise)
            self.noise = pack_noise_rgba(noise, self.grad)
            c.noise_channels = 4
        else:
            self.noise = noise
            c.noise_channels = 1
        c.noise = _p(self.noise); c.ndim = _i3(noise.shape[::-1])
        illum_mode = 0
        for k, v in ILLUM.items():
            if k and k in (s.defines or ""):
                illum_mode = v
        if illum_mode != 1 and s.with_gradients:
            # scalar builds read the noise through .a of the RGBA texture: same values as the LUMINANCE path
            # with quirk_luminance_alpha off
            self.noise = noise
            c.noise = _p(self.noise); c.noise_channels = 1
        if s.filter_row is None:
            self.kernel, inv = box_filter(256)
        else:
            self.kernel, inv = filter_from_row(s.filter_row)
        c.kernel = _p(self.kernel); c.kwidth = self.kernel.size; c.inv_filter_area = inv
        self.tf = np.ascontiguousarray(s.tf, dtype=np.uint8)
        c.tf = _p(self.tf)
        size = _i3((nx, ny, nz)); sd = (ctypes.c_float * 3)(*s.slice_dist)
        L.vvo_volume_geometry(size, sd, c.extent, c.scale, c.scale_inv, c.center)
        p = s.lic_params()
        c.step_size_vol = p.stepSizeVol; c.gradient_scale = p.gradientScale; c.illum_scale = p.illumScale; c.freq_scale = p.freqScale
        c.num_iterations = p.numIterations; c.steps_fwd = p.stepsForward; c.steps_bwd = p.stepsBackward; c.step_size_lic = p.stepSizeLIC
        c.cam_quat = (ctypes.c_float * 4)(*s.camera["quat"]); c.cam_pos = (ctypes.c_float * 3)(*s.camera["pos"])
        c.cam_dist = s.camera["dist"]; c.fovy = s.camera["fovy"]
        c.near_clip = s.camera.get("near", 0.1); c.far_clip = s.camera.get("far", 50.0)          # VV/camera.cpp:42-47
        win = getattr(s, "window", None)
        c.window_aspect = float(np.float32(win[0]) / np.float32(win[1])) if win else 0.0           # _aspect = (float)_w/_h
        c.light_quat = (ctypes.c_float * 4)(*s.light["quat"]); c.light_dist = s.light["dist"]; c.spec_exp = 40.0
        c.width = s.width; c.height = s.height
        c.illum_mode = illum_mode; c.tf_mode = s.tf_mode; c.gate_mode = s.gate_mode; c.noise_gate = s.noise_gate
        c.lowres = s.lowres; c.quirk_scalevolinv = s.quirk_scalevolinv
        c.quirk_luminance_alpha = 0 if s.with_gradients else s.quirk_luminance_alpha   # Q7 only bites GL_LUMINANCE noise
        c.speed_of_flow = 1 if "SPEED_OF_FLOW" in (s.defines or "") else 0
        c.licvol_fp16 = s.licvol_fp16; c.weight_bits = weight_bits; c.fbo_fp16 = fbo_fp16; c.fbo_pingpong = fbo_pingpong
        if getattr(s, "mc_offsets", None) is not None:
            assert "USE_MC_OFFSET" in (s.defines or ""), "mc_offsets need #define USE_MC_OFFSET"
            self.mc = half_round(np.ascontiguousarray(s.mc_offsets, dtype=np.float32).reshape(s.height, s.width))
            c.mc_offsets = self.mc.ctypes.data
        planes = list(getattr(s, "clip_planes", ()) or ())
        c.num_clip_planes = len(planes)
        for i, e in enumerate(planes):
            for k in range(4):
                c.clip_planes[i][k] = float(e[k])
        if illum_tables is not None:
            self.illum_tables = [np.ascontiguousarray(t, dtype=np.float32) for t in illum_tables]
            for i, t in enumerate(self.illum_tables):
                c.illum_tex[i] = t.ctypes.data
            c.illum_dim = (ctypes.c_int * 2)(self.illum_tables[0].shape[1], self.illum_tables[0].shape[0])
        self.c = c

    def slicing_blend8(self):
        """slicing without the FBO (the start-up state of the reference): RGBA8 back buffer [h][w][4], samples, total"""
        s = self.s
        out = np.zeros((s.height, s.width, 4), dtype=np.uint8)
        cnt = np.zeros((s.height, s.width), dtype=np.uint32)
        tot = lib().vvo_slicing_blend8(ctypes.byref(self.c), _p(out), _p(cnt))
        return out, cnt, int(tot)

    def slice_fragment_colors(self, x, y):
        _, _, n = self.slicing_setup()
        buf = np.zeros((n, 4), np.float32)
        k = lib().vvo_slice_fragment_colors(ctypes.byref(self.c), x, y, _p(buf), n)
        assert k >= 0
        return buf[:k].copy()

    def raycast(self, rect=None):
        """returns (rgba float [h][w][4], samples uint32 [h][w], total)"""
        s = self.s
        out = np.zeros((s.height, s.width, 4), dtype=np.float32)
        cnt = np.zeros((s.height, s.width), dtype=np.uint32)
        if rect is None:
            tot = lib().vvo_raycast_lic(ctypes.byref(self.c), _p(out), _p(cnt))
        else:
            tot = lib().vvo_raycast_lic_rect(ctypes.byref(self.c), rect[0], rect[1], rect[2], rect[3], _p(out), _p(cnt))
        return out, cnt, int(tot)

    def slicing(self):
        """returns (rgba float [h][w][4], shaded fragments per pixel, total)"""
        s = self.s
        out = np.zeros((s.height, s.width, 4), dtype=np.float32)
        cnt = np.zeros((s.height, s.width), dtype=np.uint32)
        tot = lib().vvo_slicing_lic(ctypes.byref(self.c), _p(out), _p(cnt))
        return out, cnt, int(tot)

    def slicing_setup(self):
        o = np.zeros(5, np.float32)
        lib().vvo_slicing_setup(ctypes.byref(self.c), _p(o))
        return o[:3].copy(), float(o[3]), int(o[4])

    def lic_volume(self, dims=None, z0=0, z1=None):
        nz, ny, nx = self.s.field.shape[:3]
        w, h, d = dims or (nx, ny, nz)
        out = np.zeros((d, h, w), dtype=np.float32)
        lib().vvo_lic_volume(ctypes.byref(self.c), w, h, d, z0, d if z1 is None else z1, _p(out))
        return out

    def raycast_licvolume(self, licvol):
        s = self.s
        self.licvol = np.ascontiguousarray(licvol, dtype=np.float32)
        self.c.licvol = _p(self.licvol); self.c.ldim = _i3(self.licvol.shape[::-1])
        out = np.zeros((s.height, s.width, 4), dtype=np.float32)
        cnt = np.zeros((s.height, s.width), dtype=np.uint32)
        tot = lib().vvo_raycast_licvolume(ctypes.byref(self.c), _p(out), _p(cnt))
        return out, cnt, int(tot)

    def compute_lic(self, pos):
        out = np.zeros(4, dtype=np.float32)
        lib().vvo_compute_lic(ctypes.byref(self.c), _p(np.asarray(pos, dtype=np.float32)), _p(out))
        return out

    def debug_walk(self, pos, dir_sign, nsteps):
        """one direction of computeLIC's walk: rows (newPos.xyz, step.rgb, noise tap, kernel weight, Pos2.xyz, step2.rgb, 0, 0)"""
        out = np.zeros((nsteps, 16), dtype=np.float32)
        lib().vvo_debug_walk(ctypes.byref(self.c), _p(np.asarray(pos, dtype=np.float32)), int(dir_sign), int(nsteps), _p(out))
        return out

    def uniforms(self):
        out = np.zeros(16, dtype=np.float32)
        lib().vvo_derive_uniforms(ctypes.byref(self.c), _p(out))
        return out


def quantize_rgba8(rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    out = np.empty(rgba.shape, dtype=np.uint8)
    lib().vvo_quantize_rgba8(_p(rgba), rgba.size, _p(out))
    return out


def display_window(rgba, win_w, win_h):
    """the display pass over a win_w x win_h window showing the stored frame rgba [rh][rw][4] (low-res preset: NEAREST up-scaling)"""
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    out = np.empty((win_h, win_w, 4), dtype=np.float32)
    lib().vvo_display_window(_p(rgba), rgba.shape[1], rgba.shape[0], win_w, win_h, _p(out))
    return out


def background(rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    out = np.empty_like(rgba)
    lib().vvo_background(_p(rgba), rgba.size // 4, _p(out))
    return out
