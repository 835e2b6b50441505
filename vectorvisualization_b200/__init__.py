"""vectorvisualization_b200 -- B200-native 3D-LIC renderer (drop-in for liaoyg/VectorVisualization's LIC path).

This module is only the Python mirror of the C ABI in include/vv_c_api.h (ctypes); all work happens in
libvv_b200.so (hand-written CUDA for sm_100a).  There is no CPU fallback: if the library is missing or no CUDA
device is present, construction raises.

`Renderer` keeps the method names of the reference's `class Renderer` (VV/renderer.h:28-125) and of the objects
it is handed (VectorDataSet / NoiseDataSet / LICFilter / TransferEdit / Camera / LICParams).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VV_B200_LIB") or os.path.join(_HERE, "libvv_b200.so")   # override only for A/B experiments

# RenderTechnique, VV/types.h:63-71
VOLIC_VOLUME, VOLIC_RAYCAST, VOLIC_SLICING, VOLIC_LICVOLUME, VOLIC_VOLUMEANI = range(5)
# VVDataType
UCHAR, USHORT, FLOAT = 1, 2, 3
# options
TF_B, TF_A, TF_R, TF_LENGTH, TF_SCALAR = range(5)
GATE_ALWAYS, GATE_TF_ALPHA = 0, 1
(OPT_TF_MODE, OPT_GATE_MODE, OPT_NOISE_GATE, OPT_QUIRK_SCALEVOLINV, OPT_QUIRK_LUMINANCE_ALPHA, OPT_LICVOL_FP16,
 OPT_FIELD_LAYOUT, OPT_COUNT_SAMPLES, OPT_LICVOL_SIZE, OPT_SPEC_EXP, OPT_SAMPLE_MAP, OPT_RAYCAST_MODE,
 OPT_LIC_CTAS_PER_SM, OPT_WALK_FAST_PATHS, OPT_DEPTH_MAJOR, OPT_BAND_ROWS, OPT_NOISE_LAYOUT, OPT_FIRST_WINDOW, OPT_WINDOW_GROWTH, OPT_PARTITION_UNIT) = range(1, 21)
LAYOUT_F4, LAYOUT_PAIR, LAYOUT_QUAD, LAYOUT_AUTO = 0, 1, 2, 3   # AUTO (default): QUAD up to 48 GiB of packed field, else PAIR
BLOCK = 16  # pixels per image-block edge (sort-first partition unit)


class LICParams(ctypes.Structure):
    """struct LICParams, VV/types.h:91-109 (same defaults)"""
    _fields_ = [("stepSizeVol", ctypes.c_float), ("gradientScale", ctypes.c_float), ("illumScale", ctypes.c_float),
                ("freqScale", ctypes.c_float), ("numIterations", ctypes.c_int), ("stepsForward", ctypes.c_int),
                ("stepsBackward", ctypes.c_int), ("stepSizeLIC", ctypes.c_float)]

    def __init__(self, **kw):
        super().__init__(1.0 / 128.0, 30.0, 1.0, 1.0, 255, 32, 32, 0.01)
        for k, v in kw.items():
            setattr(self, k, v)


class AppState(ctypes.Structure):
    """VVAppState: the application state the reference keeps in globals (VV/3DLIC.h:29-55), driven by keyApply / Renderer.keyboard"""
    _fields_ = [("lic", LICParams), ("technique", ctypes.c_int), ("lowres", ctypes.c_int), ("float_target", ctypes.c_int),
                ("continuous", ctypes.c_int), ("recording", ctypes.c_int), ("animation", ctypes.c_int), ("screenshot", ctypes.c_int),
                ("clip_active", ctypes.c_int * 3), ("selected_clip", ctypes.c_int), ("defines", ctypes.c_char * 64)]

    def __init__(self):
        super().__init__()
        load_library().vv_app_state_init(ctypes.byref(self))


KEY_UPDATE_SCENE, KEY_RELOAD_SHADER, KEY_UPDATE_LICVOLUME, KEY_UPDATE_SLICES, KEY_QUIT, KEY_SCREENSHOT, KEY_SWITCH_RECORDING, \
    KEY_SET_TECHNIQUE = 1, 2, 4, 8, 16, 32, 64, 128


def keyApply(state, key, special=False):
    """vv_key_apply: one key of VV/3DLIC.cpp's keyboard / keyboardSpecial applied to an AppState (host logic only)"""
    k = key if isinstance(key, int) else ord(key)
    return load_library().vv_key_apply(ctypes.byref(state), k, int(bool(special)))


class TimeCursor(ctypes.Structure):
    """VVTimeCursor: interpIndex / InterpSize and the current time step of an animated VectorDataSet (VV/dataset.cpp:202-210)"""
    _fields_ = [("time_begin", ctypes.c_int), ("time_end", ctypes.c_int), ("current", ctypes.c_int), ("interp_index", ctypes.c_int),
                ("interp_size", ctypes.c_int)]

    def __init__(self, time_begin=0, time_end=0, interp_size=10):
        super().__init__()
        load_library().vv_time_cursor_init(ctypes.byref(self), time_begin, time_end, interp_size)

    def next(self):
        return load_library().vv_time_cursor_next(ctypes.byref(self))

    def tick(self):
        """one createTextureIterp + checkInterpolateStage: (fraction index of this tick's texture, moved on to the next pair)"""
        adv = ctypes.c_int(0)
        used = load_library().vv_time_cursor_tick(ctypes.byref(self), ctypes.byref(adv))
        return used, bool(adv.value)


class DatInfo(ctypes.Structure):
    _fields_ = [("raw_file", ctypes.c_char * 512), ("resolution", ctypes.c_int * 3), ("slice_thickness", ctypes.c_float * 3),
                ("data_type", ctypes.c_int), ("data_dim", ctypes.c_int), ("time_begin", ctypes.c_int), ("time_end", ctypes.c_int)]


class Args(ctypes.Structure):
    _fields_ = [("vol_file", ctypes.c_char * 512), ("noise_file", ctypes.c_char * 512), ("tf_file", ctypes.c_char * 512),
                ("filter_file", ctypes.c_char * 512), ("redirect_file", ctypes.c_char * 512), ("halton_file", ctypes.c_char * 512),
                ("use_gradients", ctypes.c_int), ("use_lambda2", ctypes.c_int), ("show_help", ctypes.c_int)]


class VVError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vv error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """dlopen libvv_b200.so; raises if it has not been built (python vectorvisualization_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libvv_b200.so is missing (%s): build it with `python vectorvisualization_b200/build.py`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P, I, F, U64, CP = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint64, ctypes.c_char_p
    sig = {
        "vv_create": ([ctypes.POINTER(P), I], I), "vv_destroy": ([P], None), "vv_init": ([P, CP], I),
        "vv_load_glsl_shader": ([P, CP], I), "vv_resize": ([P, I, I], I), "vv_last_error": ([], CP), "vv_version": ([], CP),
        "vv_set_technique": ([P, I], I), "vv_update_lic_volume": ([P], I), "vv_update_slices": ([P], I),
        "vv_set_vector_field": ([P, P, P, I, ctypes.POINTER(I), ctypes.POINTER(F)], I),
        "vv_set_time_interp": ([P, I, I], I), "vv_set_scalar": ([P, P, I, ctypes.POINTER(I)], I),
        "vv_set_noise": ([P, P, ctypes.POINTER(I), I], I), "vv_generate_white_noise": ([P, I, ctypes.c_uint32, F, I], I),
        "vv_set_filter": ([P, P, I, I], I), "vv_set_box_filter": ([P, I], I), "vv_set_tf": ([P, P], I),
        "vv_set_default_tf": ([P], I), "vv_set_lic_params": ([P, ctypes.POINTER(LICParams)], I),
        "vv_default_lic_params": ([ctypes.POINTER(LICParams)], None),
        "vv_set_camera": ([P, ctypes.POINTER(F), ctypes.POINTER(F), F, F, F, F], I),
        "vv_set_light": ([P, ctypes.POINTER(F), F], I), "vv_update_light_pos": ([P], I),
        "vv_enable_lowres": ([P, I], I), "vv_set_window": ([P, I, I], I),
        "vv_time_cursor_init": ([P, I, I, I], None), "vv_time_cursor_next": ([P], I), "vv_time_cursor_tick": ([P, P], I),
        "vv_idle": ([P], I), "vv_get_time_cursor": ([P, P], I),
        "vv_app_state_init": ([P], None), "vv_key_apply": ([P, I, I], I), "vv_keyboard": ([P, P, I, I], I),
        "vv_enable_float_target": ([P, I], I), "vv_set_option": ([P, I, I], I),
        "vv_set_mc_offsets": ([P, P, I, I], I), "vv_update_mc_offset_tex": ([P, I, I, ctypes.c_uint32], I),
        "vv_set_clip_plane": ([P, I, P, I], I),
        "vv_set_snapshot": ([P, CP, CP, I], I), "vv_screenshot": ([P], I), "vv_switch_recording": ([P], I),
        "vv_last_snapshot_path": ([P], CP),
        "vv_render": ([P, I], I), "vv_read_rgba8": ([P, P, ctypes.c_size_t], I), "vv_read_rgba32f": ([P, P, ctypes.c_size_t], I),
        "vv_read_display_rgba8": ([P, P, ctypes.c_size_t], I),
        "vv_read_display_window_rgba8": ([P, P, ctypes.c_size_t, I, I], I),
        "vv_read_lic_volume": ([P, P, ctypes.c_size_t, ctypes.POINTER(I)], I),
        "vv_read_field_texture": ([P, P, ctypes.c_size_t], I),
        "vv_read_noise_texture": ([P, P, ctypes.c_size_t, ctypes.POINTER(I)], I),
        "vv_read_sample_map": ([P, P, ctypes.c_size_t], I),
        "vv_make_illum_tables": ([F, I, I, P, P, P], I),
        "vv_save_png": ([P, CP, I], I), "vv_save_raw": ([P, CP], I),
        "vv_last_ray_samples": ([P], U64), "vv_last_kernel_ms": ([P], F), "vv_last_launch_count": ([P], I), "vv_field_layout": ([P], I),
        "vv_synchronize": ([P], I), "vv_set_partition": ([P, I, I], I), "vv_set_licvol_slab": ([P, I, I], I),
        "vv_get_tile_buffer": ([P, ctypes.POINTER(P), ctypes.POINTER(I), ctypes.POINTER(I)], I),
        "vv_assemble_tiles": ([P, P, I], I), "vv_get_lic_volume_ptr": ([P, ctypes.POINTER(P), ctypes.POINTER(I)], I),
        "vv_set_stream": ([P, P], I),
        "vv_debug_walk": ([P, P, I, I, I, P, ctypes.c_size_t], I),
        "vv_p2p_export": ([P, P, ctypes.POINTER(P)], I), "vv_p2p_connect": ([P, P, ctypes.POINTER(P), I], I),
        "vv_p2p_render": ([P], I), "vv_p2p_status": ([P], I), "vv_p2p_disconnect": ([P], I),
        "vv_parse_dat": ([CP, ctypes.POINTER(DatInfo)], I), "vv_read_raw": ([ctypes.POINTER(DatInfo), I, P, ctypes.c_size_t], I),
        "vv_load_dat": ([P, CP], I), "vv_load_scalar_dat": ([P, CP], I), "vv_load_noise": ([P, CP, I], I),
        "vv_load_filter_png": ([P, CP], I), "vv_load_tf_png": ([P, CP], I),
        "vv_grd_read": ([CP, P, P], I), "vv_grd_write": ([CP, P, P], I),
        "vv_set_noise_with_gradients": ([P, P, P, P], I), "vv_read_noise_gradients": ([P, P, ctypes.c_size_t], I),
        "vv_parse_args": ([I, ctypes.POINTER(CP), ctypes.POINTER(Args)], I), "vv_usage": ([], CP),
        "vv_png_read": ([CP, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.POINTER(I), ctypes.POINTER(I), ctypes.POINTER(I)], I),
        "vv_free": ([P], None), "vv_png_write": ([CP, P, I, I, I], I),
    }
    for name, (args, res) in sig.items():
        if name == "vv_debug_walk" and os.environ.get("VV_B200_LIB") and not hasattr(lib, name):
            continue              # an older A/B build selected with VV_B200_LIB (scripts/ab.py): the diagnostic entry point may be missing
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


API_SYMBOLS = None  # filled lazily by tests from include/vv_c_api.h


def _chk(rc):
    if rc != 0:
        raise VVError(rc, load_library().vv_last_error().decode(errors="replace"))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def parse_args(argv):
    """ParseArguments::parse (VV/parseArg.cpp:97-365); argv includes the program name."""
    lib = load_library()
    arr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])
    out = Args()
    _chk(lib.vv_parse_args(len(argv), arr, ctypes.byref(out)))
    return out


def parse_dat(path):
    lib = load_library()
    info = DatInfo()
    _chk(lib.vv_parse_dat(path.encode(), ctypes.byref(info)))
    return info


def make_illum_tables(spec_exp=40.0, width=256, height=256):
    """Zoeckler (LA) and Mallo (diffuse, specular) tables as the library builds them (VV/illumination.cpp:96-390)"""
    lib = load_library()
    z = np.zeros((height, width, 2), np.float32)
    d = np.zeros((height, width), np.float32)
    s = np.zeros((height, width), np.float32)
    _chk(lib.vv_make_illum_tables(spec_exp, width, height, _ptr(z), _ptr(d), _ptr(s)))
    return z, d, s


def grd_read(file_name, dims):
    """loadGradients(DATRAW_UCHAR), VV/gradient.cpp:112-149: "<file_name>.grd" -> uint8 [nz][ny][nx][3]; dims = (nx, ny, nz)"""
    lib = load_library()
    out = np.empty((dims[2], dims[1], dims[0], 3), np.uint8)
    _chk(lib.vv_grd_read(file_name.encode(), (ctypes.c_int * 3)(*dims), _ptr(out)))
    return out


def grd_write(file_name, gradients):
    """saveGradients(DATRAW_UCHAR), VV/gradient.cpp:152-187; gradients uint8 [nz][ny][nx][3]"""
    lib = load_library()
    g = np.ascontiguousarray(gradients, dtype=np.uint8)
    nz, ny, nx = g.shape[:3]
    _chk(lib.vv_grd_write(file_name.encode(), (ctypes.c_int * 3)(nx, ny, nz), _ptr(g)))


def png_read(path):
    lib = load_library()
    data = ctypes.POINTER(ctypes.c_uint8)()
    w, h, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _chk(lib.vv_png_read(path.encode(), ctypes.byref(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)))
    try:
        return np.ctypeslib.as_array(data, shape=(h.value, w.value, c.value)).copy()
    finally:
        lib.vv_free(data)


class Renderer:
    """Headless counterpart of `class Renderer` (VV/renderer.h:28-125)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        _chk(self._lib.vv_create(ctypes.byref(self._h), device))
        self.width = self.height = 0
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- lifecycle (VV/renderer.h:31-37,115)
    def init(self, defines=None):
        _chk(self._lib.vv_init(self._h, defines.encode() if defines else None))

    def loadGLSLShader(self, defines=None):
        _chk(self._lib.vv_load_glsl_shader(self._h, defines.encode() if defines else None))

    def resize(self, w, h):
        _chk(self._lib.vv_resize(self._h, w, h))
        self.width, self.height = w, h

    def keyboard(self, state, key, special=False):
        """keyboard / keyboardSpecial of VV/3DLIC.cpp applied to `state` and carried out on this renderer; returns the action mask"""
        k = key if isinstance(key, int) else ord(key)
        act = self._lib.vv_keyboard(self._h, ctypes.byref(state), k, int(bool(special)))
        if act < 0:
            _chk(1)
        return act

    def idle(self):
        """one animation tick of idle() (VV/3DLIC.cpp:129-172) for a field loaded with loadDat"""
        _chk(self._lib.vv_idle(self._h))

    def timeCursor(self):
        c = TimeCursor()
        _chk(self._lib.vv_get_time_cursor(self._h, ctypes.byref(c)))
        return c

    def setWindow(self, w, h):
        """Camera::setWindow (VV/transform.h:79-80): the aspect ratio of the projection when it is not the frame's (low-res preset)"""
        _chk(self._lib.vv_set_window(self._h, w, h))

    def setTechnique(self, t):
        _chk(self._lib.vv_set_technique(self._h, t))

    def updateLICVolume(self):
        _chk(self._lib.vv_update_lic_volume(self._h))

    def updateSlices(self):
        _chk(self._lib.vv_update_slices(self._h))

    # -- inputs
    def setVectorField(self, field, next_field=None, slice_dist=(1.0, 1.0, 1.0)):
        """field: [z][y][x][3] float32 (FLOAT3) or uint8 (UCHAR3)"""
        field = np.ascontiguousarray(field)
        assert field.ndim == 4 and field.shape[3] == 3
        dtype = {np.dtype(np.float32): FLOAT, np.dtype(np.uint8): UCHAR}[field.dtype]
        dims = (ctypes.c_int * 3)(field.shape[2], field.shape[1], field.shape[0])
        sd = (ctypes.c_float * 3)(*slice_dist)
        nxt = None
        if next_field is not None:
            next_field = np.ascontiguousarray(next_field, dtype=field.dtype)
            nxt = _ptr(next_field)
        _chk(self._lib.vv_set_vector_field(self._h, _ptr(field), nxt, dtype, dims, sd))

    def setTimeInterp(self, index, size=10):
        _chk(self._lib.vv_set_time_interp(self._h, index, size))

    def setScalar(self, vol):
        vol = np.ascontiguousarray(vol)
        dtype = {np.dtype(np.float32): FLOAT, np.dtype(np.uint8): UCHAR}[vol.dtype]
        dims = (ctypes.c_int * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
        _chk(self._lib.vv_set_scalar(self._h, _ptr(vol), dtype, dims))

    def setNoise(self, noise, with_gradients=False):
        noise = np.ascontiguousarray(noise, dtype=np.uint8)
        dims = (ctypes.c_int * 3)(noise.shape[2], noise.shape[1], noise.shape[0])
        _chk(self._lib.vv_set_noise(self._h, _ptr(noise), dims, int(with_gradients)))

    def generateWhiteNoise(self, n=256, seed=0, p=1.0 / 6.0, with_gradients=False):
        _chk(self._lib.vv_generate_white_noise(self._h, n, seed, p, int(with_gradients)))

    def setLICFilter(self, row=None, channels=1):
        if row is None:
            _chk(self._lib.vv_set_box_filter(self._h, 256))
        else:
            row = np.ascontiguousarray(row, dtype=np.uint8)
            _chk(self._lib.vv_set_filter(self._h, _ptr(row), row.size // channels, channels))

    def setTF(self, tf=None):
        if tf is None:
            _chk(self._lib.vv_set_default_tf(self._h))
        else:
            tf = np.ascontiguousarray(tf, dtype=np.uint8)
            assert tf.shape == (256, 5)
            _chk(self._lib.vv_set_tf(self._h, _ptr(tf)))

    def setLICParams(self, p):
        _chk(self._lib.vv_set_lic_params(self._h, ctypes.byref(p)))

    def setCamera(self, quat=(0, 0, 0, 1), pos=(0, 0, 0), dist=4.0, fovy=35.0, near=0.1, far=50.0):
        _chk(self._lib.vv_set_camera(self._h, (ctypes.c_float * 4)(*quat), (ctypes.c_float * 3)(*pos), dist, fovy, near, far))

    def setLight(self, quat=(0, 0, 0, 1), dist=1.0):
        _chk(self._lib.vv_set_light(self._h, (ctypes.c_float * 4)(*quat), dist))

    def updateLightPos(self):
        _chk(self._lib.vv_update_light_pos(self._h))

    def enableLowRes(self, enable):
        _chk(self._lib.vv_enable_lowres(self._h, int(enable)))

    def setSnapshot(self, directory=None, file_name=None, animation_on=False):
        _chk(self._lib.vv_set_snapshot(self._h, directory.encode() if directory is not None else None,
                                       file_name.encode() if file_name is not None else None, int(animation_on)))

    def screenshot(self):
        _chk(self._lib.vv_screenshot(self._h))

    def switchRecording(self):
        rc = self._lib.vv_switch_recording(self._h)
        if rc < 0:
            _chk(rc)
        return bool(rc)

    def lastSnapshotPath(self):
        return self._lib.vv_last_snapshot_path(self._h).decode()

    def setMCOffsets(self, offsets):
        """Renderer::updateMCOffsetTex with explicit values: float32 [height][width] in [0,1] (None removes the texture)"""
        if offsets is None:
            _chk(self._lib.vv_set_mc_offsets(self._h, None, 0, 0))
            return
        a = np.ascontiguousarray(offsets, dtype=np.float32)
        _chk(self._lib.vv_set_mc_offsets(self._h, _ptr(a), a.shape[1], a.shape[0]))

    def updateMCOffsetTex(self, width, height, seed=0):
        _chk(self._lib.vv_update_mc_offset_tex(self._h, width, height, seed))

    def setClipPlane(self, index, equation, active=True):
        """ClipPlane::setNormal(x, y, z, d) + activation; n.q + d >= 0 is kept (q relative to the volume centre)"""
        eq = (ctypes.c_double * 4)(*equation) if equation is not None else None
        _chk(self._lib.vv_set_clip_plane(self._h, index, eq, int(active)))

    def enableFBO(self, enable):
        _chk(self._lib.vv_enable_float_target(self._h, int(enable)))

    def setOption(self, option, value):
        _chk(self._lib.vv_set_option(self._h, option, int(value)))

    # -- loaders
    def loadDat(self, path):
        _chk(self._lib.vv_load_dat(self._h, path.encode()))

    def loadScalarDat(self, path):
        _chk(self._lib.vv_load_scalar_dat(self._h, path.encode()))

    def loadNoise(self, path, with_gradients=False):
        _chk(self._lib.vv_load_noise(self._h, path.encode(), int(with_gradients)))

    def setNoiseWithGradients(self, noise, gradients):
        """noise uint8 [nz][ny][nx] + quantised gradients uint8 [nz][ny][nx][3] (the .grd contents)"""
        a = np.ascontiguousarray(noise, dtype=np.uint8)
        g = np.ascontiguousarray(gradients, dtype=np.uint8)
        nz, ny, nx = a.shape
        assert g.shape == (nz, ny, nx, 3)
        _chk(self._lib.vv_set_noise_with_gradients(self._h, _ptr(a), _ptr(g), (ctypes.c_int * 3)(nx, ny, nz)))

    def readNoiseGradients(self, shape):
        out = np.empty(tuple(shape) + (3,), dtype=np.uint8)
        _chk(self._lib.vv_read_noise_gradients(self._h, _ptr(out), out.nbytes))
        return out

    def loadFilterPNG(self, path):
        _chk(self._lib.vv_load_filter_png(self._h, path.encode()))

    def loadTF(self, name):
        _chk(self._lib.vv_load_tf_png(self._h, name.encode()))

    # -- frame
    def render(self, update=True):
        _chk(self._lib.vv_render(self._h, int(update)))

    def readRGBA8(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _chk(self._lib.vv_read_rgba8(self._h, _ptr(out), out.nbytes))
        return out

    def readRGBA32F(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _chk(self._lib.vv_read_rgba32f(self._h, _ptr(out), out.nbytes))
        return out

    def readDisplayRGBA8(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _chk(self._lib.vv_read_display_rgba8(self._h, _ptr(out), out.nbytes))
        return out

    def readDisplayWindowRGBA8(self, win_w, win_h):
        """the displayed frame over a win_w x win_h window (low-res preset: NEAREST up-scaling of the half-size frame)"""
        out = np.empty((win_h, win_w, 4), dtype=np.uint8)
        _chk(self._lib.vv_read_display_window_rgba8(self._h, _ptr(out), out.nbytes, win_w, win_h))
        return out

    def readLICVolume(self):
        dims = (ctypes.c_int * 3)()
        ptr = ctypes.c_void_p()
        _chk(self._lib.vv_get_lic_volume_ptr(self._h, ctypes.byref(ptr), dims))
        out = np.empty((dims[2], dims[1], dims[0]), dtype=np.float32)
        _chk(self._lib.vv_read_lic_volume(self._h, _ptr(out), out.nbytes, dims))
        return out

    def debugWalk(self, pos, dir_sign, n_steps, variant=0):
        """one direction of the LIC walk from a texture-space position: rows (newPos.xyz, step.rgb, noise tap, kernel weight,
        Pos2.xyz, step2.rgb, 0, 0)"""
        out = np.zeros((n_steps, 16), dtype=np.float32)
        p = np.asarray(pos, dtype=np.float32)
        _chk(self._lib.vv_debug_walk(self._h, _ptr(p), int(dir_sign), int(n_steps), int(variant), _ptr(out), out.nbytes))
        return out

    def readFieldTexture(self, shape):
        """RGBA16F vector texture contents as float32 [z][y][x][4]"""
        out = np.empty(tuple(shape) + (4,), dtype=np.float32)
        _chk(self._lib.vv_read_field_texture(self._h, _ptr(out), out.nbytes))
        return out

    def readNoiseTexture(self, shape, channels):
        out = np.empty(tuple(shape) + ((channels,) if channels > 1 else ()), dtype=np.uint8)
        ch = ctypes.c_int()
        _chk(self._lib.vv_read_noise_texture(self._h, _ptr(out), out.nbytes, ctypes.byref(ch)))
        assert ch.value == channels
        return out

    def readSampleMap(self):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        _chk(self._lib.vv_read_sample_map(self._h, _ptr(out), out.nbytes))
        return out

    def savePNG(self, path, displayed=False):
        _chk(self._lib.vv_save_png(self._h, path.encode(), int(displayed)))

    def saveRaw(self, path):
        _chk(self._lib.vv_save_raw(self._h, path.encode()))

    # -- measurement
    def lastRaySamples(self):
        return int(self._lib.vv_last_ray_samples(self._h))

    def lastKernelMs(self):
        return float(self._lib.vv_last_kernel_ms(self._h))

    def lastLaunchCount(self):
        return int(self._lib.vv_last_launch_count(self._h))

    def fieldLayout(self):
        """the layout the vector field is packed in (OPT_FIELD_LAYOUT resolved): LAYOUT_F4 / LAYOUT_PAIR / LAYOUT_QUAD"""
        return int(self._lib.vv_field_layout(self._h))

    def synchronize(self):
        _chk(self._lib.vv_synchronize(self._h))

    # -- multi-GPU hooks
    def setPartition(self, rank, world):
        _chk(self._lib.vv_set_partition(self._h, rank, world))

    def setLICVolumeSlab(self, z0, z1):
        _chk(self._lib.vv_set_licvol_slab(self._h, z0, z1))

    def tileBuffer(self):
        """(device pointer, blocks per rank, total blocks) of the block-major tile buffer"""
        ptr, nl, nt = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _chk(self._lib.vv_get_tile_buffer(self._h, ctypes.byref(ptr), ctypes.byref(nl), ctypes.byref(nt)))
        return ptr.value, nl.value, nt.value

    def assembleTiles(self, gathered_dev_ptr, world):
        _chk(self._lib.vv_assemble_tiles(self._h, ctypes.c_void_p(gathered_dev_ptr), world))

    # ---- peer-to-peer frame exchange (see include/vv_c_api.h: vv_p2p_*) ----
    def p2pExport(self):
        """allocates the gather buffer; returns (64-byte cudaIpcMemHandle_t, base device pointer)"""
        h = (ctypes.c_ubyte * 64)()
        base = ctypes.c_void_p()
        _chk(self._lib.vv_p2p_export(self._h, h, ctypes.byref(base)))
        return bytes(h), base.value

    def p2pConnect(self, handles=None, local_bases=None):
        """handles: list of `world` 64-byte IPC handles in rank order; local_bases: base pointers of ranks in this process"""
        world = len(handles) if handles is not None else len(local_bases)
        hb = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles)) if handles is not None else None
        lb = (ctypes.c_void_p * world)(*[ctypes.c_void_p(b) for b in local_bases]) if local_bases is not None else None
        _chk(self._lib.vv_p2p_connect(self._h, hb, lb, world))

    def p2pRender(self):
        _chk(self._lib.vv_p2p_render(self._h))

    def p2pStatus(self):
        _chk(self._lib.vv_p2p_status(self._h))

    def p2pDisconnect(self):
        _chk(self._lib.vv_p2p_disconnect(self._h))

    def licVolumePtr(self):
        dims = (ctypes.c_int * 3)()
        ptr = ctypes.c_void_p()
        _chk(self._lib.vv_get_lic_volume_ptr(self._h, ctypes.byref(ptr), dims))
        return ptr.value, (dims[0], dims[1], dims[2])

    def setStream(self, cuda_stream):
        _chk(self._lib.vv_set_stream(self._h, ctypes.c_void_p(cuda_stream) if cuda_stream else None))
