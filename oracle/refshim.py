"""ctypes wrapper of oracle/_ref/libvv_ref.so -- TEST INFRASTRUCTURE ONLY.

libvv_ref.so is the reference's own GLSL shader source (VV/shader/*.glsl) compiled as C++ through oracle/glsl_shim.h
by oracle/build_ref.py.  It exists only where /root/reference was available at build time; the built .so travels to
the GPU box with the repo snapshot (git-ignored, not gpurun-ignored).
"""
import ctypes
import os

import numpy as np

from . import vvo

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libvv_ref.so")

W_CLAMP_TO_EDGE, W_REPEAT, W_CLAMP = 0, 1, 2
F_L8, F_LA8, F_RGBA8, F_RGBA32F, F_L32F, F_LA32F = range(6)


class RefTex(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dim", ctypes.c_int * 3), ("fmt", ctypes.c_int), ("wrap", ctypes.c_int)]


class RefUniforms(ctypes.Structure):
    _fields_ = [(n, RefTex) for n in ("volume", "scalar", "noise", "kernel", "tf_rgba", "tf_alphaopac", "licvol", "zoeckler",
                                      "mallo_diff", "mallo_spec")] + [
        ("texMax", ctypes.c_float * 4), ("scaleVol", ctypes.c_float * 4), ("scaleVolInv", ctypes.c_float * 4),
        ("stepSize", ctypes.c_float), ("gradient", ctypes.c_float * 3), ("numIterations", ctypes.c_int),
        ("alphaCorrection", ctypes.c_float), ("licParams", ctypes.c_float * 3), ("licKernel", ctypes.c_float * 3),
        ("camera", ctypes.c_float * 4), ("light_position", ctypes.c_float * 4), ("light_ambient", ctypes.c_float * 4),
        ("light_diffuse", ctypes.c_float * 4), ("light_specular", ctypes.c_float * 4), ("spot_exponent", ctypes.c_float),
        ("mc_offset", RefTex), ("frag_x0", ctypes.c_int), ("frag_y0", ctypes.c_int), ("frag_w", ctypes.c_int),
        ("slice_count", ctypes.c_int), ("fbo_fp16", ctypes.c_int)]


_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libvv_ref.so not built (python oracle/build_ref.py, needs /root/reference)")
        _lib = ctypes.CDLL(LIB)
        _lib.vvref_set_gl_state.argtypes = [ctypes.POINTER(RefUniforms)]
    return _lib


def _tex(arr, dims, fmt, wrap):
    t = RefTex()
    t.data = arr.ctypes.data if arr is not None else None
    t.dim = (ctypes.c_int * 3)(*dims)
    t.fmt = fmt
    t.wrap = wrap
    return t


class RefScene:
    """Runs the reference's shader code on the inputs of a configs.Scene.  Texture CONTENTS and uniform VALUES come
    from the oracle's restatement of the host code (vvo.OracleScene); the shading is the reference's."""

    def __init__(self, s, luminance_noise=False, illum_tables=None):
        self.s = s
        self.o = vvo.OracleScene(s, illum_tables=illum_tables)
        o, c = self.o, self.o.c
        u = RefUniforms()
        nz, ny, nx = s.field.shape[:3]
        u.volume = _tex(o.vec, (nx, ny, nz), F_RGBA32F, W_CLAMP_TO_EDGE)
        if s.scalar is not None:
            u.scalar = _tex(o.scalar, o.scalar.shape[::-1], F_L8, W_CLAMP_TO_EDGE)
        else:
            self._zero = np.zeros((1, 1, 1), np.uint8)
            u.scalar = _tex(self._zero, (1, 1, 1), F_L8, W_CLAMP_TO_EDGE)
        noise = np.ascontiguousarray(s.noise, dtype=np.uint8)
        ndim = noise.shape[::-1]
        if s.with_gradients or not (s.quirk_luminance_alpha or luminance_noise):
            # the README always passes -g: RGBA8 texture (gradient.xyz, noise) -> .a is the noise value.
            # Without -g the texture is GL_LUMINANCE and .a == 1 (Q7), used only when that quirk is requested.
            grad = o.grad if s.with_gradients else np.zeros(noise.shape + (3,), np.uint8)
            self.noise = vvo.pack_noise_rgba(noise, grad)
            u.noise = _tex(self.noise, ndim, F_RGBA8, W_REPEAT)
        else:
            self.noise = noise
            u.noise = _tex(self.noise, ndim, F_L8, W_REPEAT)
        u.kernel = _tex(o.kernel, (o.kernel.size, 1, 1), F_L8, W_CLAMP)
        self.tf_rgba = np.ascontiguousarray(o.tf[:, :4])
        self.tf_ao = np.ascontiguousarray(o.tf[:, 3:5])
        u.tf_rgba = _tex(self.tf_rgba, (256, 1, 1), F_RGBA8, W_CLAMP_TO_EDGE)
        u.tf_alphaopac = _tex(self.tf_ao, (256, 1, 1), F_LA8, W_CLAMP_TO_EDGE)
        if illum_tables is not None:
            z, md, ms = o.illum_tables
            u.zoeckler = _tex(z, (z.shape[1], z.shape[0], 1), F_LA32F, W_CLAMP_TO_EDGE)
            u.mallo_diff = _tex(md, (md.shape[1], md.shape[0], 1), F_L32F, W_CLAMP_TO_EDGE)
            u.mallo_spec = _tex(ms, (ms.shape[1], ms.shape[0], 1), F_L32F, W_CLAMP_TO_EDGE)
        un = o.uniforms()
        u.stepSize = un[0]
        u.gradient = (ctypes.c_float * 3)(*un[1:4])
        u.licParams = (ctypes.c_float * 3)(*un[4:7])
        u.licKernel = (ctypes.c_float * 3)(*un[7:10])
        u.alphaCorrection = un[10]
        u.numIterations = int(un[11])
        cam = np.zeros(3, np.float32)
        rot = np.zeros(9, np.float32)
        vvo.lib().vvo_view(ctypes.byref(c), vvo._p(cam), vvo._p(rot))
        u.camera = (ctypes.c_float * 4)(cam[0], cam[1], cam[2], 1.0)
        lp = np.zeros(4, np.float32)
        vvo.lib().vvo_light_position(ctypes.byref(c), vvo._p(lp))
        u.light_position = (ctypes.c_float * 4)(*lp)
        u.light_ambient = (ctypes.c_float * 4)(0, 0, 0, 1)      # GL defaults of LIGHT0
        u.light_diffuse = (ctypes.c_float * 4)(1, 1, 1, 1)
        u.light_specular = (ctypes.c_float * 4)(1, 1, 1, 1)
        u.spot_exponent = 40.0                                   # VV/3DLIC.cpp:736, VV/illumination.h:52
        if getattr(s, "mc_offsets", None) is not None:
            # mcOffsetSampler: GL_LUMINANCE16F_ARB rectangle texture, NEAREST (VV/renderer.cpp:636-679)
            self.mc = vvo.half_round(np.ascontiguousarray(s.mc_offsets, dtype=np.float32).reshape(s.height, s.width))
            u.mc_offset = _tex(self.mc, (s.width, s.height, 1), F_L32F, W_CLAMP_TO_EDGE)
        self.u = u
        self._set_scale(raycast=True)

    def _set_scale(self, raycast):
        """scaleVol / scaleVolInv / texMax as the program sees them: Q1 applies to the ray-cast program in ILLUM_* builds;
        the LIC-volume and volume-ray-cast programs never have scaleVolInv active (VV/renderer.cpp:807-922, 941-944)"""
        c = self.o.c
        sc = np.zeros(9, np.float32)
        if raycast:
            vvo.lib().vvo_scale_uniforms(ctypes.byref(c), vvo._p(sc))
        else:
            sc[0:3] = list(c.scale); sc[3:6] = list(c.scale_inv)
            sc[6:9] = [c.extent[i] * c.scale[i] for i in range(3)]
        self.u.scaleVol = (ctypes.c_float * 4)(sc[0], sc[1], sc[2], 0.0)
        self.u.scaleVolInv = (ctypes.c_float * 4)(sc[3], sc[4], sc[5], 1.0 if not raycast else 0.0)
        self.u.texMax = (ctypes.c_float * 4)(sc[6], sc[7], sc[8], 0.0)

    def program(self, kind):
        d = self.s.defines or ""
        if kind == "raycast":
            base = "raycast_none"
            for k, v in (("ILLUM_GRADIENT", "gradient"), ("ILLUM_MALLO", "mallo"), ("ILLUM_ZOECKLER", "zoeckler"), ("SPEED_OF_FLOW", "sof")):
                if k in d:
                    base = "raycast_" + v
            if self.s.gate_mode == 1:
                base += "_gatetf"
            tfm = {0: "", 1: "_tfa", 2: "_tfr", 3: "_tflength", 4: "_tfscalar"}[self.s.tf_mode]
            if getattr(self.s, "mc_offsets", None) is not None:
                assert base + tfm in ("raycast_none", "raycast_gradient"), "MC-offset variants are built for the .b / always-gate programs"
                return base + "_mc"
            return base + tfm
        if kind == "licvol":
            if "ILLUM_GRADIENT" in d:
                return "licvol_gradient"
            return "licvol_sof" if "SPEED_OF_FLOW" in d else "licvol_none"
        return "volraycast"

    def _run(self, prog, tc):
        L = lib()
        fn = getattr(L, "vvref_run_" + prog)
        fn.argtypes = [ctypes.POINTER(RefUniforms), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        n = tc.shape[0]
        out = np.zeros((n, 4), np.float32)
        cnt = np.zeros(n, np.uint32)
        L.vvref_set_gl_state(ctypes.byref(self.u))
        fn(ctypes.byref(self.u), vvo._p(tc), n, vvo._p(out), vvo._p(cnt))
        return out, cnt

    def _rays(self, rect):
        s = self.s
        x0, y0, x1, y1 = rect
        tc = np.zeros(((y1 - y0) * (x1 - x0), 4), np.float32)
        vvo.lib().vvo_pixel_rays(ctypes.byref(self.o.c), x0, y0, x1, y1, vvo._p(tc))
        return tc

    def raycast(self, rect=None, texcoords=None):
        """texcoords: optional [h][w][4] fragment texcoord0 (w = 1 where the pixel has a fragment) to shade instead of the
        oracle's analytic ray entry points -- e.g. the fragments oracle/softgl.py rasterises from the reference's own draw calls"""
        s = self.s
        rect = rect or (0, 0, s.width, s.height)
        x0, y0, x1, y1 = rect
        self._set_scale(raycast=True)
        self.u.frag_x0, self.u.frag_y0, self.u.frag_w = x0, y0, x1 - x0
        tc = self._rays(rect) if texcoords is None else np.ascontiguousarray(np.asarray(texcoords, np.float32)[y0:y1, x0:x1].reshape(-1, 4))
        out, cnt = self._run(self.program("raycast"), tc)
        img = np.zeros((s.height, s.width, 4), np.float32)
        cm = np.zeros((s.height, s.width), np.uint32)
        img[y0:y1, x0:x1] = out.reshape(y1 - y0, x1 - x0, 4)
        cm[y0:y1, x0:x1] = cnt.reshape(y1 - y0, x1 - x0)
        return img, cm, int(cnt.sum())

    def slicing(self, fragments=None, fbo_fp16=1, fbo_pingpong=1):
        """lic3d_slicing_fragment.glsl over the oracle's slice geometry; the shader hard-codes TF index .a and the
        tfData.a > 0.05 gate, so the scene must use tf_mode A / gate TF_ALPHA.
        fragments: optional (starts int32 [h*w + 1], frags [n][4] = texcoord.xyz + slice index) per-pixel fragment lists in draw
        order to shade instead (oracle/softgl.py fragment_lists of the reference's own slice polygons).
        fbo_fp16 / fbo_pingpong: the frame-buffer side as Renderer::sliceVolume runs it (see ref_api.h); 0 / 0 = one fp32 accumulator"""
        s = self.s
        assert s.tf_mode == 1 and s.gate_mode == 1
        d = s.defines or ""
        prog = "slicing_none"
        for k, v in (("ILLUM_GRADIENT", "gradient"), ("ILLUM_MALLO", "mallo"), ("ILLUM_ZOECKLER", "zoeckler")):
            if k in d:
                prog = "slicing_" + v
        if getattr(s, "mc_offsets", None) is not None:
            assert prog == "slicing_none"
            prog = "slicing_none_mc"
        self._set_scale(raycast=True)
        self.u.frag_x0, self.u.frag_y0, self.u.frag_w = 0, 0, s.width
        L = lib()
        if fragments is not None:
            st = np.ascontiguousarray(fragments[0], np.int32)
            fr = np.ascontiguousarray(fragments[1], np.float32)
            assert fr.shape[1] == 4
        _, _, nslices = self.o.slicing_setup()
        self.u.slice_count = int(nslices) if fbo_pingpong else 0
        self.u.fbo_fp16 = int(fbo_fp16)
        if fragments is None:
            frags, starts = [], [0]
            buf = np.zeros((nslices, 4), np.float32)
            for y in range(s.height):
                for x in range(s.width):
                    n = vvo.lib().vvo_slice_fragments(ctypes.byref(self.o.c), x, y, vvo._p(buf), nslices)
                    frags.append(buf[:n].copy())
                    starts.append(starts[-1] + n)
            fr = np.ascontiguousarray(np.concatenate(frags, axis=0) if frags else np.zeros((0, 4), np.float32))
            st = np.asarray(starts, np.int32)
        npix = s.width * s.height
        out = np.zeros((npix, 4), np.float32)
        cnt = np.zeros(npix, np.uint32)
        fn = getattr(L, "vvref_slice_" + prog)
        fn.argtypes = [ctypes.POINTER(RefUniforms), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.vvref_set_gl_state(ctypes.byref(self.u))
        fn(ctypes.byref(self.u), vvo._p(fr), vvo._p(st), npix, vvo._p(out), vvo._p(cnt))
        return out.reshape(s.height, s.width, 4), cnt.reshape(s.height, s.width), int(cnt.sum())

    def slicing_blend_fragments(self, frags):
        """lic3d_slicingblend_fragment.glsl (slicing without the FBO) on fragment positions frags [n][3]: the premultiplied sample
        colour of every fragment, [n][4], as the shader leaves it in gl_FragColor (the GL clamps and blends it afterwards)"""
        s = self.s
        assert s.tf_mode == 1 and s.gate_mode == 1
        d = s.defines or ""
        prog = "slicingblend_none"
        for k, v in (("ILLUM_GRADIENT", "gradient"), ("ILLUM_MALLO", "mallo")):
            if k in d:
                prog = "slicingblend_" + v
        tc = np.zeros((len(frags), 4), np.float32)
        tc[:, :3] = frags
        tc[:, 3] = 1.0
        self._set_scale(raycast=True)
        self.u.frag_w = 0
        out, _ = self._run(prog, np.ascontiguousarray(tc))
        return out

    def lic_volume(self, dims=None):
        nz, ny, nx = self.s.field.shape[:3]
        w, h, d = dims or (nx, ny, nz)
        z, y, x = np.meshgrid(np.arange(d, dtype=np.float32), np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
        tc = np.stack([(x + np.float32(0.5)) / np.float32(w), (y + np.float32(0.5)) / np.float32(h), (z + np.float32(0.5)) / np.float32(d),
                       np.ones_like(x)], axis=-1).reshape(-1, 4).astype(np.float32)
        self._set_scale(raycast=False)
        out, _ = self._run(self.program("licvol"), np.ascontiguousarray(tc))
        return out[:, 0].reshape(d, h, w).copy()

    def raycast_licvolume(self, licvol):
        s = self.s
        self.licvol = np.ascontiguousarray(licvol, dtype=np.float32)
        self.u.licvol = _tex(self.licvol, self.licvol.shape[::-1], F_L32F, W_REPEAT)
        self._set_scale(raycast=False)
        out, cnt = self._run("volraycast", self._rays((0, 0, s.width, s.height)))
        return out.reshape(s.height, s.width, 4), cnt.reshape(s.height, s.width), int(cnt.sum())


def background(img, win_w=None, win_h=None):
    """background_fragment.glsl (the display pass of Renderer::renderBackground, VV/renderer.cpp:1407-1476) over a window of
    win_w x win_h pixels showing the stored frame img [rh][rw][4]; window = frame unless the low-res preset halves the frame"""
    img = np.ascontiguousarray(img, np.float32)
    rh, rw = img.shape[:2]
    ww, wh = win_w or rw, win_h or rh
    out = np.zeros((wh, ww, 4), np.float32)
    L = lib()
    L.vvref_run_background.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.vvref_run_background(img.ctypes.data, rw, rh, ww, wh, out.ctypes.data)
    return out
