"""ctypes wrapper of the reference HOST code inside oracle/_ref/libvv_ref.so -- TEST INFRASTRUCTURE ONLY.
(VV/dataset.cpp, gradient.cpp, reader.cpp, parseArg.cpp, transferEdit.cpp, mmath.cpp compiled unmodified against the
capturing GL stub of oracle/ref_shim/; see oracle/ref_host_driver.cpp.)"""
import ctypes

import numpy as np

from . import refshim

GL_CLAMP, GL_REPEAT, GL_CLAMP_TO_EDGE = 0x2900, 0x2901, 0x812F
GL_RGBA, GL_LUMINANCE, GL_RGBA16F_ARB = 0x1908, 0x1909, 0x881A


def available():
    if not refshim.available():
        return False
    try:
        return bool(refshim.lib().vvref_has_host())
    except AttributeError:
        return False


def _L():
    L = refshim.lib()
    L.vvref_vector_texture.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_noise_texture.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_noise_texture_cached.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_filter_texture.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_tf_textures.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_noise_gradients.argtypes = [ctypes.c_void_p] * 6
    L.vvref_parse_dat.argtypes = [ctypes.c_char_p] + [ctypes.c_void_p] * 6
    L.vvref_parse_args.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_slicing_setup.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    L.vvref_slice_polygon.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.vvref_clip_cap_polygon.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.vvref_renderer_state.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p,
                                       ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.vvref_cube_faces.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.vvref_quat_angle_axis.argtypes = [ctypes.c_void_p] * 3
    L.vvref_quat_mult_vec.argtypes = [ctypes.c_void_p] * 3
    return L


def vector_texture(dat_path, shape, interp=(0, 10)):
    """returns (float RGBA [z][y][x][4] as handed to glTexImage3D, geometry dict, internal format, wrap)"""
    out = np.zeros(tuple(shape) + (4,), np.float32)
    dims = (ctypes.c_int * 3)()
    geom = np.zeros(17, np.float32)
    ifmt, wrap = ctypes.c_int(), ctypes.c_int()
    rc = _L().vvref_vector_texture(dat_path.encode(), interp[0], interp[1], out.ctypes.data, out.nbytes, dims, geom.ctypes.data,
                                   ctypes.byref(ifmt), ctypes.byref(wrap))
    assert rc > 0, rc
    assert tuple(dims) == tuple(shape[::-1])
    g = dict(extent=geom[0:3], center=geom[3:6], scale=geom[6:10], scale_inv=geom[10:14], slice_dist=geom[14:17])
    return out, g, ifmt.value, wrap.value


def noise_texture(path, shape, use_gradient):
    ch = 4 if use_gradient else 1
    out = np.zeros(tuple(shape) + ((4,) if use_gradient else ()), np.uint8)
    dims = (ctypes.c_int * 3)()
    ifmt, wrap = ctypes.c_int(), ctypes.c_int()
    rc = _L().vvref_noise_texture(path.encode(), int(use_gradient), out.ctypes.data, out.nbytes, dims, ctypes.byref(ifmt), ctypes.byref(wrap))
    assert rc == out.nbytes, rc
    return out, ifmt.value, wrap.value


def noise_texture_cached(path, shape):
    """NoiseDataSet::createTexture with gradients, leaving / using the <path>.grd cache"""
    out = np.zeros(tuple(shape) + (4,), np.uint8)
    dims = (ctypes.c_int * 3)()
    ifmt, wrap = ctypes.c_int(), ctypes.c_int()
    rc = _L().vvref_noise_texture_cached(path.encode(), out.ctypes.data, out.nbytes, dims, ctypes.byref(ifmt), ctypes.byref(wrap))
    assert rc == out.nbytes, rc
    return out


def filter_texture(png_path=None):
    out = np.zeros(4096, np.uint8)
    width, inv, wrap = ctypes.c_int(), ctypes.c_float(), ctypes.c_int()
    rc = _L().vvref_filter_texture(png_path.encode() if png_path else None, out.ctypes.data, out.nbytes, ctypes.byref(width),
                                   ctypes.byref(inv), ctypes.byref(wrap))
    assert rc == width.value, rc
    return out[:width.value].copy(), inv.value, wrap.value


def tf_textures(name=None):
    rgba = np.zeros((256, 4), np.uint8)
    la = np.zeros((256, 2), np.uint8)
    loaded = ctypes.c_int()
    rc = _L().vvref_tf_textures(name.encode() if name else None, rgba.ctypes.data, la.ctypes.data, ctypes.byref(loaded))
    assert rc == 0
    return rgba, la, bool(loaded.value)


def noise_gradients(noise, slice_dist=(1.0, 1.0, 1.0)):
    noise = np.ascontiguousarray(noise, np.uint8)
    nz, ny, nx = noise.shape
    g = np.zeros((nz, ny, nx, 3), np.float32)
    f = np.zeros_like(g)
    q = np.zeros((nz, ny, nx, 3), np.uint8)
    dims = (ctypes.c_int * 3)(nx, ny, nz)
    sd = (ctypes.c_float * 3)(*slice_dist)
    _L().vvref_noise_gradients(noise.ctypes.data, dims, sd, g.ctypes.data, f.ctypes.data, q.ctypes.data)
    return g, f, q


def parse_dat(path):
    res, dist = (ctypes.c_int * 3)(), (ctypes.c_float * 3)()
    dt, dd, tb, te = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _L().vvref_parse_dat(path.encode(), res, dist, ctypes.byref(dt), ctypes.byref(dd), ctypes.byref(tb), ctypes.byref(te))
    if rc:
        return None
    return dict(resolution=tuple(res), slice_thickness=tuple(dist), data_type=dt.value, data_dim=dd.value, time=(tb.value, te.value))


def parse_args(argv):
    arr = (ctypes.c_char_p * len(argv))(*[a.encode() for a in argv])
    out = ctypes.create_string_buffer(6 * 512)
    flags = (ctypes.c_int * 2)()
    rc = _L().vvref_parse_args(len(argv), arr, out, flags)
    names = [out.raw[512 * i:512 * (i + 1)].split(b"\0")[0].decode() for i in range(6)]
    return rc == 0, dict(vol=names[0], noise=names[1], tf=names[2], filter=names[3], redirect=names[4], halton=names[5],
                         gradients=bool(flags[0]), lambda2=bool(flags[1]))


def quat_angle_axis(q):
    q = np.asarray(q, np.float32)
    ang = ctypes.c_float()
    ax = np.zeros(3, np.float32)
    _L().vvref_quat_angle_axis(q.ctypes.data, ctypes.byref(ang), ax.ctypes.data)
    return ang.value, ax


def quat_mult_vec(q, v):
    q = np.asarray(q, np.float32); v = np.asarray(v, np.float32)
    out = np.zeros(3, np.float32)
    _L().vvref_quat_mult_vec(q.ctypes.data, v.ctypes.data, out.ctypes.data)
    return out


def illum_tables():
    """(zoeckler float [256][256][2], mallo diffuse [256][256][4], mallo specular [256][256][4], internal formats, specExp)
    exactly as Illumination::createIllumTextures hands them to glTexImage2D"""
    L = _L()
    L.vvref_illum_tables.argtypes = [ctypes.c_void_p] * 6
    z = np.zeros((256, 256, 2), np.float32); d = np.zeros((256, 256, 4), np.float32); s = np.zeros((256, 256, 4), np.float32)
    dims = (ctypes.c_int * 2)(); ifmt = (ctypes.c_int * 3)(); se = ctypes.c_float()
    rc = L.vvref_illum_tables(z.ctypes.data, d.ctypes.data, s.ctypes.data, dims, ifmt, ctypes.byref(se))
    assert rc == 0 and tuple(dims) == (256, 256)
    return z, d, s, tuple(ifmt), se.value


def slicing_setup(mv, samp_dist, extent):
    """ViewSlicing::setupSlicing (VV/slicing.cpp:42-114): returns (v[3], d, numSlices)"""
    m = np.ascontiguousarray(mv, np.float32).reshape(16)
    e = np.ascontiguousarray(extent, np.float32)
    out = np.zeros(5, np.float32)
    n = _L().vvref_slicing_setup(m.ctypes.data, float(samp_dist), e.ctypes.data, out.ctypes.data)
    return out[:3].copy(), float(out[3]), n


def slice_polygon(mv, samp_dist, extent, slice_index):
    """the vertices ViewSlicing::drawSlice(slice) emits: (verts [n][3], texcoords [n][3]) in volume coordinates"""
    m = np.ascontiguousarray(mv, np.float32).reshape(16)
    e = np.ascontiguousarray(extent, np.float32)
    v = np.zeros((8, 3), np.float32); t = np.zeros((8, 3), np.float32)
    n = _L().vvref_slice_polygon(m.ctypes.data, float(samp_dist), e.ctypes.data, int(slice_index), v.ctypes.data, t.ctypes.data, 8)
    assert n >= 0
    return v[:n].copy(), t[:n].copy()


def clip_cap_polygon(plane, extent):
    """ClipPlane::drawSlice (VV/transform.cpp:432-444) for the plane (n.xyz, d): (verts, texcoords)"""
    pl = np.ascontiguousarray(plane, np.float64)
    e = np.ascontiguousarray(extent, np.float32)
    v = np.zeros((8, 3), np.float32); t = np.zeros((8, 3), np.float32)
    n = _L().vvref_clip_cap_polygon(pl.ctypes.data, e.ctypes.data, v.ctypes.data, t.ctypes.data, 8)
    assert n >= 0
    return v[:n].copy(), t[:n].copy()


def renderer_state(dat_path, filter_png, camera, light, lic_params, lowres, has_scalevolinv, width, height):
    """Renderer / Camera / Transform driven as VV/3DLIC.cpp does, GL calls captured (oracle/ref_host_driver.cpp):
    returns dict(modelview[16], light_position[4], slicing=(v[3], d, n), uniforms={name: 4 floats})"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    q, pos = f32(camera["quat"]), f32(camera["pos"])
    lq = f32(light["quat"])
    lp = f32([lic_params.stepSizeVol, lic_params.gradientScale, lic_params.illumScale, lic_params.freqScale, lic_params.numIterations,
              lic_params.stepsForward, lic_params.stepsBackward, lic_params.stepSizeLIC])
    out = np.zeros(25 + 40, np.float32)
    rc = _L().vvref_renderer_state(dat_path.encode(), filter_png.encode() if filter_png else None, q.ctypes.data, pos.ctypes.data,
                                   float(camera["dist"]), lq.ctypes.data, float(light["dist"]), lp.ctypes.data, int(lowres),
                                   int(has_scalevolinv), int(width), int(height), out.ctypes.data)
    assert rc == 0, rc
    names = ["texMax", "scaleVol", "scaleVolInv", "stepSize", "gradient", "licParams", "licKernel", "numIterations", "alphaCorrection", "viewport"]
    return dict(modelview=out[:16].copy(), light_position=out[16:20].copy(), slicing=(out[20:23].copy(), float(out[23]), int(out[24])),
                uniforms={n: out[25 + 4 * i: 29 + 4 * i].copy() for i, n in enumerate(names)})


def cube_faces(dat_path):
    """Renderer::drawCubeFaces: (verts [24][3], texcoords [24][3]), four vertices per quad"""
    v = np.zeros((24, 3), np.float32); t = np.zeros((24, 3), np.float32)
    n = _L().vvref_cube_faces(dat_path.encode(), v.ctypes.data, t.ctypes.data, 24)
    assert n == 24, n
    return v, t


def raycast_draws(dat_path, camera, width, height, lowres=0, planes=(), frames=2, slicing=False, step_size_vol=0.0):
    """Renderer::render(true) in ray-cast mode, run unmodified with the GL calls captured: the list of glBegin/glEnd primitives,
    each a dict(program, mode, cull, clip_mask, blend, blend_func, viewport[4], modelview[4][4], projection[4][4], clip_eye[6][4],
    verts[n][3], tex[n][3]).  program 77 = the LIC ray-cast program, 79 = the slicing program, 78 = the background program, 0 = fixed function.
    slicing: 0 / False = VOLIC_RAYCAST, 1 / True = VOLIC_SLICING, 2 = VOLIC_LICVOLUME (program 81).
    frames: render(true) is called that many times and the last frame is returned (2 = the steady state, see the driver)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    q, pos = f32(camera["quat"]), f32(camera["pos"])
    pl = np.zeros((3, 4), np.float64)
    pl[:, 2] = -1.0
    act = (ctypes.c_int * 3)(0, 0, 0)
    for i, p in enumerate(planes):
        pl[i] = p
        act[i] = 1
    cap = 1 << 20
    out = np.zeros(cap, np.float64)
    L = _L()
    L.vvref_raycast_draws.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_void_p, ctypes.c_int]
    n = L.vvref_raycast_draws(dat_path.encode(), q.ctypes.data, pos.ctypes.data, float(camera["dist"]), int(width), int(height),
                              int(lowres), pl.ctypes.data, act, int(frames), int(slicing), float(step_size_vol),
                              out.ctypes.data, cap)
    assert n > 0, n
    return _parse_draws(out, n)


def licvolume_draws(w, h, d):
    """Renderer::updateLICVolume (VV/renderer.cpp:1311-1374) into a w x h x d VolumeBuffer: the captured primitives, same format
    as raycast_draws; the LIC-volume program has the handle 80"""
    cap = 1 << 20
    out = np.zeros(cap, np.float64)
    L = _L()
    L.vvref_licvolume_draws.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    n = L.vvref_licvolume_draws(int(w), int(h), int(d), out.ctypes.data, cap)
    assert n > 0, n
    return _parse_draws(out, n)


def _parse_draws(out, n):
    k = 1
    draws = []
    for _ in range(int(out[0])):
        d = dict(program=int(out[k]), mode=int(out[k + 1]), cull=int(out[k + 2]), clip_mask=int(out[k + 3]),
                 blend=int(out[k + 4]), blend_func=(int(out[k + 5]), int(out[k + 6])), fbo_tex=int(out[k + 7]), bound_tex=int(out[k + 8]),
                 viewport=[int(x) for x in out[k + 9:k + 13]])
        k += 13
        d["modelview"] = out[k:k + 16].reshape(4, 4).T.copy(); k += 16          # column-major -> [row][col]
        d["projection"] = out[k:k + 16].reshape(4, 4).T.copy(); k += 16
        d["clip_eye"] = out[k:k + 24].reshape(6, 4).copy(); k += 24
        nv = int(out[k]); k += 1
        v = out[k:k + 6 * nv].reshape(nv, 6); k += 6 * nv
        d["verts"], d["tex"] = v[:, :3].copy(), v[:, 3:].copy()
        draws.append(d)
    assert k == n
    return draws


def has_app():
    try:
        return available() and bool(_L().vvref_has_app())
    except AttributeError:
        return False


def app_keyboard(dat_path, keys, ref_dir="/root/reference/VectorVisualization"):
    """keyboard() / keyboardSpecial() of VV/3DLIC.cpp (compiled unmodified, oracle/ref_app_driver.cpp) fed with `keys` -- a
    list of (key, special) -- from the application's start-up state.  Returns dict(lic=(stepSizeVol, gradientScale, illumScale,
    freqScale, numIterations, stepsForward, stepsBackward, stepSizeLIC), technique, lowres, fbo, recording, animation,
    clip_active[3], selected_clip, screenshot, shader_loads, continuous, store_frame, defines, hud)"""
    n = len(keys)
    k = (ctypes.c_ubyte * max(n, 1))(*[(x if isinstance(x, int) else ord(x)) & 0xff for x, _ in keys])
    sp = (ctypes.c_int * max(n, 1))(*[int(bool(f)) for _, f in keys])
    out = (ctypes.c_float * 21)()
    d = ctypes.create_string_buffer(256)
    h = ctypes.create_string_buffer(1024)
    rc = _L().vvref_keyboard(dat_path.encode(), ref_dir.encode() if ref_dir else None, k, sp, n, out, d, 256, h, 1024)
    assert rc == 0, rc
    o = list(out)
    return dict(lic=(np.float32(o[0]), np.float32(o[1]), np.float32(o[2]), np.float32(o[3]), int(o[4]), int(o[5]), int(o[6]), np.float32(o[7])),
                technique=int(o[8]), lowres=int(o[9]), fbo=int(o[10]), recording=int(o[11]), animation=int(o[12]),
                clip_active=[int(x) for x in o[13:16]], selected_clip=int(o[16]), screenshot=int(o[17]), shader_loads=int(o[18]),
                continuous=int(o[19]), store_frame=int(o[20]), defines=d.value.decode().strip(), hud=h.value.decode())


def animation_ticks(dat_path, n, tex_tick=-1, tex_shape=None):
    """init() + n idle() ticks of the reference's animation on a time-dependent DAT file (oracle/ref_app_driver.cpp): returns
    ([(current time step, interpIndex the tick's texture is packed with, moved on), ...], texture of tick tex_tick or None)"""
    out = (ctypes.c_int * (3 * n))()
    tex = np.zeros(tuple(tex_shape) + (4,), np.float32) if tex_shape is not None else None
    L = _L()
    L.vvref_animation_ticks.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    rc = L.vvref_animation_ticks(dat_path.encode(), n, out, tex_tick, tex.ctypes.data if tex is not None else None, tex.nbytes if tex is not None else 0)
    assert rc == 0, rc
    return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(n)], tex


def app_mouse(dat_path, width, height, events):
    """mouseInteract / mouseMotionInteract of VV/3DLIC.cpp fed with events (type, button, x, y, modifiers) -- type 0 press, 1 motion,
    2 select clip plane `button` -- on freshly constructed camera / light / clip planes.  Returns (objects float32 [5][13]: camera,
    light, clip planes 0..2 as (_q_internal[4], _q[4], _dist, _pos[3], locked), plane equations float64 [3][4])"""
    ev = np.ascontiguousarray(np.asarray(events, np.int32).reshape(-1, 5))
    out = np.zeros((5, 13), np.float32)
    normals = np.zeros((3, 4), np.float64)
    L = _L()
    L.vvref_mouse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    rc = L.vvref_mouse(dat_path.encode(), int(width), int(height), ev.ctypes.data, len(ev), out.ctypes.data, normals.ctypes.data)
    assert rc == 0, rc
    return out, normals


def scalar_texture(dat_path, shape, dtype):
    """VolumeDataSet::loadData + createTexture: (data as handed to glTexImage3D, internal format, wrap, GL source type, LINEAR?)"""
    out = np.zeros(tuple(shape), dtype)
    dims = (ctypes.c_int * 3)()
    ifmt, wrap, typ, lin = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _L().vvref_scalar_texture(dat_path.encode(), out.ctypes.data, out.nbytes, dims, ctypes.byref(ifmt), ctypes.byref(wrap),
                                   ctypes.byref(typ), ctypes.byref(lin))
    assert rc == out.nbytes, rc
    assert tuple(dims) == tuple(shape[::-1])
    return out, ifmt.value, wrap.value, typ.value, bool(lin.value)


def upload_log_reset():
    _L().vvref_upload_log_reset()


def uploaded_textures():
    """state of every texture the reference uploaded since upload_log_reset(), in glTexImage* order: list of dicts(target,
    internal_format, format, type, min_filter, mag_filter, wrap_s, wrap_t, wrap_r, dims)"""
    ids = (ctypes.c_uint * 256)()
    n = _L().vvref_upload_log(ids, 256)
    out = []
    for k in range(min(n, 256)):
        st = (ctypes.c_int * 12)()
        if _L().vvref_texture_state(ids[k], st) != 0:
            continue                                   # deleted since
        out.append(dict(id=int(ids[k]), target=st[0], internal_format=st[1], format=st[2], type=st[3], min_filter=st[4], mag_filter=st[5],
                        wrap_s=st[6], wrap_t=st[7], wrap_r=st[8], dims=(st[9], st[10], st[11])))
    return out


def renderer_textures(width, height, lic_dims):
    """the textures Renderer creates itself (FBO colour targets, MC-offset texture, LIC volume buffer layers)"""
    assert _L().vvref_renderer_textures(int(width), int(height), *[int(v) for v in lic_dims]) == 0
