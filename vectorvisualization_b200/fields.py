"""Synthetic inputs for the 3D-LIC path.

The reference ships no data, noise, kernels or colour tables (VV/README.txt:3-4 refers to ../data, ../noise,
../kernel, ../colortables which are not in its tree), only the file formats of VV/README.txt:19-75.  These
generators are builder-defined and deterministic; every consumer (CUDA path, oracle, reference shim) reads the
byte-identical arrays they return.  Writers emit the reference's own on-disk formats so its loaders' semantics
(VV/reader.cpp, VV/dataset.cpp:1347-1389, VV/dataset.cpp:1415-1467, VV/transferEdit.cpp:224-337) are exercised.

All volumes are [z][y][x] (x fastest), little-endian.
"""
import os
import struct
import zlib

import numpy as np

# --------------------------------------------------------------------------------------- vector fields


def _centres(n):
    """voxel centres of [-1,1]: x_i = -1 + (2i+1)/n"""
    return -1.0 + (2.0 * np.arange(n, dtype=np.float64) + 1.0) / n


def abc_flow(n):
    """ABC flow, A=sqrt3, B=sqrt2, C=1, arguments x pi (cfg1)."""
    c = _centres(n) * np.pi
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    A, B, C = np.sqrt(3.0), np.sqrt(2.0), 1.0
    v = np.stack([A * np.sin(z) + C * np.cos(y), B * np.sin(x) + A * np.cos(z), C * np.sin(y) + B * np.cos(x)], axis=-1)
    return np.ascontiguousarray(v, dtype=np.float32)


def rankine_vortex(n, core=0.25, vz=0.2):
    """Rankine vortex about z through the origin (cfg2)."""
    c = _centres(n)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    r = np.sqrt(x * x + y * y)
    with np.errstate(divide="ignore", invalid="ignore"):
        vt = np.where(r <= core, r / core, core / r)
        vx = np.where(r > 0, -y / r * vt, 0.0)
        vy = np.where(r > 0, x / r * vt, 0.0)
    v = np.stack([vx, vy, np.full_like(vx, vz)], axis=-1)
    return np.ascontiguousarray(v, dtype=np.float32)


def tornado(n, time=0.0):
    """Crawfis tornado, the gen_tornado procedure as written down in SURVEY.md Appendix C (cfg3)."""
    t = np.arange(n, dtype=np.float64) / (n - 1)
    z, y, x = np.meshgrid(t, t, t, indexing="ij")
    xc = 0.5 + 0.1 * np.sin(0.04 * time + 10.0 * z)
    yc = 0.5 + 0.1 * np.cos(0.03 * time + 3.0 * z)
    r = 0.1 + 0.4 * z * z + 0.1 * z * np.sin(8.0 * z)
    r2 = 0.2 + 0.1 * z
    tt = np.sqrt((x - xc) ** 2 + (y - yc) ** 2)
    s = np.abs(r - tt)
    s = np.where(s > r2, 0.8 - s, 1.0)
    z0 = np.maximum(0.0, 0.1 * (0.1 - tt * z))
    tp = np.sqrt(tt * tt + z0 * z0)
    s = (r + r2 - tp) * s / (tp + 1e-11) / (1.0 + z)
    v = np.stack([s * (y - yc) + 0.1 * (x - xc), -s * (x - xc) + 0.1 * (y - yc), s * z0], axis=-1)
    return np.ascontiguousarray(v, dtype=np.float32)


def _mt_uniform(seed, count):
    """`count` draws of std::mt19937(seed), as u = draw / 2^32 in [0,1)."""
    rs = np.random.RandomState(seed)
    raw = np.frombuffer(rs.bytes(4 * count), dtype="<u4")
    return raw.astype(np.float64) / 4294967296.0


def _lattice(seed, period=64):
    u = _mt_uniform(seed, period ** 3 * 3)
    return (2.0 * u - 1.0).reshape(period, period, period, 3)


def _sample_lattice(lat, px, py, pz):
    """periodic tri-linear interpolation of a [P][P][P][3] lattice at lattice coordinates (px,py,pz)"""
    P = lat.shape[0]
    fx, fy, fz = np.floor(px), np.floor(py), np.floor(pz)
    tx, ty, tz = (px - fx)[..., None], (py - fy)[..., None], (pz - fz)[..., None]
    x0, y0, z0 = fx.astype(np.int64) % P, fy.astype(np.int64) % P, fz.astype(np.int64) % P
    x1, y1, z1 = (x0 + 1) % P, (y0 + 1) % P, (z0 + 1) % P
    c00 = lat[z0, y0, x0] * (1 - tx) + lat[z0, y0, x1] * tx
    c10 = lat[z0, y1, x0] * (1 - tx) + lat[z0, y1, x1] * tx
    c01 = lat[z1, y0, x0] * (1 - tx) + lat[z1, y0, x1] * tx
    c11 = lat[z1, y1, x0] * (1 - tx) + lat[z1, y1, x1] * tx
    return (c00 * (1 - ty) + c10 * ty) * (1 - tz) + (c01 * (1 - ty) + c11 * ty) * tz


def curl_noise(n, seed, octaves=3, base_freq=4.0, slab=16):
    """curl of a 3-octave value-noise vector potential (cfg4 / cfg5): divergence-free turbulence.

    Psi(p) = sum_o 2^-o G_o(2^o * base_freq * p), G_o a periodic (64) tri-linear lattice of U[-1,1]^3 values from
    mt19937(seed + o); v = curl Psi by central differences with spacing 1/n in the [-1,1] coordinate.
    Evaluated slab by slab in z to bound memory."""
    if n >= 256:
        try:   # large fields: the same arithmetic (float64) on the GPU; 512^3 takes minutes in numpy
            import torch
            if torch.cuda.is_available():
                return _curl_noise_torch(n, seed, octaves, base_freq, slab)
        except ImportError:
            pass
    lats = [_lattice(seed + o) for o in range(octaves)]
    c = _centres(n)
    eps = 1.0 / n
    out = np.empty((n, n, n, 3), dtype=np.float32)

    def psi(x, y, z):
        acc = 0.0
        for o, lat in enumerate(lats):
            f = (2.0 ** o) * base_freq
            acc = acc + (2.0 ** -o) * _sample_lattice(lat, x * f, y * f, z * f)
        return acc

    for z0 in range(0, n, slab):
        zz, yy, xx = np.meshgrid(c[z0:z0 + slab], c, c, indexing="ij")
        dpx = (psi(xx + eps, yy, zz) - psi(xx - eps, yy, zz)) / (2 * eps)
        dpy = (psi(xx, yy + eps, zz) - psi(xx, yy - eps, zz)) / (2 * eps)
        dpz = (psi(xx, yy, zz + eps) - psi(xx, yy, zz - eps)) / (2 * eps)
        # curl = (dPz/dy - dPy/dz, dPx/dz - dPz/dx, dPy/dx - dPx/dy)
        out[z0:z0 + slab, ..., 0] = dpy[..., 2] - dpz[..., 1]
        out[z0:z0 + slab, ..., 1] = dpz[..., 0] - dpx[..., 2]
        out[z0:z0 + slab, ..., 2] = dpx[..., 1] - dpy[..., 0]
    return out


def _curl_noise_torch(n, seed, octaves, base_freq, slab, device="cuda"):
    """torch mirror of curl_noise (same float64 operations, evaluated on `device`); synthetic-input plumbing only"""
    import torch
    lats = [torch.from_numpy(_lattice(seed + o)).to(device) for o in range(octaves)]
    c = torch.from_numpy(_centres(n)).to(device)
    eps = 1.0 / n
    out = np.empty((n, n, n, 3), dtype=np.float32)

    def sample(lat, px, py, pz):
        P = lat.shape[0]
        fx, fy, fz = torch.floor(px), torch.floor(py), torch.floor(pz)
        tx, ty, tz = (px - fx)[..., None], (py - fy)[..., None], (pz - fz)[..., None]
        x0, y0, z0 = fx.long() % P, fy.long() % P, fz.long() % P
        x1, y1, z1 = (x0 + 1) % P, (y0 + 1) % P, (z0 + 1) % P
        c00 = lat[z0, y0, x0] * (1 - tx) + lat[z0, y0, x1] * tx
        c10 = lat[z0, y1, x0] * (1 - tx) + lat[z0, y1, x1] * tx
        c01 = lat[z1, y0, x0] * (1 - tx) + lat[z1, y0, x1] * tx
        c11 = lat[z1, y1, x0] * (1 - tx) + lat[z1, y1, x1] * tx
        return (c00 * (1 - ty) + c10 * ty) * (1 - tz) + (c01 * (1 - ty) + c11 * ty) * tz

    def psi(x, y, z):
        acc = 0.0
        for o, lat in enumerate(lats):
            f = (2.0 ** o) * base_freq
            acc = acc + (2.0 ** -o) * sample(lat, x * f, y * f, z * f)
        return acc

    for z0 in range(0, n, slab):
        zz, yy, xx = torch.meshgrid(c[z0:z0 + slab], c, c, indexing="ij")
        dpx = (psi(xx + eps, yy, zz) - psi(xx - eps, yy, zz)) / (2 * eps)
        dpy = (psi(xx, yy + eps, zz) - psi(xx, yy - eps, zz)) / (2 * eps)
        dpz = (psi(xx, yy, zz + eps) - psi(xx, yy, zz - eps)) / (2 * eps)
        v = torch.stack([dpy[..., 2] - dpz[..., 1], dpz[..., 0] - dpx[..., 2], dpx[..., 1] - dpy[..., 0]], dim=-1)
        out[z0:z0 + slab] = v.to(torch.float32).cpu().numpy()
    return out


def uniform_field(n, direction=(1.0, 0.0, 0.0)):
    """constant vector field: streamlines are straight lines (closed-form LIC tests)"""
    v = np.empty((n, n, n, 3), dtype=np.float32)
    v[...] = np.asarray(direction, dtype=np.float32)
    return v


# --------------------------------------------------------------------------------------- noise / scalar


def white_noise(n, seed, p):
    """n^3 u8 white noise: 255 with probability p else 0; mt19937(seed), one draw per voxel in file order.
    p = 1/6 is the reference's built-in 'sparse' noise (VV/dataset.cpp:1159-1162); p = 1/2 is 'dense'."""
    u = _mt_uniform(seed, n ** 3)
    return np.where(u < np.float64(np.float32(p)), 255, 0).astype(np.uint8).reshape(n, n, n)


SPARSE_P = 1.0 / 6.0
DENSE_P = 0.5


def constant_scalar(n=64, value=51):
    """scalar volume of constant 0.2 (= 51/255): passes the (0.1, 0.3) band of VV/shader/inc_lic.glsl:80"""
    return np.full((n, n, n), value, dtype=np.uint8)


# --------------------------------------------------------------------------------------- filter kernels


def filter_kernel(name, width=256):
    """one-row u8 LIC filter kernels: box / triangle / gaussian / cos2 (VV/screenshot-filter-*.png)"""
    i = np.arange(width, dtype=np.float64)
    c = (i + 0.5 - width / 2.0) / (width / 2.0)
    if name == "box":
        k = np.full(width, 255.0)
    elif name == "triangle":
        k = 255.0 * (1.0 - np.abs(c))
    elif name in ("gaussian", "gus"):
        k = 255.0 * np.exp(-c * c / (2.0 * (1.0 / 3.0) ** 2))
    elif name in ("cos2", "cos^2"):
        k = 255.0 * np.cos(np.pi * c / 2.0) ** 2
    else:
        raise ValueError("unknown filter kernel %r" % name)
    return np.floor(k + 0.5).astype(np.uint8)


# --------------------------------------------------------------------------------------- transfer functions


def default_tf():
    """TransferEdit ctor (VV/transferEdit.cpp:76-82): rgb = i, alpha = LIC opacity = max(0, i - 20)"""
    tf = np.zeros((256, 5), dtype=np.uint8)
    i = np.arange(256)
    tf[:, 0] = tf[:, 1] = tf[:, 2] = i
    tf[:, 3] = tf[:, 4] = np.where(i < 20, 0, i - 20)
    return tf


def tf_preset(name):
    """builder-defined presets named after VV/screenshot-tf-*.png / screenshot-op-*.png"""
    tf = default_tf()
    i = np.arange(256)
    t = i / 255.0
    if name == "default":
        return tf
    if name == "tf-length":
        r = np.where(t < 0.5, 2 * t, 1.0)
        g = np.where(t < 0.5, 2 * t, 2 * (1 - t))
        b = np.where(t < 0.5, 1.0, 2 * (1 - t))
        tf[:, 0] = np.floor(255 * r + 0.5)
        tf[:, 1] = np.floor(255 * g + 0.5)
        tf[:, 2] = np.floor(255 * b + 0.5)
        # semi-transparent (alpha <= 0.1): no ray sample can reach src.a > 0.95, so the shader's early termination
        # (Q4) never fires and every ray is marched over its full chord -- the ray-sample count of a frame is then the
        # analytic one (~21.7 M for cfg3 at the default view)
        tf[:, 3] = np.floor(255 * 0.1 * np.clip((i - 20) / 235.0, 0, 1) + 0.5)
        tf[:, 4] = i
        return tf
    if name == "op-high":
        tf[:, 4] = 230
        return tf
    if name == "op-low":
        tf[:, 4] = 40
        return tf
    if name == "op-random":
        tf[:, 4] = (_mt_uniform(7, 256) * 256).astype(np.uint8)
        return tf
    raise ValueError("unknown TF preset %r" % name)


# --------------------------------------------------------------------------------------- cameras


def quat_from_axis_angle(axis, angle_deg):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    h = np.deg2rad(angle_deg) / 2.0
    return tuple(np.float32(v) for v in (a[0] * np.sin(h), a[1] * np.sin(h), a[2] * np.sin(h), np.cos(h)))


CAMERA_DEFAULT = dict(quat=(0.0, 0.0, 0.0, 1.0), pos=(0.0, 0.0, 0.0), dist=4.0, fovy=35.0)       # VV/camera.cpp:42-47
CAMERA_CLOSE = dict(quat=quat_from_axis_angle((1, 1, 0), 35.0), pos=(0.0, 0.0, 0.0), dist=2.5, fovy=35.0)

# --------------------------------------------------------------------------------------- writers (reference formats)


def write_dat(path, volume, fmt=None, slice_thickness=(1, 1, 1), time_steps=None):
    """write <path>.dat + <path>.raw in the reference's format (VV/README.txt:25-32, VV/reader.cpp:115-185).
    volume: [z][y][x] (scalar) or [z][y][x][3]; time_steps: list of volumes -> '<name>_%d.raw' + TimeDependent."""
    vols = time_steps if time_steps is not None else [volume]
    v0 = vols[0]
    dim = 3 if v0.ndim == 4 else 1
    if fmt is None:
        fmt = {np.dtype(np.uint8): "UCHAR", np.dtype(np.float32): "FLOAT", np.dtype(np.uint16): "USHORT"}[v0.dtype]
    base = os.path.splitext(path)[0]
    name = os.path.basename(base)
    nz, ny, nx = v0.shape[:3]
    with open(base + ".dat", "w") as f:
        if time_steps is not None:
            f.write("ObjectFileName: %s_%%d.raw\n" % name)
            f.write("TimeDependent: 0 %d\n" % (len(vols) - 1))
        else:
            f.write("ObjectFileName: %s.raw\n" % name)
        f.write("Resolution:     %d %d %d\n" % (nx, ny, nz))
        f.write("SliceThickness: %g %g %g\n" % tuple(slice_thickness))
        f.write("Format:         %s%s\n" % (fmt, "3" if dim == 3 else ""))
    for t, v in enumerate(vols):
        raw = base + ("_%d.raw" % t if time_steps is not None else ".raw")
        np.ascontiguousarray(v).tofile(raw)
    return base + ".dat"


def write_noise(path, noise):
    """3 x int32 dims + u8 voxels (VV/dataset.cpp:1361-1377)"""
    nz, ny, nx = noise.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", nx, ny, nz))
        f.write(np.ascontiguousarray(noise, dtype=np.uint8).tobytes())
    return path


def write_png(path, img):
    """minimal 8-bit PNG writer: img [h][w] or [h][w][c], c in 1..4 (rows top to bottom)"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    return path


def write_tf(name, tf):
    """<name>_rgba.png (RGBA) + <name>_alpha.png (gray+alpha = alpha, LIC opacity), VV/README.txt:63-75"""
    base, ext = os.path.splitext(name)
    ext = ext or ".png"
    write_png(base + "_rgba" + ext, tf[None, :, :4])
    write_png(base + "_alpha" + ext, tf[None, :, 3:5])
    return name
