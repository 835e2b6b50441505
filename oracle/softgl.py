"""softgl.py -- TEST INFRASTRUCTURE ONLY: the geometry half of an OpenGL 2.1 pipeline in numpy (float64).

The reference leaves three things to the GL implementation between Renderer::drawCubeFaces / drawClippedPolygon
(VV/renderer.cpp:682-736, 1294-1309) and the fragment shader's gl_TexCoord[0]: clipping (view volume and the user planes
of glClipPlane), back-face culling and rasterisation with perspective-correct interpolation.  This module applies those
steps exactly as the OpenGL 2.1 specification states them to the primitives the reference's own code emits (captured by
oracle/ref_host_driver.cpp: vvref_raycast_draws), so the oracle's analytic ray entry points can be checked against the
fragments a conforming GL would hand to the reference's shader:

  eye = M v, clip = P eye                                                  (2.11)
  user planes: (p1' p2' p3' p4') . gl_ClipVertex >= 0, gl_ClipVertex = M v  (2.12; VV/shader/volic_vertex.glsl:12)
  view volume: -w <= x, y, z <= w; attributes interpolated linearly along clipped edges (2.12, 2.14.8)
  window = viewport(clip.xyz / clip.w)                                     (2.11.1)
  facing: sign of  a = 1/2 sum x_i y_i+1 - x_i+1 y_i  in window coordinates, counter-clockwise = front, back faces culled (2.14.1, 3.5.1)
  rasterisation: fragment centres (x + 1/2, y + 1/2) inside the polygon; attribute
                 f = sum(b_i f_i / w_i) / sum(b_i / w_i) with barycentric b_i  (3.5.1, eq. 3.6)

Not modelled (implementation-defined): sub-pixel snapping of window coordinates, fp32 interpolation, which of two
primitives sharing an edge owns a fragment centre lying exactly on it (here: the top-left rule).
"""
import numpy as np

GL_QUADS, GL_TRIANGLE_FAN, GL_POLYGON, GL_TRIANGLES = 7, 6, 9, 4


def _clip_polygon(verts, dist):
    """Sutherland-Hodgman against one half-space; verts [n][k] (any attributes, interpolated linearly), dist [n] signed
    distances (>= 0 kept)"""
    n = len(verts)
    out = []
    for i in range(n):
        a, b = verts[i], verts[(i + 1) % n]
        da, db = dist[i], dist[(i + 1) % n]
        if da >= 0:
            out.append(a)
        if (da >= 0) != (db >= 0):
            t = da / (da - db)
            out.append(a + t * (b - a))
    return np.asarray(out, np.float64).reshape(-1, verts.shape[1])


def _polygons(draw):
    v = np.concatenate([draw["verts"], draw["tex"]], axis=1)
    m = draw["mode"]
    if m == GL_QUADS:
        return [v[i:i + 4] for i in range(0, len(v) - 3, 4)]
    if m in (GL_TRIANGLE_FAN, GL_POLYGON):          # the reference's fans are planar convex polygons
        return [v] if len(v) >= 3 else []
    if m == GL_TRIANGLES:
        return [v[i:i + 3] for i in range(0, len(v) - 2, 3)]
    raise ValueError("primitive mode %d not handled" % m)


def _to_window(draw, poly):
    """object-space polygon [n][6] -> clipped polygon in window coordinates: (xw, yw, zw, 1/w_c, s, t, r) per vertex"""
    M, P = draw["modelview"], draw["projection"]
    pos = np.concatenate([poly[:, :3], np.ones((len(poly), 1))], axis=1)
    eye = pos @ M.T
    clip = eye @ P.T
    work = np.concatenate([clip, eye, poly[:, 3:6]], axis=1)          # [n][4 + 4 + 3]
    for i in range(6):
        if draw["clip_mask"] >> i & 1 and len(work) >= 3:
            work = _clip_polygon(work, work[:, 4:8] @ draw["clip_eye"][i])
    for axis in range(3):
        for sign in (1.0, -1.0):
            if len(work) >= 3:
                work = _clip_polygon(work, work[:, 3] - sign * work[:, axis])     # w - x >= 0 and w + x >= 0
    if len(work) < 3:
        return None
    w = work[:, 3]
    ndc = work[:, :3] / w[:, None]
    x0, y0, vw, vh = draw["viewport"]
    win = np.stack([vw / 2.0 * ndc[:, 0] + (x0 + vw / 2.0), vh / 2.0 * ndc[:, 1] + (y0 + vh / 2.0), 0.5 * ndc[:, 2] + 0.5], axis=1)
    return np.concatenate([win, (1.0 / w)[:, None], work[:, 8:11]], axis=1)


def _signed_area(win):
    x, y = win[:, 0], win[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _edge(ax, ay, bx, by, px, py):
    """edge function of a -> b at p (> 0 to the left); evaluated in a canonical endpoint order so that the two triangles
    sharing an edge see exactly opposite values (a real rasteriser gets this from fixed-point window coordinates)"""
    if (bx, by) < (ax, ay):
        return -((ax - bx) * (py - by) - (ay - by) * (px - bx))
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax)


def _raster_triangle(tri, width, height, emit, edge_dist=None):
    """tri [3][7] counter-clockwise in window coordinates; emit(ys, xs, attrs[n][3]); edge_dist [H][W] (optional) is lowered to
    the distance (in pixels) between each nearby fragment centre and the triangle's edge lines"""
    xs, ys = tri[:, 0], tri[:, 1]
    x_lo, x_hi = max(int(np.floor(xs.min() - 0.5)) - 1, 0), min(int(np.ceil(xs.max() - 0.5)) + 1, width - 1)
    y_lo, y_hi = max(int(np.floor(ys.min() - 0.5)) - 1, 0), min(int(np.ceil(ys.max() - 0.5)) + 1, height - 1)
    if x_lo > x_hi or y_lo > y_hi:
        return
    gy, gx = np.mgrid[y_lo:y_hi + 1, x_lo:x_hi + 1]
    px, py = gx + 0.5, gy + 0.5
    area2 = _edge(xs[0], ys[0], xs[1], ys[1], xs[2], ys[2])
    if not area2 > 0:
        return
    inside = np.ones(px.shape, bool)
    bary = []
    for i in range(3):
        a, b = (i + 1) % 3, (i + 2) % 3           # edge opposite to vertex i
        e = _edge(xs[a], ys[a], xs[b], ys[b], px, py)
        dx, dy = xs[b] - xs[a], ys[b] - ys[a]
        # top-left rule (y up, counter-clockwise): a left edge goes down (dy < 0), a top edge is horizontal going left (dx < 0)
        owns = (dy < 0) or (dy == 0 and dx < 0)
        inside &= (e > 0) | ((e == 0) & owns)
        bary.append(e / area2)
        if edge_dist is not None and (dx != 0 or dy != 0):
            sub = edge_dist[y_lo:y_hi + 1, x_lo:x_hi + 1]
            np.minimum(sub, np.abs(e) / np.hypot(dx, dy), out=sub)
    if not inside.any():
        return
    iw = tri[:, 3]
    den = sum(bary[i] * iw[i] for i in range(3))
    attrs = np.stack([sum(bary[i] * iw[i] * tri[i, 4 + k] for i in range(3)) / den for k in range(3)], axis=-1)
    emit(gy[inside], gx[inside], attrs[inside])


def _run(draws, width, height, program, emit, edge_dist, ordinal=None):
    """ordinal: optional one-element list that is kept at the number of selected primitives before the current one (the slice
    index of the slicing pass, which issues one glBegin / glEnd per slice whether or not the polygon is empty)"""
    facing = []
    nth = -1
    for di, d in enumerate(draws):
        if program is not None and d["program"] != program:
            continue
        nth += 1
        if ordinal is not None:
            ordinal[0] = nth
        for poly in _polygons(d):
            win = _to_window(d, poly)
            if win is None:
                continue
            a = _signed_area(win)
            culled = bool(d["cull"]) and not a > 0           # glCullFace(GL_BACK), glFrontFace(GL_CCW): the GL defaults
            facing.append((di, a, culled))
            if culled or a == 0:
                continue
            if a < 0:
                win = win[::-1]
            for k in range(1, len(win) - 1):                 # fan from vertex 0
                _raster_triangle(win[[0, k, k + 1]], width, height, emit, edge_dist)
    return facing


def rasterize(draws, width, height, program=None):
    """Fragments of the primitives drawn with `program` (None: all), in draw order.
    Returns (tex [H][W][4] float64 -- texcoord0 of the LAST fragment per pixel (no depth test, no blending: the ray-cast
             pass of VV/renderer.cpp:1093-1120), w = 1 where there is one;
             count [H][W] -- fragments per pixel;
             edge_dist [H][W] -- distance in pixels from the fragment centre to the nearest edge line of a rasterised
             triangle (inf if none nearby): coverage there is implementation-defined when it is ~0;
             facing -- list of (primitive index, signed window area, culled))"""
    tex = np.zeros((height, width, 4), np.float64)
    count = np.zeros((height, width), np.int32)
    edge_dist = np.full((height, width), np.inf)

    def emit(ys, xs, attrs):
        tex[ys, xs, :3] = attrs
        tex[ys, xs, 3] = 1.0
        np.add.at(count, (ys, xs), 1)

    facing = _run(draws, width, height, program, emit, edge_dist)
    return tex, count, edge_dist, facing


def fragment_lists(draws, width, height, program=None):
    """Every fragment of every pixel in draw order (the slicing pass, VV/renderer.cpp:1176-1225: slice i reads what slices
    0..i-1 left in the frame buffer).  Returns (starts int32 [H*W + 1], frags float64 [n][4], edge_dist [H][W]): the fragments
    of pixel p = y * W + x are frags[starts[p]:starts[p + 1]], columns = texcoord0.xyz and the index of the primitive among those
    drawn with `program` (= the slice index: which of the two ping-pong targets the fragment is written to)."""
    pix, att = [], []
    edge_dist = np.full((height, width), np.inf)
    ordinal = [0]

    def emit(ys, xs, attrs):
        pix.append(ys.astype(np.int64) * width + xs)
        att.append(np.concatenate([attrs, np.full((len(attrs), 1), float(ordinal[0]))], axis=1))

    _run(draws, width, height, program, emit, edge_dist, ordinal)
    starts = np.zeros(height * width + 1, np.int32)
    if not pix:
        return starts, np.zeros((0, 4)), edge_dist
    pix = np.concatenate(pix)
    att = np.concatenate(att, axis=0)
    order = np.argsort(pix, kind="stable")                  # emission order is kept inside a pixel
    starts[1:] = np.cumsum(np.bincount(pix, minlength=height * width))
    return starts, att[order], edge_dist
