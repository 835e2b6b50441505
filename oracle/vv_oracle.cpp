/*
 * vv_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see vv_oracle.h).
 *
 * fp32 restatement of the reference's GLSL hot path and of the host code that
 * defines its inputs.  Compiled with -ffp-contract=off so that every +,-,*,/
 * and sqrt is a single IEEE-754 binary32 operation, in the order the shader
 * source writes them.
 *
 * VV/ = /root/reference/VectorVisualization/
 */
#include "vv_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* GLSL normalize(): v / sqrt(dot(v,v)) */
inline V3 normalize(V3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
inline V3 rgb(V4 a) { return {a.x, a.y, a.z}; }

/* ---- fp16 round trip (RGBA16F upload, Q11 / Q14) ---------------------- */
inline float half_round(float x) { return (float)(_Float16)x; }

/* ---- GL samplers (GL 2.1 spec 3.8.8) ----------------------- */
enum Wrap { CLAMP_TO_EDGE, REPEAT, CLAMP_BORDER /* GL_CLAMP, border colour 0 */ };

struct Axis { int i0, i1; float f; bool b0, b1; /* b*: texel is border */ bool fused; /* weight_bits == -1, see lerp_axis */ };

inline float quant_weight(float f, int bits)
{
    if (bits <= 0) return f;
    float q = (float)(1 << bits);
    return std::floor(f * q + 0.5f) / q;
}

inline Axis axis_linear(float s, int n, Wrap wrap, int weight_bits)
{
    Axis a;
    a.b0 = a.b1 = false;
    if (wrap == REPEAT)
        s = s - std::floor(s);           /* REPEAT ignores the integer part of s */
    else if (wrap == CLAMP_BORDER)
        s = clampf(s, 0.0f, 1.0f);       /* GL_CLAMP clamps s to [0,1] */
    a.fused = (weight_bits == -1);
    /* weight_bits == -1: the same filtering rule with fused multiply-adds -- u = fma(s, n, -0.5) and a + f (b - a) as sub + fma (one
     * rounding fewer each) -- which is the "fp32 software trilinear interpolation" the CUDA path implements.  GL leaves the rounding of
     * the filter arithmetic to the implementation; the default (0) is the formula as the specification writes it. */
    float u = a.fused ? std::fmaf(s, (float)n, -0.5f) : s * (float)n - 0.5f;
    float fl = std::floor(u);
    a.f = quant_weight(u - fl, weight_bits);
    int i0 = (int)fl, i1 = i0 + 1;
    switch (wrap) {
    case CLAMP_TO_EDGE:
        i0 = std::min(std::max(i0, 0), n - 1);
        i1 = std::min(std::max(i1, 0), n - 1);
        break;
    case REPEAT:
        i0 = ((i0 % n) + n) % n;
        i1 = ((i1 % n) + n) % n;
        break;
    case CLAMP_BORDER:
        a.b0 = (i0 < 0 || i0 >= n);
        a.b1 = (i1 < 0 || i1 >= n);
        i0 = std::min(std::max(i0, 0), n - 1);
        i1 = std::min(std::max(i1, 0), n - 1);
        break;
    }
    a.i0 = i0; a.i1 = i1;
    return a;
}

inline float lerp_gl(float a, float b, float f) { return (1.0f - f) * a + f * b; }
inline float lerp_axis(float a, float b, const Axis &ax) { return ax.fused ? std::fmaf(ax.f, b - a, a) : lerp_gl(a, b, ax.f); }

/* generic trilinear fetch: C channels, texel decode by functor */
template <int C, class Fetch>
inline void trilinear(const int dim[3], float sx, float sy, float sz, Wrap wrap, int wb, Fetch fetch, float *out)
{
    Axis ax = axis_linear(sx, dim[0], wrap, wb);
    Axis ay = axis_linear(sy, dim[1], wrap, wb);
    Axis az = axis_linear(sz, dim[2], wrap, wb);
    float t[2][2][2][C];
    const int xs[2] = {ax.i0, ax.i1}, ys[2] = {ay.i0, ay.i1}, zs[2] = {az.i0, az.i1};
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i)
                fetch(xs[i], ys[j], zs[k], t[k][j][i]);
    for (int c = 0; c < C; ++c) {
        float x00 = lerp_axis(t[0][0][0][c], t[0][0][1][c], ax);
        float x10 = lerp_axis(t[0][1][0][c], t[0][1][1][c], ax);
        float x01 = lerp_axis(t[1][0][0][c], t[1][0][1][c], ax);
        float x11 = lerp_axis(t[1][1][0][c], t[1][1][1][c], ax);
        float y0 = lerp_axis(x00, x10, ay);
        float y1 = lerp_axis(x01, x11, ay);
        out[c] = lerp_axis(y0, y1, az);
    }
}

struct Ctx {
    const VVOScene *s;
    /* effective uniforms (renderer.cpp:925-996) */
    float stepSize;
    float gradient[3];
    float licParams[3];
    float licKernel[3];
    float alphaCorrection;
    int numIterations;
    V3 texMax, scaleVol, scaleVolInv;
    V3 camera;          /* gl_ModelViewMatrixInverse[3].xyz */
    float rot[9];       /* R (object -> eye), row-major */
    V3 lightPos;        /* gl_LightSource[0].position.xyz */
    float tanHalf, aspect;
};

/* volumeSampler: RGBA16F, LINEAR, CLAMP_TO_EDGE  (dataset.cpp:327-357) */
inline V4 texVolume(const Ctx &c, V3 p)
{
    const VVOScene *s = c.s;
    const float *T = s->vec;
    const int *d = s->vdim;
    float o[4];
    trilinear<4>(d, p.x, p.y, p.z, CLAMP_TO_EDGE, s->weight_bits < 0 ? -1 : 0,   /* (the 8-bit weight model is not applied to fp16 textures) */
                 [&](int x, int y, int z, float *t) {
                     const float *q = T + 4 * (((size_t)z * d[1] + y) * d[0] + x);
                     t[0] = q[0]; t[1] = q[1]; t[2] = q[2]; t[3] = q[3];
                 }, o);
    return {o[0], o[1], o[2], o[3]};
}

/* scalarSampler: LUMINANCE8, LINEAR, CLAMP_TO_EDGE (dataset.cpp:1025-1038) -> (L,L,L,1) */
inline V4 texScalar(const Ctx &c, V3 p)
{
    const VVOScene *s = c.s;
    if (!s->scalar) return {0.0f, 0.0f, 0.0f, 1.0f};
    const uint8_t *T = s->scalar;
    const int *d = s->sdim;
    float o[1];
    trilinear<1>(d, p.x, p.y, p.z, CLAMP_TO_EDGE, s->weight_bits,
                 [&](int x, int y, int z, float *t) {
                     t[0] = (float)T[((size_t)z * d[1] + y) * d[0] + x] / 255.0f;
                 }, o);
    return {o[0], o[0], o[0], 1.0f};
}

/* noiseSampler: LUMINANCE8 or RGBA8, LINEAR, REPEAT (dataset.cpp:1283-1335) */
inline V4 texNoise(const Ctx &c, V3 p)
{
    const VVOScene *s = c.s;
    const uint8_t *T = s->noise;
    const int *d = s->ndim;
    if (s->noise_channels == 4) {
        float o[4];
        trilinear<4>(d, p.x, p.y, p.z, REPEAT, s->weight_bits,
                     [&](int x, int y, int z, float *t) {
                         const uint8_t *q = T + 4 * (((size_t)z * d[1] + y) * d[0] + x);
                         for (int k = 0; k < 4; ++k) t[k] = (float)q[k] / 255.0f;
                     }, o);
        return {o[0], o[1], o[2], o[3]};
    }
    float o[1];
    trilinear<1>(d, p.x, p.y, p.z, REPEAT, s->weight_bits,
                 [&](int x, int y, int z, float *t) {
                     t[0] = (float)T[((size_t)z * d[1] + y) * d[0] + x] / 255.0f;
                 }, o);
    /* GL_LUMINANCE -> (L,L,L,1): .a is the constant 1 in the reference (Q7).  Default: .a = L. */
    return {o[0], o[0], o[0], s->quirk_luminance_alpha ? 1.0f : o[0]};
}

/* licVolumeSampler: RGBA16F, LINEAR, REPEAT (VolumeBuffer.cpp:47-55); only .r is used */
inline float texLicVol(const Ctx &c, V3 p)
{
    const VVOScene *s = c.s;
    const float *T = s->licvol;
    const int *d = s->ldim;
    float o[1];
    trilinear<1>(d, p.x, p.y, p.z, REPEAT, s->weight_bits < 0 ? -1 : 0,
                 [&](int x, int y, int z, float *t) { t[0] = T[((size_t)z * d[1] + y) * d[0] + x]; }, o);
    return o[0];
}

/* licKernelSampler: LUMINANCE8, LINEAR, GL_CLAMP (dataset.cpp:1489-1499) -> .r */
inline float texKernel(const Ctx &c, float x)
{
    const VVOScene *s = c.s;
    Axis a = axis_linear(x, s->kwidth, CLAMP_BORDER, s->weight_bits);
    float t0 = a.b0 ? 0.0f : (float)s->kernel[a.i0] / 255.0f;
    float t1 = a.b1 ? 0.0f : (float)s->kernel[a.i1] / 255.0f;
    return lerp_gl(t0, t1, a.f);
}

/* transferRGBASampler: RGBA8 256, LINEAR, CLAMP_TO_EDGE (transferEdit.cpp:505-521) */
inline V4 texTF(const Ctx &c, float x)
{
    const VVOScene *s = c.s;
    Axis a = axis_linear(x, 256, CLAMP_TO_EDGE, s->weight_bits);
    float o[4];
    for (int k = 0; k < 4; ++k)
        o[k] = lerp_gl((float)s->tf[5 * a.i0 + k] / 255.0f, (float)s->tf[5 * a.i1 + k] / 255.0f, a.f);
    return {o[0], o[1], o[2], o[3]};
}

/* transferAlphaOpacSampler: LUMINANCE_ALPHA8 (L = alpha channel 3, A = LIC opacity channel 4),
 * transferEdit.cpp:524-540 ; returns .a */
inline float texOpacA(const Ctx &c, float x)
{
    const VVOScene *s = c.s;
    Axis a = axis_linear(x, 256, CLAMP_TO_EDGE, s->weight_bits);
    return lerp_gl((float)s->tf[5 * a.i0 + 4] / 255.0f, (float)s->tf[5 * a.i1 + 4] / 255.0f, a.f);
}

/* 2D illumination tables: float, LINEAR, CLAMP_TO_EDGE (illumination.cpp createTex) */
inline void texIllum2D(const Ctx &c, int which, int ch, float sx, float sy, float *out)
{
    const VVOScene *s = c.s;
    const float *T = s->illum_tex[which];
    int w = s->illum_dim[0], h = s->illum_dim[1];
    Axis ax = axis_linear(sx, w, CLAMP_TO_EDGE, 0), ay = axis_linear(sy, h, CLAMP_TO_EDGE, 0);
    for (int k = 0; k < ch; ++k) {
        float a = lerp_gl(T[(ay.i0 * w + ax.i0) * ch + k], T[(ay.i0 * w + ax.i1) * ch + k], ax.f);
        float b = lerp_gl(T[(ay.i1 * w + ax.i0) * ch + k], T[(ay.i1 * w + ax.i1) * ch + k], ax.f);
        out[k] = lerp_gl(a, b, ay.f);
    }
}

/* ---- quaternion / view helpers (mmath.cpp:213-232, 150-200) -------------- */
inline void quat_angle_axis(const float q[4], float *angle, float axis[3])
{
    /* Quaternion_getAngleAxis, VV/mmath.cpp:213-232 */
    double d = std::sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2]);
    if (d > 1e-6) {
        d = 1.f / d;
        axis[0] = (float)(q[0] * d); axis[1] = (float)(q[1] * d); axis[2] = (float)(q[2] * d);
        if (1.0 - std::fabs(q[3]) > 1e-6) *angle = 2.f * (float)std::acos(q[3]);
        else *angle = 0.0f;
    } else {
        axis[0] = 0.f; axis[1] = 0.f; axis[2] = 1.f; *angle = 0.f;
    }
}

/* glRotatef(angle_deg, axis): GL 2.1 spec 2.11.2; evaluated in double, stored as float */
inline void rotation_matrix(float angle_rad, const float axis[3], float R[9])
{
    double deg = (double)(float)(angle_rad * 180.0 / M_PI);   /* camera.cpp:67 casts to float degrees */
    double a = deg * M_PI / 180.0;
    double x = axis[0], y = axis[1], z = axis[2];
    double n = std::sqrt(x * x + y * y + z * z);
    if (n > 0) { x /= n; y /= n; z /= n; }
    double c = std::cos(a), s = std::sin(a), t = 1.0 - c;
    double M[9] = {t * x * x + c,     t * x * y - s * z, t * x * z + s * y,
                   t * x * y + s * z, t * y * y + c,     t * y * z - s * x,
                   t * x * z - s * y, t * y * z + s * x, t * z * z + c};
    for (int i = 0; i < 9; ++i) R[i] = (float)M[i];
}

/* Quaternion_multVector3 (mmath.cpp): v' = q v q^-1 */
inline V3 quat_rotate(const float q[4], V3 v)
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    /* t = 2 * cross(q.xyz, v); v' = v + w*t + cross(q.xyz, t) */
    double tx = 2.0 * (y * v.z - z * v.y), ty = 2.0 * (z * v.x - x * v.z), tz = 2.0 * (x * v.y - y * v.x);
    return {(float)(v.x + w * tx + (y * tz - z * ty)),
            (float)(v.y + w * ty + (z * tx - x * tz)),
            (float)(v.z + w * tz + (x * ty - y * tx))};
}

/* raycast_program: the LIC ray-cast / slicing programs, where scaleVolInv can be an active uniform (Q1);
 * the LIC-volume and volume-ray-cast programs never reference it (VV/renderer.cpp:807-922) */
void make_ctx(const VVOScene *s, Ctx &c, bool raycast_program = true)
{
    c.s = s;
    /* renderer.cpp:947-995 */
    if (s->lowres) {
        c.stepSize = 2.0f * s->step_size_vol;
        c.gradient[0] = s->gradient_scale; c.gradient[1] = s->illum_scale; c.gradient[2] = 0.7f * s->freq_scale;
        c.licParams[0] = 15.0f; c.licParams[1] = 15.0f; c.licParams[2] = 1.0f / 64.0f;
        c.licKernel[0] = 0.5f / 15.0f; c.licKernel[1] = 0.5f / 15.0f;
        c.licKernel[2] = s->inv_filter_area / (30.0f);
        c.alphaCorrection = 2.0f * s->step_size_vol * 128.0f;
    } else {
        c.stepSize = s->step_size_vol;
        c.gradient[0] = s->gradient_scale; c.gradient[1] = s->illum_scale; c.gradient[2] = s->freq_scale;
        c.licParams[0] = (float)s->steps_fwd; c.licParams[1] = (float)s->steps_bwd; c.licParams[2] = s->step_size_lic;
        c.licKernel[0] = 0.5f / s->steps_fwd; c.licKernel[1] = 0.5f / s->steps_bwd;
        c.licKernel[2] = s->inv_filter_area / (s->steps_fwd + s->steps_bwd);
        c.alphaCorrection = s->step_size_vol * 128.0f;
    }
    c.numIterations = s->num_iterations;
    /* renderer.cpp:934-944, incl. Q1: scaleVolInv is written into scaleVol's slot in programs
     * where scaleVolInv is an active uniform (the ILLUM_* builds); scaleVolInv itself stays 0. */
    c.texMax = {s->extent[0] * s->scale[0], s->extent[1] * s->scale[1], s->extent[2] * s->scale[2]};
    bool inv_active = raycast_program && (s->illum_mode != VVO_ILLUM_NONE);
    if (inv_active && s->quirk_scalevolinv) {
        c.scaleVol = {s->scale_inv[0], s->scale_inv[1], s->scale_inv[2]};
        c.scaleVolInv = {0.0f, 0.0f, 0.0f};
    } else {
        c.scaleVol = {s->scale[0], s->scale[1], s->scale[2]};
        c.scaleVolInv = {s->scale_inv[0], s->scale_inv[1], s->scale_inv[2]};
    }
    /* view: camera.cpp:56-68 + renderer.cpp:146 */
    float angle, axis[3];
    quat_angle_axis(s->cam_quat, &angle, axis);
    rotation_matrix(angle, axis, c.rot);
    /* camera = center + R^T ((0,0,dist) - cam_pos) */
    double t[3] = {-(double)s->cam_pos[0], -(double)s->cam_pos[1], (double)s->cam_dist - (double)s->cam_pos[2]};
    double cam[3];
    for (int i = 0; i < 3; ++i)
        cam[i] = (double)s->center[i] + (double)c.rot[0 * 3 + i] * t[0] + (double)c.rot[1 * 3 + i] * t[1] + (double)c.rot[2 * 3 + i] * t[2];
    c.camera = {(float)cam[0], (float)cam[1], (float)cam[2]};
    c.tanHalf = (float)std::tan((double)s->fovy * M_PI / 360.0);
    c.aspect = (s->window_aspect > 0.0f) ? s->window_aspect : (float)s->width / (float)s->height;
    /* light: renderer.cpp:431-466: M = T(center) R(-angle, axis) T(cam_pos); lightPos = q_light (0,0,dist) */
    V3 lp = quat_rotate(s->light_quat, V3{0.0f, 0.0f, s->light_dist});
    float Rm[9];
    rotation_matrix(-angle, axis, Rm);
    double l[3] = {(double)lp.x + s->cam_pos[0], (double)lp.y + s->cam_pos[1], (double)lp.z + s->cam_pos[2]};
    c.lightPos = {(float)(Rm[0] * l[0] + Rm[1] * l[1] + Rm[2] * l[2] + s->center[0]),
                  (float)(Rm[3] * l[0] + Rm[4] * l[1] + Rm[5] * l[2] + s->center[1]),
                  (float)(Rm[6] * l[0] + Rm[7] * l[1] + Rm[8] * l[2] + s->center[2])};
}

/* Ray through the centre of pixel (x,y) (GL window coords, origin bottom-left).
 * The reference rasterises the cube's front faces with texcoord0 = vertex (renderer.cpp:682-736,
 * renderer.h:168-172): the fragment's gl_TexCoord[0] is the point where the pixel ray enters the box
 * [0,extent]^3.  Analytic slab test in double, rounded once to float; the coordinate of the entry
 * face is exact (it is constant over the rasterised quad). */
/* Plane j as the GL holds it from the second frame on: (n / |n|, d); |n| <= VS_EPS zeroes the normal (slicing.cpp:337-348) */
inline void clip_equation(const VVOScene *s, int j, double out[4])
{
    const double *e = s->clip_planes[j];
    const double len = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    for (int k = 0; k < 3; ++k) out[k] = (len > 1e-8) ? e[k] / len : 0.0;
    out[3] = e[3];
}

/* view-volume clipping of the proxy geometry (GL 2.1 spec 2.12): the pixel ray's parameter t is the eye-space depth of
 * o + t d because d = R^T (ex, ey, -1) */
inline bool in_depth_range(const VVOScene *s, double t) { return t >= (double)s->near_clip && t <= (double)s->far_clip; }

bool pixel_ray(const Ctx &c, int px, int py, V3 &entry)
{
    const VVOScene *s = c.s;
    double ex = (2.0 * (px + 0.5) / s->width - 1.0) * (double)c.tanHalf * (double)c.aspect;
    double ey = (2.0 * (py + 0.5) / s->height - 1.0) * (double)c.tanHalf;
    double ez = -1.0;
    /* object-space direction = R^T d_eye */
    double d[3];
    for (int i = 0; i < 3; ++i)
        d[i] = (double)c.rot[0 * 3 + i] * ex + (double)c.rot[1 * 3 + i] * ey + (double)c.rot[2 * 3 + i] * ez;
    double o[3] = {c.camera.x, c.camera.y, c.camera.z};
    double tn = -1e300, tf = 1e300;
    int face = -1; double faceval = 0.0;
    for (int i = 0; i < 3; ++i) {
        double lo = 0.0, hi = s->extent[i];
        if (d[i] == 0.0) {
            if (o[i] < lo || o[i] > hi) return false;
            continue;
        }
        double t0 = (lo - o[i]) / d[i], t1 = (hi - o[i]) / d[i];
        double fv = lo;
        if (t0 > t1) { std::swap(t0, t1); fv = hi; }
        if (t0 > tn) { tn = t0; face = i; faceval = fv; }
        if (t1 < tf) tf = t1;
    }
    /* front faces only (back faces culled, renderer.cpp:1099): camera outside the box; the part of a face nearer than the
     * near plane is clipped away, and the fragment with it */
    bool have = (tn < tf) && in_depth_range(s, tn) && face >= 0;
    double p[3] = {0.0, 0.0, 0.0};
    if (have) {
        for (int i = 0; i < 3; ++i) p[i] = o[i] + tn * d[i];
        p[face] = faceval;
        for (int i = 0; i < 3; ++i)
            p[i] = std::min(std::max(p[i], 0.0), (double)s->extent[i]);
    }
    const int nc = s->num_clip_planes;
    if (nc > 0) {
        /* User clip planes.  Renderer::render enables them around the whole draw (renderer.cpp:156-163, 199): GL discards
         * the part of the cube's front faces with n.q + d < 0 (q = position relative to the volume centre).
         * drawClippedPolygon (renderer.cpp:1294-1309) then draws, per active plane in order, the box's cross-section
         * n^.q = -(d - 0.0001) (ClipPlane / ViewSlicing::drawSingleSlice, slicing.cpp:368-560) with texcoord = position;
         * it is wound to face viewers looking along +n and back faces are culled; the other planes clip it too.  No depth
         * test and no blending in the FBO (renderer.cpp:1097-1101): the last fragment drawn under a pixel is the one the
         * shader's result is kept for. */
        double ctr[3] = {s->center[0], s->center[1], s->center[2]};
        double eq[3][4];
        for (int j = 0; j < nc; ++j) clip_equation(s, j, eq[j]);
        auto kept = [&](const double q[3], int skip) {
            for (int j = 0; j < nc; ++j) {
                if (j == skip) continue;
                const double *e = eq[j];
                if (e[0] * q[0] + e[1] * q[1] + e[2] * q[2] + e[3] < 0.0) return false;
            }
            return true;
        };
        if (have) {
            double q[3] = {p[0] - ctr[0], p[1] - ctr[1], p[2] - ctr[2]};
            have = kept(q, -1);
        }
        double oc[3] = {o[0] - ctr[0], o[1] - ctr[1], o[2] - ctr[2]};
        for (int i = 0; i < nc; ++i) {
            const double *e = s->clip_planes[i];
            double len = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
            if (!(len > 1e-8)) continue;                                   /* VS_EPS, slicing.h */
            double n[3] = {e[0] / len, e[1] / len, e[2] / len};
            double dist = -(e[3] - 0.0001);
            double dn = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
            if (!(dn > 0.0)) continue;                                     /* cap faces away from the viewer: culled */
            double t = (dist - (n[0] * oc[0] + n[1] * oc[1] + n[2] * oc[2])) / dn;
            if (!in_depth_range(s, t)) continue;
            double q[3] = {oc[0] + t * d[0], oc[1] + t * d[1], oc[2] + t * d[2]};
            bool inside = true;
            for (int k = 0; k < 3; ++k)
                if (std::fabs(q[k]) > 0.5 * (double)s->extent[k]) inside = false;
            if (!inside || !kept(q, i)) continue;
            have = true;
            for (int k = 0; k < 3; ++k)
                p[k] = std::min(std::max(q[k] + 0.5 * (double)s->extent[k], 0.0), (double)s->extent[k]);
        }
    }
    if (!have) return false;
    entry = {(float)p[0], (float)p[1], (float)p[2]};
    return true;
}

/* ---- view-aligned slicing geometry: ViewSlicing::setupSlicing / drawSlice, slicing.cpp:42-114 ------------- */
struct Slicing {
    float v[3];       /* unit view vector in object space (slicing.cpp:71-81) */
    float d;          /* depth range covered by the slices, 2 * max projection (slicing.cpp:87-99) */
    int numSlices;    /* (int)(d / sampDist) + 1 (slicing.cpp:101) */
};

/* Renderer::updateSlices (renderer.cpp:1270-1292): model-view of the frame, sampDist = stepSizeVol (x2 in low-res) */
Slicing setup_slicing(const Ctx &c)
{
    const VVOScene *s = c.s;
    Slicing sl;
    /* m[2], m[6], m[10] = third row of R; m[3] = m[7] = m[11] = 0; m[14] = z translation; m[15] = 1 */
    double tz = ((double)s->cam_pos[2] - (double)s->cam_dist) -
                ((double)c.rot[6] * s->center[0] + (double)c.rot[7] * s->center[1] + (double)c.rot[8] * s->center[2]);
    float m14 = (float)tz;
    float invNorm = 1.0f / (m14 - 1.0f);
    sl.v[0] = (c.rot[6] - 0.0f) * invNorm;
    sl.v[1] = (c.rot[7] - 0.0f) * invNorm;
    sl.v[2] = (c.rot[8] - 0.0f) * invNorm;
    invNorm = 1.0f / std::sqrt(sl.v[0] * sl.v[0] + sl.v[1] * sl.v[1] + sl.v[2] * sl.v[2]);
    sl.v[0] *= invNorm; sl.v[1] *= invNorm; sl.v[2] *= invNorm;
    float xMax = 0.5f * s->extent[0], yMax = 0.5f * s->extent[1], zMax = 0.5f * s->extent[2];
    float v[3] = {std::fabs(sl.v[0]), std::fabs(sl.v[1]), std::fabs(sl.v[2])};
    float dv[7] = {v[0] * xMax, v[1] * yMax, v[2] * zMax, v[0] * xMax + v[1] * yMax, v[0] * xMax + v[2] * zMax,
                   v[1] * yMax + v[2] * zMax, v[0] * xMax + v[1] * yMax + v[2] * zMax};
    sl.d = dv[6];
    for (int i = 0; i < 6; ++i)
        if (sl.d < dv[i]) sl.d = dv[i];
    sl.d *= 2.0;
    sl.numSlices = (int)(sl.d / c.stepSize) + 1;
    return sl;
}

/* plane offset of slice i from the volume centre along v (slicing.cpp:114) */
inline float slice_offset(const Slicing &sl, int slice) { return -0.5f * sl.d + (slice + 0.5f) * sl.d / sl.numSlices; }

struct PixelRay { double o[3], d[3]; };

inline PixelRay pixel_dir(const Ctx &c, int px, int py)
{
    const VVOScene *s = c.s;
    PixelRay r;
    double ex = (2.0 * (px + 0.5) / s->width - 1.0) * (double)c.tanHalf * (double)c.aspect;
    double ey = (2.0 * (py + 0.5) / s->height - 1.0) * (double)c.tanHalf;
    double ez = -1.0;
    for (int i = 0; i < 3; ++i)
        r.d[i] = (double)c.rot[0 * 3 + i] * ex + (double)c.rot[1 * 3 + i] * ey + (double)c.rot[2 * 3 + i] * ez;
    r.o[0] = c.camera.x; r.o[1] = c.camera.y; r.o[2] = c.camera.z;
    return r;
}

/* The fragment of slice i under a pixel: intersection of the pixel ray with the plane v.(p - center) = d_i, if it lies
 * in the box [0,extent]^3 (the slice polygon is that plane clipped to the box; texcoord0 = vertex, slicing.cpp:246-261). */
inline bool slice_fragment(const Ctx &c, const Slicing &sl, const PixelRay &r, int slice, V3 &geomPos)
{
    const VVOScene *s = c.s;
    double a = (r.o[0] - (double)s->center[0]) * (double)sl.v[0] + (r.o[1] - (double)s->center[1]) * (double)sl.v[1] +
               (r.o[2] - (double)s->center[2]) * (double)sl.v[2];
    double b = r.d[0] * (double)sl.v[0] + r.d[1] * (double)sl.v[1] + r.d[2] * (double)sl.v[2];
    if (b == 0.0) return false;
    double t = ((double)slice_offset(sl, slice) - a) / b;
    if (!in_depth_range(s, t)) return false;
    double p[3];
    for (int i = 0; i < 3; ++i) {
        p[i] = r.o[i] + t * r.d[i];
        if (p[i] < 0.0 || p[i] > (double)s->extent[i]) return false;
    }
    /* user clip planes are enabled around sliceVolume too (renderer.cpp:156-163): the slice polygons are clipped */
    for (int j = 0; j < s->num_clip_planes; ++j) {
        double e[4];
        clip_equation(s, j, e);
        double q[3] = {p[0] - (double)s->center[0], p[1] - (double)s->center[1], p[2] - (double)s->center[2]};
        if (e[0] * q[0] + e[1] * q[1] + e[2] * q[2] + e[3] < 0.0) return false;
    }
    geomPos = {(float)p[0], (float)p[1], (float)p[2]};
    return true;
}

/* ---- inc_lic.glsl -------------------------------------------------------- */

/* freqSamplingGrad, inc_lic.glsl:61-68: raw RGBA noise texel at pos (no frequency scale, no gate; Q8) */
inline V4 freqSamplingGrad(const Ctx &c, V3 pos) { return texNoise(c, pos); }

/* freqSampling, inc_lic.glsl:71-90 */
inline float freqSampling(const Ctx &c, V3 pos)
{
    if (c.s->noise_gate) {
        V4 scalarData = texScalar(c, pos);
        if (scalarData.x > 0.1f && scalarData.x < 0.3f)
            return texNoise(c, pos * c.gradient[2]).w;
        else
            return 0.0f;
    }
    return texNoise(c, pos * c.gradient[2]).w;
}

/* singleLICstep, inc_lic.glsl:93-145.  GRAD selects the USE_NOISE_GRADIENTS build. */
template <bool GRAD>
inline V4 singleLICstep(const Ctx &c, V3 licdir, V3 &newPos, V4 &step, float kernelOffset, float logEyeDist, float dir, float *dbg = nullptr)
{
    if (c.s->speed_of_flow) licdir = licdir * step.w;                    /* :108-110 */
    licdir = licdir * (c.licParams[2] * (logEyeDist * 0.5f + 0.3f));      /* :114 */
    V3 Pos2 = newPos + licdir;                                            /* :115 */
    V4 step2 = texVolume(c, Pos2);                                        /* :116 */
    if (dbg) { dbg[0] = Pos2.x; dbg[1] = Pos2.y; dbg[2] = Pos2.z; dbg[3] = step2.x; dbg[4] = step2.y; dbg[5] = step2.z; }   /* vvo_debug_walk */
    V3 licdir2 = 2.0f * rgb(step2) - V3{1.0f, 1.0f, 1.0f};               /* :117 */
    licdir2 = licdir2 * dir;                                              /* :118 */
    if (c.s->speed_of_flow) licdir2 = licdir2 * step.w;                   /* :120-122 */
    licdir2 = licdir2 * (c.licParams[2] * (logEyeDist * 0.5f + 0.3f));    /* :123 */
    newPos = newPos + 0.5f * (licdir + licdir2);                          /* :125 */
    step = texVolume(c, newPos);                                          /* :128 */
    V4 noise;
    if (GRAD) noise = freqSamplingGrad(c, newPos);                        /* :135 */
    else { float n = freqSampling(c, newPos); noise = {n, n, n, n}; }     /* :138 */
    noise = noise * texKernel(c, kernelOffset);                           /* :142 */
    return noise;
}

/* computeLIC, inc_lic.glsl:152-202 */
template <bool GRAD>
inline V4 computeLIC(const Ctx &c, V3 pos, V4 vectorFieldSample)
{
    const float logEyeDist = 0.0f;   /* Q3: declared, never written, read; defined as 0 here */
    float kernelOffset = 0.5f;
    V4 illum;
    if (GRAD) illum = freqSamplingGrad(c, pos);
    else { float n = freqSampling(c, pos); illum = {n, n, n, n}; }
    illum = illum * texKernel(c, 0.5f);                                   /* :172 */

    float dir = -1.0f;                                                     /* :174 */
    V3 newPos = pos;
    V4 step = vectorFieldSample;
    for (int i = 0; i < (int)c.licParams[1]; ++i) {                        /* :178 */
        V3 licdir = -2.0f * rgb(step) + V3{1.0f, 1.0f, 1.0f};             /* :180 */
        kernelOffset -= c.licKernel[1];                                    /* :182 */
        illum = illum + singleLICstep<GRAD>(c, licdir, newPos, step, kernelOffset, logEyeDist, dir);
    }
    dir = 1.0f;                                                            /* :188 */
    newPos = pos;
    step = vectorFieldSample;
    kernelOffset = 0.5f;
    for (int i = 0; i < (int)c.licParams[0]; ++i) {                        /* :192 */
        V3 licdir = 2.0f * rgb(step) - V3{1.0f, 1.0f, 1.0f};              /* :194 */
        kernelOffset += c.licKernel[0];                                    /* :196 */
        illum = illum + singleLICstep<GRAD>(c, licdir, newPos, step, kernelOffset, logEyeDist, dir);
    }
    return illum;                                                          /* :201 vec4(illum) */
}

/* ---- inc_illum.glsl ------------------------------------------------------ */

inline float opacity_correct(const Ctx &c, float a)
{
    /* color.a = 1.0 - pow(1.0 - color.a, alphaCorrection)   inc_illum.glsl:38,108,151,169 */
    return 1.0f - std::pow(1.0f - a, c.alphaCorrection);
}

/* illumLIC, inc_illum.glsl:158-172 */
inline V4 illumLIC(const Ctx &c, float illum, V4 tfData)
{
    V4 color;
    color.x = illum * tfData.x * c.gradient[1];
    color.y = illum * tfData.y * c.gradient[1];
    color.z = illum * tfData.z * c.gradient[1];
    color.w = texOpacA(c, illum * 1.3f) * tfData.w;
    color.w = opacity_correct(c, color.w);
    return color;
}

/* illumGradient, inc_illum.glsl:1-41.  gl_LightSource[0]: ambient 0, diffuse 1, specular 1 (GL defaults),
 * spotExponent = spec_exp (3DLIC.cpp:736). */
inline V4 illumGradient(const Ctx &c, V4 illum, V4 tfData, V3 pos, V3 dir)
{
    V3 lightDir = normalize(c.lightPos - pos * c.scaleVolInv);            /* :8 */
    V3 viewDir = normalize(-dir);                                          /* :9 */
    V3 normal = normalize(-rgb(illum));                                    /* :11 */
    V3 reflectDir = normalize(2.0f * dot(lightDir, normal) * normal - lightDir);   /* :13 */
    float spec = clampf(dot(reflectDir, viewDir), 0.0f, 1.0f);             /* :15 */
    spec = std::pow(spec, c.s->spec_exp);                                  /* :16 */
    V3 specular = V3{1.0f, 1.0f, 1.0f} * (spec * illum.w);                 /* :18 */
    /* mix(vec3(0), tf.rgb, illum.a) = 0*(1-a) + tf*a */
    V3 color = V3{0.0f * (1.0f - illum.w) + tfData.x * illum.w,
                  0.0f * (1.0f - illum.w) + tfData.y * illum.w,
                  0.0f * (1.0f - illum.w) + tfData.z * illum.w};           /* :22 */
    float diff = clampf(dot(lightDir, normal), 0.0f, 1.0f);                /* :24 */
    V3 diffuse = V3{1.0f, 1.0f, 1.0f} * diff * c.gradient[1];              /* :26 */
    color = color * (diffuse + V3{0.3f, 0.3f, 0.3f} + V3{0.0f, 0.0f, 0.0f}) + specular;   /* :29 */
    color = color * c.gradient[1];                                         /* :31 */
    V4 out = {color.x, color.y, color.z, 0.0f};
    out.w = texOpacA(c, illum.w * 1.3f) * tfData.w;                        /* :35 */
    out.w = opacity_correct(c, out.w);                                     /* :38 */
    return out;
}

/* illumMallo, inc_illum.glsl:46-112 */
inline V4 illumMallo(const Ctx &c, float illum, V4 tfData, V3 pos, V3 dir, V3 tangent)
{
    V3 lightDir = normalize(c.lightPos - pos * c.scaleVolInv);            /* :62 */
    V3 viewDir = normalize(-dir);
    tangent = normalize(2.0f * tangent - V3{1.0f, 1.0f, 1.0f});            /* :65 */
    V3 binormal = normalize(cross(tangent, viewDir));
    V3 normal = cross(binormal, tangent);
    V3 halfway = normalize(viewDir + lightDir);
    float lt[4] = {dot(lightDir, normal), dot(lightDir, tangent), dot(halfway, normal), dot(halfway, tangent)};
    float tmpx = 1.0f / std::sqrt(1.0f - lt[1] * lt[1]);                   /* :78 */
    float tmpy = 1.0f / std::sqrt(1.0f - lt[3] * lt[3]);                   /* :79 */
    float nz = lt[0] * tmpx, nw = lt[2] * tmpy;                            /* :80 lt.zw = lt.xz * tmp */
    lt[2] = nz; lt[3] = nw;
    for (int i = 0; i < 4; ++i) lt[i] = 0.5f * lt[i] + 0.5f;               /* :82 */
    float d1[1], s1[1];
    texIllum2D(c, 1, 1, lt[2], lt[1], d1);                                 /* :85  lt.zy */
    texIllum2D(c, 2, 1, lt[2], lt[3], s1);                                 /* :87  lt.zw */
    float w = std::pow(tmpy, -c.s->spec_exp);                              /* :91 */
    float specular = clampf(s1[0] * w, 0.0f, 1.0f);
    float diffuse = d1[0] * c.gradient[1];                                 /* :96 */
    V4 out;
    out.x = (0.0f * (1.0f - illum) + tfData.x * illum) * diffuse + specular;
    out.y = (0.0f * (1.0f - illum) + tfData.y * illum) * diffuse + specular;
    out.z = (0.0f * (1.0f - illum) + tfData.z * illum) * diffuse + specular;
    out.w = texOpacA(c, illum * 1.3f) * tfData.w;
    out.w = opacity_correct(c, out.w);
    return out;
}

/* illumZoeckler, inc_illum.glsl:117-154 (Q17: LUMINANCE_ALPHA read as .rg -> r = g = luminance) */
inline V4 illumZoeckler(const Ctx &c, float illum, V4 tfData, V3 pos, V3 dir, V3 tangent)
{
    V3 lightDir = normalize(c.lightPos - pos * c.scaleVolInv);            /* :124 */
    V3 viewDir = normalize(dir);                                           /* :125 */
    tangent = normalize(2.0f * tangent - V3{1.0f, 1.0f, 1.0f});
    float cx = 0.5f * dot(lightDir, tangent) + 0.5f;
    float cy = 0.5f * dot(viewDir, tangent) + 0.5f;
    float la[2];
    texIllum2D(c, 0, 2, cx, cy, la);
    float sr = la[0], sg = la[0];                                          /* .rg of (L,L,L,A) */
    sr *= c.gradient[1];
    V4 out;
    out.x = (0.0f * (1.0f - illum) + tfData.x * illum) * sr + 0.9f * sg;
    out.y = (0.0f * (1.0f - illum) + tfData.y * illum) * sr + 0.9f * sg;
    out.z = (0.0f * (1.0f - illum) + tfData.z * illum) * sr + 0.9f * sg;
    out.w = texOpacA(c, illum * 1.3f) * tfData.w;
    out.w = opacity_correct(c, out.w);
    return out;
}

inline float tf_index(const Ctx &c, V4 vectorData, V4 scalarData)
{
    switch (c.s->tf_mode) {
    case VVO_TF_A: return vectorData.w;
    case VVO_TF_R: return vectorData.x;
    case VVO_TF_LENGTH:   /* length(vectorData): the vec4 length, lic3d_fragment.glsl:55 */
        return std::sqrt(vectorData.x * vectorData.x + vectorData.y * vectorData.y + vectorData.z * vectorData.z + vectorData.w * vectorData.w);
    case VVO_TF_SCALAR: return scalarData.x;
    default: return vectorData.z;
    }
}

inline bool outside_box(const Ctx &c, V3 pos)
{
    /* any(bvec3(clamp(pos, 0, texMax) - pos))   lic3d_fragment.glsl:91 */
    return (clampf(pos.x, 0.0f, c.texMax.x) - pos.x) != 0.0f ||
           (clampf(pos.y, 0.0f, c.texMax.y) - pos.y) != 0.0f ||
           (clampf(pos.z, 0.0f, c.texMax.z) - pos.z) != 0.0f;
}

/* main() of lic3d_fragment.glsl:5-99 for one fragment */
template <bool GRAD>
V4 frag_raycast_lic(const Ctx &c, V3 geomPos, uint32_t &nsamples, float mc)
{
    bool outside = false;
    V3 pos = geomPos * c.scaleVol;                                         /* :14 */
    V3 geomDir = normalize(geomPos - c.camera);                            /* :20 */
    V3 dir = geomDir * c.scaleVol;                                         /* :21 */
    V4 dest = {0, 0, 0, 0}, src = {0, 0, 0, 0};
    nsamples = 0;
    if (mc >= 0.0f) pos = pos + dir * c.stepSize * mc;                     /* :31-33 USE_MC_OFFSET */
    for (int j = 0; !outside && j < c.numIterations; ++j) {                /* :38 */
        for (int i = 0; i < c.numIterations; ++i) {                        /* :40 */
            ++nsamples;
            V4 vectorData = texVolume(c, pos);                             /* :44 */
            V4 scalarData = texScalar(c, pos);                             /* :52 */
            V4 tfData = texTF(c, tf_index(c, vectorData, scalarData));     /* :54 */
            bool gate = (c.s->gate_mode == VVO_GATE_TF_ALPHA) ? (tfData.w > 0.05f)
                                                              : (scalarData.y > -0.0001f);   /* :59-61 */
            if (gate) {
                V4 illum = computeLIC<GRAD>(c, pos, vectorData);           /* :64 */
                illum.w *= c.licKernel[2] * c.gradient[0];                 /* :67 */
                switch (c.s->illum_mode) {                                 /* :71-80 */
                case VVO_ILLUM_GRADIENT: src = illumGradient(c, illum, tfData, pos, dir); break;
                case VVO_ILLUM_MALLO: src = illumMallo(c, illum.w, tfData, pos, dir, rgb(vectorData)); break;
                case VVO_ILLUM_ZOECKLER: src = illumZoeckler(c, illum.w, tfData, pos, dir, rgb(vectorData)); break;
                default: src = illumLIC(c, illum.w, tfData); break;
                }
                src.x *= src.w; src.y *= src.w; src.z *= src.w;            /* :83 */
                float k = 1.0f - dest.w;
                dest = {clampf(k * src.x + dest.x, 0.0f, 1.0f), clampf(k * src.y + dest.y, 0.0f, 1.0f),
                        clampf(k * src.z + dest.z, 0.0f, 1.0f), clampf(k * src.w + dest.w, 0.0f, 1.0f)};   /* :84 */
            }
            pos = pos + dir * c.stepSize;                                  /* :88 */
            outside = outside_box(c, pos) || (src.w > 0.95f);              /* :91 (Q4: src.a) */
            if (outside) break;
        }
    }
    return dest;
}

/* main() of raycast_lic3d_fragment.glsl:5-73 */
V4 frag_raycast_licvolume(const Ctx &c, V3 geomPos, uint32_t &nsamples)
{
    bool outside = false;
    V3 pos = geomPos * c.scaleVol;
    V3 geomDir = normalize(geomPos - c.camera);
    V3 dir = geomDir * c.scaleVol;
    V4 dest = {0, 0, 0, 0}, src;
    nsamples = 0;
    for (int j = 0; !outside && j < c.numIterations; ++j) {
        for (int i = 0; i < c.numIterations; ++i) {
            ++nsamples;
            V4 vectorData = texVolume(c, pos);                             /* :40 */
            float volumeData_r = texLicVol(c, pos);                        /* :41 */
            V4 tfData = texTF(c, vectorData.z);                            /* :47 */
            src = illumLIC(c, volumeData_r, tfData);                       /* :51 */
            src.x *= src.w; src.y *= src.w; src.z *= src.w;
            float k = 1.0f - dest.w;
            dest = {clampf(k * src.x + dest.x, 0.0f, 1.0f), clampf(k * src.y + dest.y, 0.0f, 1.0f),
                    clampf(k * src.z + dest.z, 0.0f, 1.0f), clampf(k * src.w + dest.w, 0.0f, 1.0f)};
            pos = pos + dir * c.stepSize;
            outside = outside_box(c, pos) || (dest.w > 0.95f);             /* :65 (dest.a here) */
            if (outside) break;
        }
    }
    return dest;
}

/* main() of lic3d_slicing_fragment.glsl:5-74 for one fragment; dest = the frame buffer under the fragment
 * (imageFBOSampler, :11).  The shader hard-codes TF index .a and gate tfData.a > 0.05; tf_mode / gate_mode select them here. */
template <bool GRAD>
V4 frag_slicing(const Ctx &c, V3 geomPos, V4 dest, bool &shaded, float mc)
{
    V4 src = {0, 0, 0, 0};
    shaded = false;
    if (dest.w < 0.95f) {                                                  /* :14 */
        shaded = true;
        V3 pos = geomPos * c.scaleVol;                                     /* :18 */
        V3 geomDir = normalize(geomPos - c.camera);                        /* :24 */
        V3 dir = geomDir * c.scaleVol;                                     /* :25 */
        if (mc >= 0.0f) pos = pos + dir * c.stepSize * mc;                 /* :31-33 USE_MC_OFFSET */
        V4 vectorData = texVolume(c, pos);                                 /* :36 */
        V4 scalarData = {0, 0, 0, 1};
        if (c.s->tf_mode == VVO_TF_SCALAR) scalarData = texScalar(c, pos);
        V4 tfData = texTF(c, tf_index(c, vectorData, scalarData));         /* :43 */
        bool gate = (c.s->gate_mode == VVO_GATE_TF_ALPHA) ? (tfData.w > 0.05f) : true;   /* :46 */
        if (gate) {
            V4 illum = computeLIC<GRAD>(c, pos, vectorData);               /* :49 */
            illum.w *= c.licKernel[2] * c.gradient[0];                     /* :52 */
            switch (c.s->illum_mode) {                                     /* :56-65 */
            case VVO_ILLUM_GRADIENT: src = illumGradient(c, illum, tfData, pos, dir); break;
            case VVO_ILLUM_MALLO: src = illumMallo(c, illum.w, tfData, pos, dir, rgb(vectorData)); break;
            case VVO_ILLUM_ZOECKLER: src = illumZoeckler(c, illum.w, tfData, pos, dir, rgb(vectorData)); break;
            default: src = illumLIC(c, illum.w, tfData); break;
            }
            src.x *= src.w; src.y *= src.w; src.z *= src.w;                /* :68 */
            float k = 1.0f - dest.w;
            dest = {clampf(k * src.x + dest.x, 0.0f, 1.0f), clampf(k * src.y + dest.y, 0.0f, 1.0f),
                    clampf(k * src.z + dest.z, 0.0f, 1.0f), clampf(k * src.w + dest.w, 0.0f, 1.0f)};   /* :69 */
        }
    }
    return dest;
}

template <class F>
uint64_t for_pixels(const VVOScene *s, int x0, int y0, int x1, int y1, float *out_rgba, uint32_t *out_samples, F frag, bool raycast_program = true)
{
    Ctx c;
    make_ctx(s, c, raycast_program);
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            V3 entry;
            V4 col = {0, 0, 0, 0};
            uint32_t n = 0;
            const float mc = s->mc_offsets ? s->mc_offsets[(size_t)y * s->width + x] : -1.0f;
            if (pixel_ray(c, x, y, entry)) col = frag(c, entry, n, mc);
            float *o = out_rgba + 4 * ((size_t)y * s->width + x);
            o[0] = col.x; o[1] = col.y; o[2] = col.z; o[3] = col.w;
            if (out_samples) out_samples[(size_t)y * s->width + x] = n;
            total += n;
        }
    }
    return total;
}

} // namespace

extern "C" {

uint64_t vvo_raycast_lic_rect(const VVOScene *s, int x0, int y0, int x1, int y1, float *out_rgba, uint32_t *out_samples)
{
    bool grad = (s->illum_mode == VVO_ILLUM_GRADIENT);   /* ILLUM_GRADIENT => USE_NOISE_GRADIENTS (inc_header.glsl:17-19) */
    if (grad)
        return for_pixels(s, x0, y0, x1, y1, out_rgba, out_samples,
                          [](const Ctx &c, V3 e, uint32_t &n, float mc) { return frag_raycast_lic<true>(c, e, n, mc); });
    return for_pixels(s, x0, y0, x1, y1, out_rgba, out_samples,
                      [](const Ctx &c, V3 e, uint32_t &n, float mc) { return frag_raycast_lic<false>(c, e, n, mc); });
}

uint64_t vvo_raycast_lic(const VVOScene *s, float *out_rgba, uint32_t *out_samples)
{
    return vvo_raycast_lic_rect(s, 0, 0, s->width, s->height, out_rgba, out_samples);
}

uint64_t vvo_raycast_licvolume(const VVOScene *s, float *out_rgba, uint32_t *out_samples)
{
    return for_pixels(s, 0, 0, s->width, s->height, out_rgba, out_samples,
                      [](const Ctx &c, V3 e, uint32_t &n, float) { return frag_raycast_licvolume(c, e, n); }, false);
}

/* Renderer::sliceVolume (renderer.cpp:1123-1267): numSlices view-aligned polygons drawn front to back, each fragment
 * runs lic3d_slicing_fragment.glsl on the frame buffer content left by the previous slices (FBO ping-pong variant).
 * out_samples counts the fragments that did any work (dest.a < 0.95). */
uint64_t vvo_slicing_lic(const VVOScene *s, float *out_rgba, uint32_t *out_samples)
{
    Ctx c;
    make_ctx(s, c);
    const Slicing sl = setup_slicing(c);
    const bool grad = (s->illum_mode == VVO_ILLUM_GRADIENT);
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int y = 0; y < s->height; ++y)
        for (int x = 0; x < s->width; ++x) {
            PixelRay r = pixel_dir(c, x, y);
            V4 dest = {0, 0, 0, 0};
            /* the two ping-pong targets, both cleared (renderer.cpp:1164-1165 clears the one attached before the loop, :1210-1211
             * the other one at i = 0); slice i is drawn into buf[(i + 1) & 1] and samples buf[i & 1] */
            V4 buf[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            uint32_t n = 0;
            const float mc = s->mc_offsets ? s->mc_offsets[(size_t)y * s->width + x] : -1.0f;
            for (int i = 0; i < sl.numSlices; ++i) {
                V3 g;
                if (!slice_fragment(c, sl, r, i, g)) continue;
                bool shaded;
                if (s->fbo_pingpong) dest = buf[i & 1];
                dest = grad ? frag_slicing<true>(c, g, dest, shaded, mc) : frag_slicing<false>(c, g, dest, shaded, mc);
                if (s->fbo_fp16) dest = {half_round(dest.x), half_round(dest.y), half_round(dest.z), half_round(dest.w)};
                if (s->fbo_pingpong) buf[(i + 1) & 1] = dest;
                if (shaded) ++n;
            }
            if (s->fbo_pingpong) dest = buf[sl.numSlices & 1];      /* what is displayed / stored: the target of slice N - 1 */
            float *o = out_rgba + 4 * ((size_t)y * s->width + x);
            o[0] = dest.x; o[1] = dest.y; o[2] = dest.z; o[3] = dest.w;
            if (out_samples) out_samples[(size_t)y * s->width + x] = n;
            total += n;
        }
    return total;
}

} /* extern "C" */

/* lic3d_slicingblend_fragment.glsl = lic3d_slicing_fragment.glsl without the frame-buffer read: with dest = 0 the FBO shader's
 * skip never fires and its blend clamp((1 - 0) src + 0) is the clamp the GL applies to a fragment colour before blending */
template <bool GRAD>
static V4 frag_slicing_blend(const Ctx &c, V3 g, bool &shaded, float mc)
{
    V4 src = frag_slicing<GRAD>(c, g, V4{0, 0, 0, 0}, shaded, mc);
    shaded = (src.x != 0.0f || src.y != 0.0f || src.z != 0.0f || src.w != 0.0f);
    return src;
}

static inline uint8_t unorm8(float v) { return (uint8_t)std::floor(clampf(v, 0.0f, 1.0f) * 255.0f + 0.5f); }

extern "C" {

int vvo_slice_fragment_colors(const VVOScene *s, int x, int y, float *out_rgba, int cap)
{
    Ctx c;
    make_ctx(s, c);
    const Slicing sl = setup_slicing(c);
    const bool grad = (s->illum_mode == VVO_ILLUM_GRADIENT);
    PixelRay r = pixel_dir(c, x, y);
    const float mc = s->mc_offsets ? s->mc_offsets[(size_t)y * s->width + x] : -1.0f;
    int n = 0;
    for (int i = 0; i < sl.numSlices; ++i) {
        V3 g;
        if (!slice_fragment(c, sl, r, i, g)) continue;
        if (n >= cap) return -1;
        bool shaded;
        V4 v = grad ? frag_slicing_blend<true>(c, g, shaded, mc) : frag_slicing_blend<false>(c, g, shaded, mc);
        out_rgba[4 * n] = v.x; out_rgba[4 * n + 1] = v.y; out_rgba[4 * n + 2] = v.z; out_rgba[4 * n + 3] = v.w;
        ++n;
    }
    return n;
}

uint64_t vvo_slicing_blend8(const VVOScene *s, uint8_t *out_rgba8, uint32_t *out_samples)
{
    Ctx c;
    make_ctx(s, c);
    const Slicing sl = setup_slicing(c);
    const bool grad = (s->illum_mode == VVO_ILLUM_GRADIENT);
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int y = 0; y < s->height; ++y)
        for (int x = 0; x < s->width; ++x) {
            PixelRay r = pixel_dir(c, x, y);
            uint8_t dst[4] = {0, 0, 0, 0};                                          /* glClear with (0, 0, 0, 0), renderer.cpp:1164-1165 */
            uint32_t n = 0;
            const float mc = s->mc_offsets ? s->mc_offsets[(size_t)y * s->width + x] : -1.0f;
            auto blend = [&](const float src[4]) {                                  /* glBlendFunc(ONE_MINUS_DST_ALPHA, ONE) */
                const float da = (float)dst[3] / 255.0f;
                for (int k = 0; k < 4; ++k) dst[k] = unorm8(src[k] * (1.0f - da) + (float)dst[k] / 255.0f);
            };
            for (int i = 0; i < sl.numSlices; ++i) {
                V3 g;
                if (!slice_fragment(c, sl, r, i, g)) continue;
                bool shaded;
                V4 v = grad ? frag_slicing_blend<true>(c, g, shaded, mc) : frag_slicing_blend<false>(c, g, shaded, mc);
                const float src[4] = {v.x, v.y, v.z, v.w};
                blend(src);
                if (shaded) ++n;
            }
            const float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};                        /* the screen-filling white plane, :1238-1255 */
            blend(white);
            uint8_t *o = out_rgba8 + 4 * ((size_t)y * s->width + x);
            for (int k = 0; k < 4; ++k) o[k] = dst[k];
            if (out_samples) out_samples[(size_t)y * s->width + x] = n;
            total += n;
        }
    return total;
}

/* slicing set-up read-back for tests: out[5] = v.xyz, d, numSlices */
void vvo_slicing_setup(const VVOScene *s, float *out)
{
    Ctx c;
    make_ctx(s, c);
    Slicing sl = setup_slicing(c);
    out[0] = sl.v[0]; out[1] = sl.v[1]; out[2] = sl.v[2]; out[3] = sl.d; out[4] = (float)sl.numSlices;
}

/* fragments of one pixel column for the reference-shader driver: out[numSlices][4] = (geomPos.xyz, valid) */
int vvo_slice_fragments(const VVOScene *s, int x, int y, float *out, int cap)
{
    Ctx c;
    make_ctx(s, c);
    Slicing sl = setup_slicing(c);
    PixelRay r = pixel_dir(c, x, y);
    int n = 0;
    for (int i = 0; i < sl.numSlices && n < cap; ++i) {
        V3 g;
        if (!slice_fragment(c, sl, r, i, g)) continue;
        out[4 * n] = g.x; out[4 * n + 1] = g.y; out[4 * n + 2] = g.z; out[4 * n + 3] = (float)i;   /* .w: the slice index */
        ++n;
    }
    return n;
}

/* lic3d_volume_fragment.glsl:2-21, one fragment per voxel centre of a w x h x d target
 * (renderer.cpp:1353-1358, VolumeBuffer.cpp:100-108).  The fragment's texcoord is interpolated over a
 * full-screen quad: ((x+.5)/w, (y+.5)/h, (z+.5)/d). */
void vvo_lic_volume(const VVOScene *s, int w, int h, int d, int z0, int z1, float *out)
{
    Ctx c;
    make_ctx(s, c, false);
    bool grad = (s->illum_mode == VVO_ILLUM_GRADIENT);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                V3 geomPos = {((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h, ((float)z + 0.5f) / (float)d};
                V3 pos = geomPos * c.scaleVol;                             /* :5 */
                V4 vectorData = texVolume(c, pos);                         /* :8 */
                V4 l = grad ? computeLIC<true>(c, pos, vectorData) : computeLIC<false>(c, pos, vectorData);
                float r = l.x;                                             /* :13 .r */
                r *= c.licKernel[2] * c.gradient[0];                       /* :15 */
                if (s->licvol_fp16) r = half_round(r);
                out[((size_t)z * h + y) * w + x] = r;
            }
}

void vvo_compute_lic(const VVOScene *s, const float pos[3], float out[4])
{
    Ctx c;
    make_ctx(s, c);
    V3 p = {pos[0], pos[1], pos[2]};
    V4 v = texVolume(c, p);
    V4 l = (s->illum_mode == VVO_ILLUM_GRADIENT) ? computeLIC<true>(c, p, v) : computeLIC<false>(c, p, v);
    out[0] = l.x; out[1] = l.y; out[2] = l.z; out[3] = l.w;
}

/* one direction of computeLIC's walk from pos, step by step: out[16 i ..] = (newPos.xyz, step.rgb, noise tap, kernel weight, Pos2.xyz, step2.rgb, 0, 0) */
void vvo_debug_walk(const VVOScene *s, const float pos[3], int dir_sign, int nsteps, float *out)
{
    Ctx c;
    make_ctx(s, c);
    V3 p = {pos[0], pos[1], pos[2]}, newPos = p;
    V4 step = texVolume(c, p);
    const float dir = dir_sign < 0 ? -1.0f : 1.0f;
    float kernelOffset = 0.5f;
    for (int i = 0; i < nsteps; ++i) {
        V3 licdir = dir < 0 ? -2.0f * rgb(step) + V3{1.0f, 1.0f, 1.0f} : 2.0f * rgb(step) - V3{1.0f, 1.0f, 1.0f};
        kernelOffset += dir < 0 ? -c.licKernel[1] : c.licKernel[0];
        V4 n = (s->illum_mode == VVO_ILLUM_GRADIENT) ? singleLICstep<true>(c, licdir, newPos, step, kernelOffset, 0.0f, dir, out + 16 * i + 8)
                                                     : singleLICstep<false>(c, licdir, newPos, step, kernelOffset, 0.0f, dir, out + 16 * i + 8);
        const float kw = texKernel(c, kernelOffset);
        out[16 * i] = newPos.x; out[16 * i + 1] = newPos.y; out[16 * i + 2] = newPos.z;
        out[16 * i + 3] = step.x; out[16 * i + 4] = step.y; out[16 * i + 5] = step.z;
        out[16 * i + 6] = (kw != 0.0f) ? n.w / kw : 0.0f; out[16 * i + 7] = kw;
        out[16 * i + 14] = out[16 * i + 15] = 0.0f;
    }
}

/* background_fragment.glsl:9-16 */
void vvo_background(const float *rgba, int n_pixels, float *out)
{
    for (int i = 0; i < n_pixels; ++i) {
        float a = rgba[4 * i + 3];
        for (int k = 0; k < 4; ++k) out[4 * i + k] = clampf((1.0f - a) * 1.0f + rgba[4 * i + k], 0.0f, 1.0f);
    }
}

void vvo_display_window(const float *rgba, int rw, int rh, int ww, int wh, float *out)
{
    const float vx = static_cast<float>(rw) / ww, vy = static_cast<float>(rh) / wh;
    for (int y = 0; y < wh; ++y)
        for (int x = 0; x < ww; ++x) {
            int sx = (int)std::floor(((float)x + 0.5f) * vx), sy = (int)std::floor(((float)y + 0.5f) * vy);
            sx = sx < 0 ? 0 : (sx > rw - 1 ? rw - 1 : sx);
            sy = sy < 0 ? 0 : (sy > rh - 1 ? rh - 1 : sy);
            vvo_background(rgba + 4 * ((size_t)sy * rw + sx), 1, out + 4 * ((size_t)y * ww + x));
        }
}

/* GL float -> UNORM8 conversion of the stored frame (renderer.cpp:216-226; GL 2.1 spec 2.14.9) */
void vvo_quantize_rgba8(const float *rgba, int n, uint8_t *out)
{
    for (int i = 0; i < n; ++i) out[i] = (uint8_t)std::floor(clampf(rgba[i], 0.0f, 1.0f) * 255.0f + 0.5f);
}

void vvo_sample_vec(const VVOScene *s, const float p[3], float out[4])
{
    Ctx c; c.s = s;
    V4 v = texVolume(c, V3{p[0], p[1], p[2]});
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
void vvo_sample_noise(const VVOScene *s, const float p[3], float out[4])
{
    Ctx c; c.s = s;
    V4 v = texNoise(c, V3{p[0], p[1], p[2]});
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
void vvo_sample_scalar(const VVOScene *s, const float p[3], float out[4])
{
    Ctx c; c.s = s;
    V4 v = texScalar(c, V3{p[0], p[1], p[2]});
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
float vvo_sample_kernel(const VVOScene *s, float x) { Ctx c; c.s = s; return texKernel(c, x); }
void vvo_sample_tf(const VVOScene *s, float x, float out_rgba[4], float out_la[2])
{
    Ctx c; c.s = s;
    V4 v = texTF(c, x);
    out_rgba[0] = v.x; out_rgba[1] = v.y; out_rgba[2] = v.z; out_rgba[3] = v.w;
    Axis a = axis_linear(x, 256, CLAMP_TO_EDGE, s->weight_bits);
    out_la[0] = lerp_gl((float)s->tf[5 * a.i0 + 3] / 255.0f, (float)s->tf[5 * a.i1 + 3] / 255.0f, a.f);
    out_la[1] = texOpacA(c, x);
}

void vvo_derive_uniforms(const VVOScene *s, float *o)
{
    Ctx c;
    make_ctx(s, c);
    o[0] = c.stepSize;
    o[1] = c.gradient[0]; o[2] = c.gradient[1]; o[3] = c.gradient[2];
    o[4] = c.licParams[0]; o[5] = c.licParams[1]; o[6] = c.licParams[2];
    o[7] = c.licKernel[0]; o[8] = c.licKernel[1]; o[9] = c.licKernel[2];
    o[10] = c.alphaCorrection;
    o[11] = (float)c.numIterations;
    o[12] = c.licParams[2] * (0.0f * 0.5f + 0.3f);
    o[13] = o[14] = o[15] = 0.0f;
}

void vvo_view(const VVOScene *s, float cam_obj[3], float rot[9])
{
    Ctx c;
    make_ctx(s, c);
    cam_obj[0] = c.camera.x; cam_obj[1] = c.camera.y; cam_obj[2] = c.camera.z;
    std::memcpy(rot, c.rot, sizeof(c.rot));
}

void vvo_light_position(const VVOScene *s, float out[4])
{
    Ctx c;
    make_ctx(s, c);
    out[0] = c.lightPos.x; out[1] = c.lightPos.y; out[2] = c.lightPos.z; out[3] = 1.0f;
}

int vvo_pixel_ray(const VVOScene *s, int x, int y, float entry[3], float dir[3])
{
    Ctx c;
    make_ctx(s, c);
    V3 e;
    if (!pixel_ray(c, x, y, e)) return 0;
    V3 d = normalize(e - c.camera);
    entry[0] = e.x; entry[1] = e.y; entry[2] = e.z;
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    return 1;
}

/* entry points of all pixel rays of [x0,x1) x [y0,y1): out[(y-y0)*(x1-x0) + (x-x0)][4] = (entry.xyz, hit ? 1 : 0) */
void vvo_pixel_rays(const VVOScene *s, int x0, int y0, int x1, int y1, float *out)
{
    Ctx c;
    make_ctx(s, c);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            float *o = out + 4 * ((size_t)(y - y0) * (x1 - x0) + (x - x0));
            V3 e;
            if (pixel_ray(c, x, y, e)) { o[0] = e.x; o[1] = e.y; o[2] = e.z; o[3] = 1.0f; }
            else { o[0] = o[1] = o[2] = o[3] = 0.0f; }
        }
}

/* effective scaleVol / scaleVolInv / texMax uniforms of the ray-cast program (incl. Q1): out[9] */
void vvo_scale_uniforms(const VVOScene *s, float *out)
{
    Ctx c;
    make_ctx(s, c);
    out[0] = c.scaleVol.x; out[1] = c.scaleVol.y; out[2] = c.scaleVol.z;
    out[3] = c.scaleVolInv.x; out[4] = c.scaleVolInv.y; out[5] = c.scaleVolInv.z;
    out[6] = c.texMax.x; out[7] = c.texMax.y; out[8] = c.texMax.z;
}

/* VectorDataSet::loadData, dataset.cpp:144-176 */
void vvo_volume_geometry(const int size[3], const float slice_dist[3], float extent[3], float scale[3], float scale_inv[3], float center[3])
{
    float volSize[3], maxVolSize = 0;
    int maxTexSize = 0;
    for (int i = 0; i < 3; ++i) {
        volSize[i] = size[i] * slice_dist[i];
        if (volSize[i] > maxVolSize) maxVolSize = volSize[i];
        if (size[i] > maxTexSize) maxTexSize = size[i];
    }
    for (int i = 0; i < 3; ++i) {
        scale[i] = maxTexSize / (size[i] * slice_dist[i]);
        scale_inv[i] = size[i] * slice_dist[i] / maxTexSize;
        extent[i] = size[i] * slice_dist[i] / maxVolSize;
        center[i] = extent[i] / 2.0f;
    }
}

/* VectorDataSet::fillTexDataFloatInterp, dataset.cpp:533-635 (FLOAT branch :586-611, normalise :623-631) */
void vvo_pack_vector_field(const float *v0, const float *v1, const int dim[3], int interp_index, int interp_size, int fp16, float *out)
{
    size_t n = (size_t)dim[0] * dim[1] * dim[2];
    float maxLen = -1.0f;
    for (size_t a = 0; a < n; ++a) {
        float t[3];
        for (int k = 0; k < 3; ++k) {
            float a0 = v0[3 * a + k];
            float a1 = v1 ? v1[3 * a + k] : a0;
            t[k] = a0 + (float)interp_index / interp_size * (a1 - a0);    /* :590 */
        }
        float len = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);    /* :592 (SQR on float) */
        if (len < 1e-5f) {                                                 /* EPS, mmath.h:43 */
            len = 0.0f;
            out[4 * a] = out[4 * a + 1] = out[4 * a + 2] = 0.5f;
        } else {
            out[4 * a] = 0.5f * t[0] / len + 0.5f;                         /* :605-607 */
            out[4 * a + 1] = 0.5f * t[1] / len + 0.5f;
            out[4 * a + 2] = 0.5f * t[2] / len + 0.5f;
        }
        if (len > maxLen) maxLen = len;
        out[4 * a + 3] = len;
    }
    for (size_t a = 0; a < n; ++a) {
        float len = out[4 * a + 3] / maxLen;                               /* :629 */
        out[4 * a + 3] = (len > 1.0f) ? 1.0f : ((len < 0.0f) ? 0.0f : len);
    }
    if (fp16)
        for (size_t i = 0; i < 4 * n; ++i) out[i] = half_round(out[i]);    /* GL_RGBA16F_ARB upload, dataset.cpp:329-347 */
}

/* VectorDataSet::fillTexDataFloat UCHAR branch, dataset.cpp:466-490: centred at 128, no magnitude normalisation
 * (Q12/Q20: the non-mutating formula) */
void vvo_pack_vector_field_u8(const uint8_t *v0, const int dim[3], int fp16, float *out)
{
    size_t n = (size_t)dim[0] * dim[1] * dim[2];
    float maxLen = -1.0f;
    for (size_t a = 0; a < n; ++a) {
        float t[3];
        for (int k = 0; k < 3; ++k) t[k] = (float)v0[3 * a + k] - 128.0f;
        float len = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        if (len < 1e-5f) {
            len = 0.0f;
            out[4 * a] = out[4 * a + 1] = out[4 * a + 2] = 0.5f;
        } else {
            out[4 * a] = 0.5f * t[0] / len + 0.5f;
            out[4 * a + 1] = 0.5f * t[1] / len + 0.5f;
            out[4 * a + 2] = 0.5f * t[2] / len + 0.5f;
        }
        if (len > maxLen) maxLen = len;
        out[4 * a + 3] = len;
    }
    for (size_t a = 0; a < n; ++a) {
        float len = out[4 * a + 3] / maxLen;
        out[4 * a + 3] = (len > 1.0f) ? 1.0f : ((len < 0.0f) ? 0.0f : len);
    }
    if (fp16)
        for (size_t i = 0; i < 4 * n; ++i) out[i] = half_round(out[i]);
}

/* computeGradients, gradient.cpp:190-374 (SOBEL == 1 build, gradient.h:37) */
void vvo_compute_gradients_f(const uint8_t *vol, const int dim[3], const float sd[3], float *g)
{
    static const int W[3][3][3][3] = {
        {{{-1, -3, -1}, {-3, -6, -3}, {-1, -3, -1}}, {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, {{1, 3, 1}, {3, 6, 3}, {1, 3, 1}}},
        {{{-1, -3, -1}, {0, 0, 0}, {1, 3, 1}}, {{-3, -6, -3}, {0, 0, 0}, {3, 6, 3}}, {{-1, -3, -1}, {0, 0, 0}, {1, 3, 1}}},
        {{{-1, 0, 1}, {-3, 0, 3}, {-1, 0, 1}}, {{-3, 0, 3}, {-6, 0, 6}, {-3, 0, 3}}, {{-1, 0, 1}, {-3, 0, 3}, {-1, 0, 1}}}};
    const int nx = dim[0], ny = dim[1], nz = dim[2];
    auto vox = [&](int x, int y, int z) { return (float)vol[((size_t)z * ny + y) * nx + x]; };
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                float *gp = g + 3 * (((size_t)z * ny + y) * nx + x);
                if (x > 0 && x < nx - 1 && y > 0 && y < ny - 1 && z > 0 && z < nz - 1) {
                    for (int dir = 0; dir < 3; ++dir) {
                        gp[dir] = 0.0f;
                        /* NB: weights[dir][i+1][j+1][k+1] is applied to voxel (x+i, y+j, z+k): :252-262 */
                        for (int i = -1; i < 2; ++i)
                            for (int j = -1; j < 2; ++j)
                                for (int k = -1; k < 2; ++k)
                                    gp[dir] += W[dir][i + 1][j + 1][k + 1] * vox(x + i, y + j, z + k);
                        gp[dir] /= 2.0f * sd[dir];
                    }
                } else {
                    gp[0] = (x < 1) ? (vox(x + 1, y, z) - vox(x, y, z)) / sd[0] : (vox(x, y, z) - vox(x - 1, y, z)) / sd[0];
                    gp[1] = (y < 1) ? (vox(x, y + 1, z) - vox(x, y, z)) / sd[1] : (vox(x, y, z) - vox(x, y - 1, z)) / sd[1];
                    gp[2] = (z < 1) ? (vox(x, y, z + 1) - vox(x, y, z)) / sd[2] : (vox(x, y, z) - vox(x, y, z - 1)) / sd[2];
                }
            }
}

/* filterGradients, gradient.cpp:377-459, including Q16 (loops k = -fw .. fw-2; border fw shrinks but the
 * kernel built for fw = 2 is indexed with the shrunken fw) */
void vvo_filter_gradients_f(const int dim[3], float *grad)
{
    const int fSize = 5;
    int fw0 = fSize / 2;
    const int nx = dim[0], ny = dim[1], nz = dim[2];
    std::vector<float> filter(fSize * fSize * fSize, 0.0f);
    float sum = 0.0f;
    for (int k = -fw0; k < fw0 - 1; ++k)
        for (int j = -fw0; j < fw0 - 1; ++j)
            for (int i = -fw0; i < fw0 - 1; ++i)
                sum += filter[((fw0 + k) * fSize + fw0 + j) * fSize + fw0 + i] = std::exp(-(float)(i * i + j * j + k * k) / 5.0f);
    for (int k = -fw0; k < fw0 - 1; ++k)
        for (int j = -fw0; j < fw0 - 1; ++j)
            for (int i = -fw0; i < fw0 - 1; ++i)
                filter[((fw0 + k) * fSize + fw0 + j) * fSize + fw0 + i] /= sum;
    std::vector<float> outv((size_t)3 * nx * ny * nz);
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                int bx = std::min(x, nx - x - 1), by = std::min(y, ny - y - 1), bz = std::min(z, nz - z - 1);
                int fw = std::min(fSize / 2, std::min(std::min(bx, by), bz));
                size_t gi = 3 * (((size_t)z * ny + y) * nx + x);
                for (int n = 0; n < 3; ++n) {
                    float acc = 0.0f;
                    for (int k = -fw; k < fw - 1; ++k)
                        for (int j = -fw; j < fw - 1; ++j)
                            for (int i = -fw; i < fw - 1; ++i) {
                                size_t ogi = 3 * (((size_t)(z + k) * ny + (y + j)) * nx + (x + i)) + n;
                                acc += filter[((fw + k) * fSize + fw + j) * fSize + fw + i] * grad[ogi];
                            }
                    outv[gi + n] = acc;
                }
            }
    std::memcpy(grad, outv.data(), outv.size() * sizeof(float));
}

/* quantize8, gradient.cpp:462-478 */
static void quantize8(float *g, uint8_t *d)
{
    float len = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (len < 1e-5f) g[0] = g[1] = g[2] = 0.0f;
    else { g[0] /= len; g[1] /= len; g[2] /= len; }
    for (int i = 0; i < 3; ++i) d[i] = (unsigned char)((g[i] + 1.0) / 2.0 * 255);
}

void vvo_noise_gradients(const uint8_t *noise, const int dim[3], const float sd[3], uint8_t *out)
{
    size_t n = (size_t)dim[0] * dim[1] * dim[2];
    std::vector<float> g(3 * n);
    vvo_compute_gradients_f(noise, dim, sd, g.data());
    vvo_filter_gradients_f(dim, g.data());
    for (size_t i = 0; i < n; ++i) quantize8(&g[3 * i], &out[3 * i]);
}

/* NoiseDataSet::createTexture gradient branch, dataset.cpp:1264-1282 */
void vvo_pack_noise_rgba(const uint8_t *noise, const uint8_t *grad, int n, uint8_t *out)
{
    for (int i = 0; i < n; ++i) {
        out[4 * i] = grad[3 * i]; out[4 * i + 1] = grad[3 * i + 1]; out[4 * i + 2] = grad[3 * i + 2];
        out[4 * i + 3] = noise[i];
    }
}

/* White noise: value 255 with probability p, else 0.  The reference draws
 * floor(0.6 rand()/(RAND_MAX+1) + 0.5)*255 from an unseeded rand() (dataset.cpp:1159-1162, P(255) = 1/6);
 * the synthetic configs use mt19937(seed), one draw per voxel in file order, u = draw / 2^32. */
void vvo_white_noise(int n, uint32_t seed, float p, uint8_t *out)
{
    std::mt19937 gen(seed);
    for (int i = 0; i < n; ++i) {
        double u = (double)gen() / 4294967296.0;
        out[i] = (u < (double)p) ? 255 : 0;
    }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

/* LICFilter::loadData + calcFilterKernelInvArea, dataset.cpp:1415-1467, 1503-1512 */
int vvo_filter_from_row(const uint8_t *row, int width, int channels, uint8_t *out, float *inv_area)
{
    int fw = next_pow2(width);
    int shift = (fw - width) / 2;
    std::memset(out, 0, fw);
    for (int i = 0; i < width; ++i) out[i + shift] = row[i * channels];
    float area = 0.0f;
    for (int i = 0; i < fw; ++i) area += out[i];
    *inv_area = 0.5f * fw * 255.0f / area;
    return fw;
}

/* LICFilter::createBoxFilter, dataset.cpp:1405-1413 */
int vvo_box_filter(int width, uint8_t *out, float *inv_area)
{
    int fw = next_pow2(width);
    std::memset(out, 255, fw);
    *inv_area = 0.5f;
    return fw;
}

/* TransferEdit::TransferEdit, transferEdit.cpp:76-82 */
void vvo_default_tf(uint8_t *tf)
{
    for (int i = 0; i < 256; ++i) {
        for (int ch = 0; ch < 3; ++ch) tf[5 * i + ch] = (unsigned char)i;
        for (int ch = 3; ch < 5; ++ch) tf[5 * i + ch] = (unsigned char)((i < 20) ? 0 : i - 20);
    }
}

/* GL float -> UNORM8 -> float of an 8-bit texture upload */
static float unorm8_roundtrip(float v) { return std::floor(clampf(v, 0.0f, 1.0f) * 255.0f + 0.5f) / 255.0f; }

/* Illumination::createIllumTexZoeckler / createIllumTexMallo / computeSpecTermMallo, illumination.cpp:96-390, with the
 * LineMat defaults of illumination.h:40-62.  createIllumTextures does not forward floatTex (illumination.cpp:59-91), so
 * the textures are 8-bit: out arrays hold the decoded UNORM8 values.  zoeckler [h][w][2], mallo_* [h][w] (4 equal channels). */
void vvo_illum_tables(float spec_exp, int w, int h, float *zoeckler, float *mallo_diff, float *mallo_spec)
{
    const float lightColor = 1.0f, ambient = 0.1f, diffuseM = 0.5f, specularM = 0.8f, diffExp = 2.0f;
    auto integrand = [](double beta, double n, double theta) {
        double y = std::cos(theta - beta);
        if (y < 0.0) y = 0.0;
        return (std::pow(y, n) * (std::cos(theta) / 2.0));
    };
    auto specTerm = [&](double alpha, double beta, double n) {
        double a = alpha - M_PI / 2.0, b = M_PI / 2.0;
        int m = 10;
        double hh = (b - a) / (2.0 * m), integral = 0.0;
        for (int i = 0; i < 2 * m; i += 2) {
            double xi = a + i * hh;
            integral += 2.0 * integrand(beta, n, xi);
            xi = a + (i + 1) * hh;
            integral += 4.0 * integrand(beta, n, xi);
        }
        integral += integrand(beta, n, b);
        integral -= integrand(beta, n, a);
        integral *= hh / 3.0;
        return integral;
    };
    int idx = 0;
    double invResX = 1.0 / (double)(w - 1), invResY = 1.0 / (double)(h - 1);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double t1 = (double)x * invResX, t2 = (double)y * invResY;
            double lt = 2.0 * t1 - 1.0, vt = 2.0 * t2 - 1.0;
            double diffuse = std::sqrt(1.0 - lt * lt);
            diffuse = std::pow(diffuse, (double)diffExp);
            double dotproduct = lt * vt - std::sqrt(1.0 - lt * lt) * std::sqrt(1.0 - vt * vt);
            dotproduct = (dotproduct < -1.0) ? -1.0 : ((dotproduct > 1.0) ? 1.0 : dotproduct);
            double traditionalDiff = ambient * lightColor + diffuse * diffuseM * lightColor;
            double traditionalSpec = std::pow(std::fabs(dotproduct), static_cast<double>(spec_exp)) * specularM * lightColor;
            traditionalDiff *= 0.5;
            traditionalDiff = (traditionalDiff < 0.0) ? 0.0 : ((traditionalDiff > 1.0) ? 1.0 : traditionalDiff);
            traditionalSpec *= 0.5;
            traditionalSpec = (traditionalSpec < 0.0) ? 0.0 : ((traditionalSpec > 1.0) ? 1.0 : traditionalSpec);
            zoeckler[idx++] = unorm8_roundtrip(static_cast<float>(traditionalDiff));
            zoeckler[idx++] = unorm8_roundtrip(static_cast<float>(traditionalSpec) * 0.9f);
        }
    idx = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double s = ((double)x + 0.5) / w, t = ((double)y + 0.5) / h;
            double alpha = std::acos(2.0 * s - 1.0), beta = std::acos(2.0 * t - 1.0);
            double lt = 2.0 * t - 1.0;
            double diffuse = std::sqrt(1.0 - lt * lt) * (std::sin(alpha) + (M_PI - alpha) * std::cos(alpha)) * 0.25;
            double specular = 3.5 * specTerm(alpha, beta, spec_exp);
            double color = diffuse * diffuseM * lightColor;
            color = (color < 0.0) ? 0.0 : ((color > 1.0) ? 1.0 : color);
            mallo_diff[idx] = unorm8_roundtrip((float)color);
            color = specular * specularM * lightColor;
            color = (color < 0.0) ? 0.0 : ((color > 1.0) ? 1.0 : color);
            mallo_spec[idx] = unorm8_roundtrip((float)color);
            idx++;
        }
}

float vvo_half_round(float x) { return half_round(x); }

int vvo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
