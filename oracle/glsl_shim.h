/*
 * glsl_shim.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A GLSL 1.20 "front end" made of C++: vec2/vec3/vec4 with the swizzles the reference's shaders use, the GLSL
 * built-ins they call, sampler types whose texture*() functions implement the OpenGL 2.1 filtering rules
 * (spec 3.8.8: LINEAR, texel centres at i + 0.5; CLAMP_TO_EDGE / REPEAT / CLAMP with a zero border; UNORM8 /
 * float decode; LUMINANCE and LUMINANCE_ALPHA expansion), and the gl_* state the shaders read.
 *
 * oracle/build_ref.py runs the reference's own shader sources (VV/shader/*.glsl, read where they lie under
 * /root/reference) through a purely lexical rewrite (parameter qualifiers, float literal suffixes, main ->
 * shader_main, `float logEyeDist;` -> `= 0.0f` (Q3)) into oracle/_ref/, and compiles them against this header.
 * What executes is the reference's shader code; what this header supplies is the part of the GL that the
 * reference leaves to the driver.
 */
#pragma once
#include <cmath>
#include <cstdint>

namespace glsl {

struct vec2;
struct vec3;
struct vec4;

// ---- swizzle proxies: live in a union with the parent's float[N] ---------------------------------
template <int N, int A, int B>
struct S2 {
    float d[N];
    operator vec2() const;
    S2 &operator=(const vec2 &v);
    S2 &operator=(const S2 &o) { float a = o.d[A], b = o.d[B]; d[A] = a; d[B] = b; return *this; }
};
template <int N, int A, int B, int C>
struct S3 {
    float d[N];
    operator vec3() const;
    S3 &operator=(const vec3 &v);
    S3 &operator=(const S3 &o) { float a = o.d[A], b = o.d[B], c = o.d[C]; d[A] = a; d[B] = b; d[C] = c; return *this; }
    S3 &operator*=(float s) { d[A] *= s; d[B] *= s; d[C] *= s; return *this; }
    S3 &operator*=(const vec3 &v);
    S3 &operator+=(const vec3 &v);
};

struct vec2 {
    union {
        float d[2];
        struct { float x, y; };
        struct { float r, g; };
        S2<2, 0, 1> xy;
        S2<2, 0, 1> rg;
    };
    vec2() : d{0, 0} {}
    explicit vec2(float s) : d{s, s} {}
    vec2(float a, float b) : d{a, b} {}
    vec2(const vec2 &o) : d{o.d[0], o.d[1]} {}
    vec2 &operator=(const vec2 &o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec2 &operator*=(float s) { d[0] *= s; d[1] *= s; return *this; }
};

struct vec3 {
    union {
        float d[3];
        struct { float x, y, z; };
        struct { float r, g, b; };
        S3<3, 0, 1, 2> xyz;
        S3<3, 0, 1, 2> rgb;
        S2<3, 0, 1> xy;
        S2<3, 0, 1> rg;
    };
    vec3() : d{0, 0, 0} {}
    explicit vec3(float s) : d{s, s, s} {}
    vec3(float a, float b, float c) : d{a, b, c} {}
    vec3(const vec2 &v, float c) : d{v.d[0], v.d[1], c} {}
    vec3(const vec3 &o) : d{o.d[0], o.d[1], o.d[2]} {}
    vec3 &operator=(const vec3 &o) { d[0] = o.d[0]; d[1] = o.d[1]; d[2] = o.d[2]; return *this; }
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec3 &operator*=(float s) { d[0] *= s; d[1] *= s; d[2] *= s; return *this; }
    vec3 &operator*=(const vec3 &o) { d[0] *= o.d[0]; d[1] *= o.d[1]; d[2] *= o.d[2]; return *this; }
    vec3 &operator+=(const vec3 &o) { d[0] += o.d[0]; d[1] += o.d[1]; d[2] += o.d[2]; return *this; }
    vec3 &operator-=(const vec3 &o) { d[0] -= o.d[0]; d[1] -= o.d[1]; d[2] -= o.d[2]; return *this; }
};

struct vec4 {
    union {
        float d[4];
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        S3<4, 0, 1, 2> xyz;
        S3<4, 0, 1, 2> rgb;
        S2<4, 0, 1> xy;
        S2<4, 0, 1> rg;
        S2<4, 2, 3> zw;
        S2<4, 0, 2> xz;
        S2<4, 2, 1> zy;
    };
    vec4() : d{0, 0, 0, 0} {}
    explicit vec4(float s) : d{s, s, s, s} {}
    vec4(float a, float b, float c, float e) : d{a, b, c, e} {}
    vec4(const vec3 &v, float e) : d{v.d[0], v.d[1], v.d[2], e} {}
    vec4(const vec4 &o) : d{o.d[0], o.d[1], o.d[2], o.d[3]} {}
    vec4 &operator=(const vec4 &o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; return *this; }
    float &operator[](int i) { return d[i]; }
    float operator[](int i) const { return d[i]; }
    vec4 &operator*=(float s) { for (int i = 0; i < 4; ++i) d[i] *= s; return *this; }
    vec4 &operator+=(const vec4 &o) { for (int i = 0; i < 4; ++i) d[i] += o.d[i]; return *this; }
};

template <int N, int A, int B> inline S2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B> inline S2<N, A, B> &S2<N, A, B>::operator=(const vec2 &v) { d[A] = v.d[0]; d[B] = v.d[1]; return *this; }
template <int N, int A, int B, int C> inline S3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int N, int A, int B, int C> inline S3<N, A, B, C> &S3<N, A, B, C>::operator=(const vec3 &v) { d[A] = v.d[0]; d[B] = v.d[1]; d[C] = v.d[2]; return *this; }
template <int N, int A, int B, int C> inline S3<N, A, B, C> &S3<N, A, B, C>::operator*=(const vec3 &v) { d[A] *= v.d[0]; d[B] *= v.d[1]; d[C] *= v.d[2]; return *this; }
template <int N, int A, int B, int C> inline S3<N, A, B, C> &S3<N, A, B, C>::operator+=(const vec3 &v) { d[A] += v.d[0]; d[B] += v.d[1]; d[C] += v.d[2]; return *this; }

struct bvec4 {
    bool d[4];
    bvec4(const vec3 &v, bool w) : d{v.d[0] != 0.0f, v.d[1] != 0.0f, v.d[2] != 0.0f, w} {}
};
inline bool any(const bvec4 &b) { return b.d[0] || b.d[1] || b.d[2] || b.d[3]; }

// ---- component-wise operators ------------------------------------------------------------------------
#define GLSL_VEC_OPS(V, N)                                                                                         \
    inline V operator+(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; } \
    inline V operator-(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; } \
    inline V operator*(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.d[i]; return r; } \
    inline V operator/(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b.d[i]; return r; } \
    inline V operator+(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + s; return r; }        \
    inline V operator-(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - s; return r; }        \
    inline V operator*(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }        \
    inline V operator/(const V &a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / s; return r; }        \
    inline V operator+(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s + a.d[i]; return r; }        \
    inline V operator-(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s - a.d[i]; return r; }        \
    inline V operator*(float s, const V &a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }        \
    inline V operator-(const V &a) { V r; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

// ---- built-ins ------------------------------------------------------------------------------------------
inline float dot(const vec3 &a, const vec3 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2]; }
inline float dot(const vec4 &a, const vec4 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2] + a.d[3] * b.d[3]; }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline float length(const vec4 &a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { float l = std::sqrt(dot(a, a)); return vec3(a.d[0] / l, a.d[1] / l, a.d[2] / l); }
inline vec3 cross(const vec3 &a, const vec3 &b)
{
    return vec3(a.d[1] * b.d[2] - a.d[2] * b.d[1], a.d[2] * b.d[0] - a.d[0] * b.d[2], a.d[0] * b.d[1] - a.d[1] * b.d[0]);
}
inline float clamp(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
inline vec3 clamp(const vec3 &v, float lo, float hi) { return vec3(clamp(v.d[0], lo, hi), clamp(v.d[1], lo, hi), clamp(v.d[2], lo, hi)); }
inline vec3 clamp(const vec3 &v, const vec3 &lo, const vec3 &hi)
{
    return vec3(clamp(v.d[0], lo.d[0], hi.d[0]), clamp(v.d[1], lo.d[1], hi.d[1]), clamp(v.d[2], lo.d[2], hi.d[2]));
}
inline vec4 clamp(const vec4 &v, float lo, float hi)
{
    return vec4(clamp(v.d[0], lo, hi), clamp(v.d[1], lo, hi), clamp(v.d[2], lo, hi), clamp(v.d[3], lo, hi));
}
/* mix(x, y, a) = x (1 - a) + y a   (GLSL 1.20 spec 8.3) */
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(const vec3 &x, const vec3 &y, float a) { return x * (1.0f - a) + y * a; }
inline vec4 mix(const vec4 &x, const vec4 &y, float a) { return x * (1.0f - a) + y * a; }
inline float pow(float x, float y) { return std::pow(x, y); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float log2(float x) { return std::log2(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float floor(float x) { return std::floor(x); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float abs(float a) { return std::fabs(a); }

// ---- textures ------------------------------------------------------------------------------------------------
enum Wrap { W_CLAMP_TO_EDGE = 0, W_REPEAT = 1, W_CLAMP = 2 /* border colour (0,0,0,0) */ };
enum Fmt { F_L8 = 0, F_LA8 = 1, F_RGBA8 = 2, F_RGBA32F = 3, F_L32F = 4, F_LA32F = 5, F_FBO = 6 /* the frame buffer under the fragment */ };

struct Texture {
    const void *data = nullptr;
    int dim[3] = {1, 1, 1};
    int fmt = F_L8;
    int wrap = W_CLAMP_TO_EDGE;
    int id = 0;              /* for per-sampler fetch counters */
};
typedef const Texture *sampler1D;
typedef const Texture *sampler2D;
typedef const Texture *sampler3D;
typedef const Texture *sampler2DRect;

extern thread_local uint32_t g_fetch_count[16];
extern thread_local float g_fbo_dest[4];   /* imageFBOSampler content at gl_FragCoord (slicing ping-pong target) */

struct AxisL { int i0, i1; float f; bool b0, b1; };

inline AxisL axis_linear(float s, int n, int wrap)
{
    AxisL a;
    a.b0 = a.b1 = false;
    if (wrap == W_REPEAT) s = s - std::floor(s);
    else if (wrap == W_CLAMP) s = clamp(s, 0.0f, 1.0f);
    float u = s * (float)n - 0.5f;
    float fl = std::floor(u);
    a.f = u - fl;
    int i0 = (int)fl, i1 = i0 + 1;
    if (wrap == W_REPEAT) {
        i0 = ((i0 % n) + n) % n;
        i1 = ((i1 % n) + n) % n;
    } else {
        if (wrap == W_CLAMP) { a.b0 = (i0 < 0 || i0 >= n); a.b1 = (i1 < 0 || i1 >= n); }
        i0 = i0 < 0 ? 0 : (i0 > n - 1 ? n - 1 : i0);
        i1 = i1 < 0 ? 0 : (i1 > n - 1 ? n - 1 : i1);
    }
    a.i0 = i0; a.i1 = i1;
    return a;
}

inline vec4 texel(const Texture *t, int x, int y, int z)
{
    size_t i = ((size_t)z * t->dim[1] + y) * t->dim[0] + x;
    switch (t->fmt) {
    case F_L8: { float l = (float)((const uint8_t *)t->data)[i] / 255.0f; return vec4(l, l, l, 1.0f); }
    case F_LA8: { const uint8_t *p = (const uint8_t *)t->data + 2 * i; float l = (float)p[0] / 255.0f; return vec4(l, l, l, (float)p[1] / 255.0f); }
    case F_RGBA8: { const uint8_t *p = (const uint8_t *)t->data + 4 * i; return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f); }
    case F_RGBA32F: { const float *p = (const float *)t->data + 4 * i; return vec4(p[0], p[1], p[2], p[3]); }
    case F_L32F: { float l = ((const float *)t->data)[i]; return vec4(l, l, l, 1.0f); }
    default: { const float *p = (const float *)t->data + 2 * i; return vec4(p[0], p[0], p[0], p[1]); }
    }
}

inline vec4 lerp4(const vec4 &a, const vec4 &b, float f) { return a * (1.0f - f) + b * f; }

inline vec4 sample_linear(const Texture *t, float s, float tt, float r, int nd)
{
    g_fetch_count[t->id & 15]++;
    AxisL ax = axis_linear(s, t->dim[0], t->wrap);
    AxisL ay = {0, 0, 0.0f, false, false}, az = {0, 0, 0.0f, false, false};
    if (nd >= 2) ay = axis_linear(tt, t->dim[1], t->wrap);
    if (nd >= 3) az = axis_linear(r, t->dim[2], t->wrap);
    const vec4 border(0.0f, 0.0f, 0.0f, 0.0f);
    auto fetch = [&](int x, bool bx, int y, bool by, int z, bool bz) { return (bx || by || bz) ? border : texel(t, x, y, z); };
    vec4 c00 = lerp4(fetch(ax.i0, ax.b0, ay.i0, ay.b0, az.i0, az.b0), fetch(ax.i1, ax.b1, ay.i0, ay.b0, az.i0, az.b0), ax.f);
    if (nd == 1) return c00;
    vec4 c10 = lerp4(fetch(ax.i0, ax.b0, ay.i1, ay.b1, az.i0, az.b0), fetch(ax.i1, ax.b1, ay.i1, ay.b1, az.i0, az.b0), ax.f);
    vec4 c0 = lerp4(c00, c10, ay.f);
    if (nd == 2) return c0;
    vec4 c01 = lerp4(fetch(ax.i0, ax.b0, ay.i0, ay.b0, az.i1, az.b1), fetch(ax.i1, ax.b1, ay.i0, ay.b0, az.i1, az.b1), ax.f);
    vec4 c11 = lerp4(fetch(ax.i0, ax.b0, ay.i1, ay.b1, az.i1, az.b1), fetch(ax.i1, ax.b1, ay.i1, ay.b1, az.i1, az.b1), ax.f);
    vec4 c1 = lerp4(c01, c11, ay.f);
    return lerp4(c0, c1, az.f);
}

inline vec4 texture1D(sampler1D t, float s) { return sample_linear(t, s, 0.0f, 0.0f, 1); }
inline vec4 texture2D(sampler2D t, const vec2 &p) { return sample_linear(t, p.d[0], p.d[1], 0.0f, 2); }
inline vec4 texture3D(sampler3D t, const vec3 &p) { return sample_linear(t, p.d[0], p.d[1], p.d[2], 3); }
/* rectangle textures use unnormalised coordinates; only reached by the slicing / background / MC-offset paths */
inline vec4 texture2DRect(sampler2DRect t, const vec2 &p)
{
    if (t->fmt == F_FBO) return vec4(g_fbo_dest[0], g_fbo_dest[1], g_fbo_dest[2], g_fbo_dest[3]);
    int x = (int)std::floor(p.d[0]), y = (int)std::floor(p.d[1]);
    x = x < 0 ? 0 : (x > t->dim[0] - 1 ? t->dim[0] - 1 : x);
    y = y < 0 ? 0 : (y > t->dim[1] - 1 ? t->dim[1] - 1 : y);
    return texel(t, x, y, 0);
}

// ---- gl_* state --------------------------------------------------------------------------------------------------
struct mat4 {
    vec4 c[4];
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
struct LightSource {
    vec4 position, ambient, diffuse, specular;
    float spotExponent;
};
extern thread_local vec4 gl_TexCoord[8];
extern thread_local vec4 gl_FragColor;
extern thread_local vec4 gl_FragCoord;
extern mat4 gl_ModelViewMatrixInverse;
extern mat4 gl_ModelViewProjectionMatrix;
extern LightSource gl_LightSource[1];

} // namespace glsl
