// vv_device.cuh -- parameter block shared by host and device, and the software samplers.
//
// Data layout in HBM (see DESIGN.md "Layouts"):
//   vector field   LAYOUT_PAIR : uint4 [z][y][x] = { half4 T[x], half4 T[min(x+1,nx-1)] }   16 B/voxel
//                  LAYOUT_F4   : float4 [z][y][x]                                             16 B/voxel
//                  T = the reference's RGBA16F texture contents (VV/dataset.cpp:290-366): rgb = 0.5 v/|v| + 0.5,
//                  a = |v|/max|v|, rounded to fp16 -- so both layouts hold the same values.
//   u8 volumes     cell8 : uint2 [z][y][x] = the 8 corner bytes of the trilinear cell whose low corner is
//                  (x,y,z), wrap mode (REPEAT for noise, CLAMP_TO_EDGE for the scalar volume) baked in;
//                  byte order x fastest: (x0y0z0, x1y0z0, x0y1z0, x1y1z0, x0y0z1, ...).
//   RGBA8 noise    quad : uint4 [z][y][x] = RGBA8 texels (x0y0, x1y0, x0y1, x1y1) of plane z, REPEAT baked in.
//   tables         TF RGBA as float4[256], LIC-opacity as float[256], per-step filter-kernel weights as float[]
//                  (staged into shared memory by every CTA).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vvb200 {

enum { LAYOUT_F4 = 0, LAYOUT_PAIR = 1 };
enum { ILLUM_NONE = 0, ILLUM_GRADIENT = 1, ILLUM_MALLO = 2, ILLUM_ZOECKLER = 3 };
enum { TF_B = 0, TF_A = 1, TF_R = 2, TF_LENGTH = 3, TF_SCALAR = 4 };
enum { GATE_ALWAYS = 0, GATE_TF_ALPHA = 1 };

constexpr int kBlockDim = 16;        // 16x16-pixel image blocks, one per CTA iteration
constexpr int kBlockPixels = 256;
constexpr int kMaxLicSteps = 1024;   // per direction (weights live in shared memory)

struct DevParams {
    // ---- textures ----
    const uint4  *field_pair;
    const float4 *field_f4;
    int fnx, fny, fnz;
    const uint2  *scalar_cell;
    int snx, sny, snz;
    const uint2  *noise_cell;     // scalar noise (LUMINANCE / .a channel)
    const uint4  *noise_quad;     // RGBA noise (gradient build)
    int nnx, nny, nnz;
    const float  *licvol;         // fp32 scalar LIC volume, sampled REPEAT
    int lnx, lny, lnz;
    const float4 *tf_rgba;        // [256]
    const float  *tf_opac;        // [256]
    const float  *kw;             // [0] centre, [1..nBwd] backward, [1+nBwd .. nBwd+nFwd] forward
    const float  *illum2d[3];     // zoeckler (2ch), mallo diffuse, mallo specular
    int illum_w, illum_h;
    // ---- uniforms (VV/renderer.cpp:925-996) ----
    float stepSize, gradScale, illumScale, freq;
    float h;                      // licParams.z * (logEyeDist*0.5 + 0.3), logEyeDist = 0 (Q3)
    float licScale;               // licKernel.b * gradient.r
    float alphaCorr, specExp;
    int   numIter, nFwd, nBwd;
    float texMax[3], scaleVol[3], scaleVolInv[3], lightPos[3], camera[3];
    // ---- view (double: bit-identical ray set-up on host oracle and device) ----
    double camD[3], rot[9], tanHalf, aspect, extent[3];
    int width, height;
    int tfMode, gateMode, quirkLumAlpha;
    // ---- partition / outputs ----
    int rank, world, nBlocksX, nBlocksY, nLocalBlocks;
    float4 *tiles;
    unsigned long long *sampleCounter;
    unsigned int *blockCounter;
    unsigned int *samplesPerPixel;   // optional, block-major like tiles
    // ---- sample-parallel pipeline ----
    float4 *rayA, *rayB;             // [tile*32 + lane]: (pos0.xyz, n) and (dir.xyz, state)
    uint2  *tileRec;                 // [tile]: (first src row, max samples of the tile)
    float4 *src;                     // [(row + k) * 32 + lane]: shaded ray samples, w < 0 = gated off
    uint2  *items, *itemsNext;       // work items (tile, k) of the current / next depth window
    unsigned int *itemCount, *itemCountNext, *itemHead, *slotAlloc, *nMaxGlobal;
    int win0, win1, win2;            // current window [win0, win1), next window ends at win2
    // ---- LIC volume target ----
    float *licvol_out;
    int ow, oh, od, oz0, oz1, licvolFp16;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lerpf(float a, float b, float f) { return fmaf(f, b - a, a); }

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }

// CLAMP_TO_EDGE + LINEAR on one axis: u = s n - 0.5 clamped to [0, n-1] (same value as clamping the
// two texel indices, GL 2.1 spec 3.8.8)
__device__ __forceinline__ void axis_clamp(float s, int n, int &i0, int &i1, float &f)
{
    float u = fmaf(s, (float)n, -0.5f);
    u = fminf(fmaxf(u, 0.0f), (float)(n - 1));
    float fl = floorf(u);
    f = u - fl;
    i0 = (int)fl;
    i1 = min(i0 + 1, n - 1);
}

// REPEAT + LINEAR: s' = s - floor(s); u = s' n - 0.5; cell index = floor(u) mod n
__device__ __forceinline__ void axis_repeat(float s, int n, int &i0, float &f)
{
    s = s - floorf(s);
    float u = fmaf(s, (float)n, -0.5f);
    float fl = floorf(u);
    f = u - fl;
    i0 = (int)fl;
    if (i0 < 0) i0 += n;
    if (i0 >= n) i0 -= n;
}

__device__ __forceinline__ float4 ld_f4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ uint4 ld_u4(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint2 ld_u2(const uint2 *p) { return __ldg(p); }

__device__ __forceinline__ float2 h2f(unsigned int w)
{
    __half2 h = *reinterpret_cast<__half2 *>(&w);
    return __half22float2(h);
}

// ---- vector field: trilinear RGBA16F fetch, CLAMP_TO_EDGE (volumeSampler, VV/dataset.cpp:350-357) ----
template <int LAYOUT, bool ALPHA>
__device__ __forceinline__ float4 fetch_field(const DevParams &P, float px, float py, float pz)
{
    int x0, x1, y0, y1, z0, z1;
    float fx, fy, fz;
    axis_clamp(px, P.fnx, x0, x1, fx);
    axis_clamp(py, P.fny, y0, y1, fy);
    axis_clamp(pz, P.fnz, z0, z1, fz);
    const unsigned int row = (unsigned int)P.fnx;
    const unsigned int slab = row * (unsigned int)P.fny;
    const unsigned int b00 = (unsigned int)z0 * slab + (unsigned int)y0 * row;
    const unsigned int b10 = (unsigned int)z0 * slab + (unsigned int)y1 * row;
    const unsigned int b01 = (unsigned int)z1 * slab + (unsigned int)y0 * row;
    const unsigned int b11 = (unsigned int)z1 * slab + (unsigned int)y1 * row;
    float4 r;
    if (LAYOUT == LAYOUT_PAIR) {
        const uint4 *F = P.field_pair;
        uint4 a = ld_u4(F + b00 + x0), b = ld_u4(F + b10 + x0), c = ld_u4(F + b01 + x0), d = ld_u4(F + b11 + x0);
        float2 a0 = h2f(a.x), a2 = h2f(a.z), b0 = h2f(b.x), b2 = h2f(b.z);
        float2 c0 = h2f(c.x), c2 = h2f(c.z), d0 = h2f(d.x), d2 = h2f(d.z);
        float2 a1 = h2f(a.y), a3 = h2f(a.w), b1 = h2f(b.y), b3 = h2f(b.w);
        float2 c1 = h2f(c.y), c3 = h2f(c.w), d1 = h2f(d.y), d3 = h2f(d.w);
        // x-lerp inside each pair
        float ar = lerpf(a0.x, a2.x, fx), ag = lerpf(a0.y, a2.y, fx), ab = lerpf(a1.x, a3.x, fx);
        float br = lerpf(b0.x, b2.x, fx), bg = lerpf(b0.y, b2.y, fx), bb = lerpf(b1.x, b3.x, fx);
        float cr = lerpf(c0.x, c2.x, fx), cg = lerpf(c0.y, c2.y, fx), cb = lerpf(c1.x, c3.x, fx);
        float dr = lerpf(d0.x, d2.x, fx), dg = lerpf(d0.y, d2.y, fx), db = lerpf(d1.x, d3.x, fx);
        r.x = lerpf(lerpf(ar, br, fy), lerpf(cr, dr, fy), fz);
        r.y = lerpf(lerpf(ag, bg, fy), lerpf(cg, dg, fy), fz);
        r.z = lerpf(lerpf(ab, bb, fy), lerpf(cb, db, fy), fz);
        if (ALPHA) {
            float aa = lerpf(a1.y, a3.y, fx), ba = lerpf(b1.y, b3.y, fx);
            float ca = lerpf(c1.y, c3.y, fx), da = lerpf(d1.y, d3.y, fx);
            r.w = lerpf(lerpf(aa, ba, fy), lerpf(ca, da, fy), fz);
        } else {
            r.w = 0.0f;
        }
    } else {
        const float4 *F = P.field_f4;
        float4 t000 = ld_f4(F + b00 + x0), t100 = ld_f4(F + b00 + x1);
        float4 t010 = ld_f4(F + b10 + x0), t110 = ld_f4(F + b10 + x1);
        float4 t001 = ld_f4(F + b01 + x0), t101 = ld_f4(F + b01 + x1);
        float4 t011 = ld_f4(F + b11 + x0), t111 = ld_f4(F + b11 + x1);
        r.x = lerpf(lerpf(lerpf(t000.x, t100.x, fx), lerpf(t010.x, t110.x, fx), fy),
                    lerpf(lerpf(t001.x, t101.x, fx), lerpf(t011.x, t111.x, fx), fy), fz);
        r.y = lerpf(lerpf(lerpf(t000.y, t100.y, fx), lerpf(t010.y, t110.y, fx), fy),
                    lerpf(lerpf(t001.y, t101.y, fx), lerpf(t011.y, t111.y, fx), fy), fz);
        r.z = lerpf(lerpf(lerpf(t000.z, t100.z, fx), lerpf(t010.z, t110.z, fx), fy),
                    lerpf(lerpf(t001.z, t101.z, fx), lerpf(t011.z, t111.z, fx), fy), fz);
        if (ALPHA)
            r.w = lerpf(lerpf(lerpf(t000.w, t100.w, fx), lerpf(t010.w, t110.w, fx), fy),
                        lerpf(lerpf(t001.w, t101.w, fx), lerpf(t011.w, t111.w, fx), fy), fz);
        else
            r.w = 0.0f;
    }
    return r;
}

// byte k of w -> float, exactly, without I2F: bits 0x4B0000bb = 8388608 + bb
__device__ __forceinline__ float byte_f(unsigned int w, int k)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + k)) - 8388608.0f;
}

// trilinear blend of the 8 corner bytes of one cell8 entry, result in [0,1]
__device__ __forceinline__ float cell8_blend(uint2 c, float fx, float fy, float fz)
{
    float x00 = lerpf(byte_f(c.x, 0), byte_f(c.x, 1), fx);
    float x10 = lerpf(byte_f(c.x, 2), byte_f(c.x, 3), fx);
    float x01 = lerpf(byte_f(c.y, 0), byte_f(c.y, 1), fx);
    float x11 = lerpf(byte_f(c.y, 2), byte_f(c.y, 3), fx);
    return lerpf(lerpf(x00, x10, fy), lerpf(x01, x11, fy), fz) * (1.0f / 255.0f);
}

// scalarSampler: LUMINANCE8, CLAMP_TO_EDGE (VV/dataset.cpp:1025-1038) -> .r
__device__ __forceinline__ float fetch_scalar(const DevParams &P, float px, float py, float pz)
{
    int x0, x1, y0, y1, z0, z1;
    float fx, fy, fz;
    axis_clamp(px, P.snx, x0, x1, fx);
    axis_clamp(py, P.sny, y0, y1, fy);
    axis_clamp(pz, P.snz, z0, z1, fz);
    uint2 c = ld_u2(P.scalar_cell + ((unsigned int)z0 * P.sny + y0) * P.snx + x0);
    return cell8_blend(c, fx, fy, fz);
}

// noiseSampler, LUMINANCE8 / alpha channel, REPEAT (VV/dataset.cpp:1283-1335) -> noise value
__device__ __forceinline__ float fetch_noise_scalar(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_repeat(px, P.nnx, x0, fx);
    axis_repeat(py, P.nny, y0, fy);
    axis_repeat(pz, P.nnz, z0, fz);
    uint2 c = ld_u2(P.noise_cell + ((unsigned int)z0 * P.nny + y0) * P.nnx + x0);
    return cell8_blend(c, fx, fy, fz);
}

__device__ __forceinline__ float4 rgba8_lerp4(unsigned int t00, unsigned int t10, unsigned int t01, unsigned int t11,
                                              float fx, float fy)
{
    float4 r;
    r.x = lerpf(lerpf(byte_f(t00, 0), byte_f(t10, 0), fx), lerpf(byte_f(t01, 0), byte_f(t11, 0), fx), fy);
    r.y = lerpf(lerpf(byte_f(t00, 1), byte_f(t10, 1), fx), lerpf(byte_f(t01, 1), byte_f(t11, 1), fx), fy);
    r.z = lerpf(lerpf(byte_f(t00, 2), byte_f(t10, 2), fx), lerpf(byte_f(t01, 2), byte_f(t11, 2), fx), fy);
    r.w = lerpf(lerpf(byte_f(t00, 3), byte_f(t10, 3), fx), lerpf(byte_f(t01, 3), byte_f(t11, 3), fx), fy);
    return r;
}

// noiseSampler, RGBA8 (gradient.xyz, noise), REPEAT -> raw texel (freqSamplingGrad, inc_lic.glsl:61-68)
__device__ __forceinline__ float4 fetch_noise_rgba(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_repeat(px, P.nnx, x0, fx);
    axis_repeat(py, P.nny, y0, fy);
    axis_repeat(pz, P.nnz, z0, fz);
    int z1 = z0 + 1;
    if (z1 >= P.nnz) z1 = 0;
    uint4 a = ld_u4(P.noise_quad + ((unsigned int)z0 * P.nny + y0) * P.nnx + x0);
    uint4 b = ld_u4(P.noise_quad + ((unsigned int)z1 * P.nny + y0) * P.nnx + x0);
    float4 p0 = rgba8_lerp4(a.x, a.y, a.z, a.w, fx, fy);
    float4 p1 = rgba8_lerp4(b.x, b.y, b.z, b.w, fx, fy);
    const float k = 1.0f / 255.0f;
    float4 r;
    r.x = lerpf(p0.x, p1.x, fz) * k;
    r.y = lerpf(p0.y, p1.y, fz) * k;
    r.z = lerpf(p0.z, p1.z, fz) * k;
    r.w = lerpf(p0.w, p1.w, fz) * k;
    return r;
}

// licVolumeSampler: fp32 scalar, REPEAT (VV/VolumeBuffer.cpp:47-55), only .r
__device__ __forceinline__ float fetch_licvol(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_repeat(px, P.lnx, x0, fx);
    axis_repeat(py, P.lny, y0, fy);
    axis_repeat(pz, P.lnz, z0, fz);
    int x1 = (x0 + 1 >= P.lnx) ? 0 : x0 + 1;
    int y1 = (y0 + 1 >= P.lny) ? 0 : y0 + 1;
    int z1 = (z0 + 1 >= P.lnz) ? 0 : z0 + 1;
    const float *L = P.licvol;
    const unsigned int row = P.lnx, slab = (unsigned int)P.lnx * P.lny;
    float t000 = __ldg(L + z0 * slab + y0 * row + x0), t100 = __ldg(L + z0 * slab + y0 * row + x1);
    float t010 = __ldg(L + z0 * slab + y1 * row + x0), t110 = __ldg(L + z0 * slab + y1 * row + x1);
    float t001 = __ldg(L + z1 * slab + y0 * row + x0), t101 = __ldg(L + z1 * slab + y0 * row + x1);
    float t011 = __ldg(L + z1 * slab + y1 * row + x0), t111 = __ldg(L + z1 * slab + y1 * row + x1);
    return lerpf(lerpf(lerpf(t000, t100, fx), lerpf(t010, t110, fx), fy),
                 lerpf(lerpf(t001, t101, fx), lerpf(t011, t111, fx), fy), fz);
}

// 1-D 256-entry tables in shared memory, LINEAR + CLAMP_TO_EDGE (VV/transferEdit.cpp:515-540)
__device__ __forceinline__ float4 tf_lookup(const float4 *s_tf, float x)
{
    int i0, i1;
    float f;
    axis_clamp(x, 256, i0, i1, f);
    float4 a = s_tf[i0], b = s_tf[i1], r;
    r.x = lerpf(a.x, b.x, f); r.y = lerpf(a.y, b.y, f); r.z = lerpf(a.z, b.z, f); r.w = lerpf(a.w, b.w, f);
    return r;
}
__device__ __forceinline__ float opac_lookup(const float *s_opac, float x)
{
    int i0, i1;
    float f;
    axis_clamp(x, 256, i0, i1, f);
    return lerpf(s_opac[i0], s_opac[i1], f);
}

} // namespace vvb200
