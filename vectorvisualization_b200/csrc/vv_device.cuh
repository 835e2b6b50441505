// vv_device.cuh -- parameter block shared by host and device, and the software samplers.
//
// Data layout in HBM (see DESIGN.md "Layouts"):
//   vector field   LAYOUT_PAIR : uint4 [z][y][x] = { half4 T[x], half4 T[min(x+1,nx-1)] }   16 B/voxel, G guard cells per side
//                  LAYOUT_QUAD : 2 x uint4 [z][y][x] = the pairs of rows y and y+1 (the xy face of the trilinear cell in one
//                                32-byte sector, read with one 256-bit load)                  32 B/voxel, same guard geometry
//                  LAYOUT_F4   : float4 [z][y][x]                                             16 B/voxel
//                  T = the reference's RGBA16F texture contents (VV/dataset.cpp:290-366): rgb = 0.5 v/|v| + 0.5,
//                  a = |v|/max|v|, rounded to fp16 -- so both layouts hold the same values.
//   u8 volumes     cell8 : uint2 [z][y][x] = the 8 corner bytes of the trilinear cell whose low corner is
//                  (x,y,z), wrap mode (REPEAT for noise, CLAMP_TO_EDGE for the scalar volume) baked in;
//                  byte order x fastest: (x0y0z0, x1y0z0, x0y1z0, x1y1z0, x0y0z1, ...).
//   RGBA8 noise    pair : uint4 = { half4 N[x], half4 N[x+1] }, bytes 0..255 as fp16 integers (hot layout)
//                  quad : uint4 [z][y][x] = RGBA8 texels (x0y0, x1y0, x0y1, x1y1) of plane z, REPEAT baked in (check).
//   padding        hot layouts carry their wrap mode as border rows / planes (edge replicated for CLAMP_TO_EDGE, wrapped
//                  for REPEAT, where the cell index floor(u) lies in [-1, n-1]), so the sampler never clamps or wraps
//                  an index: neighbours are +row / +plane.
//   tables         TF RGBA as float4[256], LIC-opacity as float[256], per-step filter-kernel weights as float[]
//                  (staged into shared memory by every CTA).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vvb200 {

enum { LAYOUT_F4 = 0, LAYOUT_PAIR = 1, LAYOUT_QUAD = 2 };
enum { ILLUM_NONE = 0, ILLUM_GRADIENT = 1, ILLUM_MALLO = 2, ILLUM_ZOECKLER = 3 };
enum { TF_B = 0, TF_A = 1, TF_R = 2, TF_LENGTH = 3, TF_SCALAR = 4 };
enum { GATE_ALWAYS = 0, GATE_TF_ALPHA = 1 };

constexpr int kBlockDim = 16;        // 16x16-pixel image blocks, one per CTA iteration
constexpr int kBlockPixels = 256;
constexpr int kMaxLicSteps = 1024;   // per direction (weights live in shared memory)
#ifndef VV_CK_STRIDE
#define VV_CK_STRIDE 16
#endif
constexpr int kCkStride = VV_CK_STRIDE;   // ray-march positions kept every kCkStride steps (ray_checkpoint_kernel)

struct DevParams {
    // ---- textures ----
    const uint4  *field_pair;     // padded [nz+2G][ny+2G][nx+2G] (edge replicated, G = fGuard >= 1), pointing at cell (0,0,0): neighbours are
                                  // +fRow / +fPlane, no index clamp; cells [-G, n-1+G] are addressable (signed element offsets)
    const float4 *field_f4;
    int fnx, fny, fnz;
    float fnf[3], fnm1f[3];       // (float)n and (float)(n-1) per axis: no I2FP of loop invariants in the walk
    unsigned int fRow, fPlane;    // element strides of field_pair
    int fGuard;                   // guard cells around the field; the unclamped samplers (XF_GUARD) need every walk position inside it
    int guardOk;                  // the host checked that fGuard covers this frame's longest walk (else the clamped samplers run)
    int noiseSameDims;            // RGBA noise has the field's dimensions: its cell coordinates inside [0,1)^3 are the field's (XF_NSHARE)
    int noiseShared;              // ... and noise_bf is stored with the field's guard geometry: the field's cell index addresses it
    float walkReach;              // (S + 2) h: a walk started at least this far inside [0,1)^3 never leaves it
    const uint2  *scalar_cell;
    int snx, sny, snz;
    float snf[3], snm1f[3];
    // REPEAT textures are stored with a wrapped border so that the cell index floor(u) in [-1, n-1] addresses memory
    // directly: the pointers below already point at cell (0,0,0) of the padded arrays (signed element offsets).
    const uint2  *noise_cell;     // scalar noise (LUMINANCE / .a channel), cell8, padded [nz+1][ny+1][nx+1]
    const uint4  *noise_quad;     // RGBA noise (gradient build), u8 xy-quad layout (unpadded check layout)
    const uint4  *noise_pair;     // RGBA noise as fp16 x-pairs {half4 T[x], half4 T[x+1]} padded [nz+2][ny+2][nx+1] (used when non-null)
    const uint4  *noise_bf;       // RGBA noise as bf16 {T[x], T[x+1] - T[x]} (bytes and their differences are exact in bf16), same padding
    int nnx, nny, nnz;
    float nnf[3];
    int ncRow, ncPlane, npRow, npPlane;   // element strides of noise_cell / noise_pair
    int nbRow, nbPlane;                   // element strides of noise_bf
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
    int   nFwdEff, nBwdEff;       // steps up to the last non-zero filter-kernel weight
    float texMax[3], scaleVol[3], scaleVolInv[3], lightPos[3], camera[3];
    // ---- view (double: bit-identical ray set-up on host oracle and device) ----
    double camD[3], rot[9], tanHalf, aspect, extent[3];
    double nearD, farD;           // gluPerspective near / far: GL clips the proxy geometry to the view volume, a fragment exists only
                                  // where its eye-space depth (= the pixel-ray parameter t, d = R^T (ex, ey, -1)) lies in [near, far]
    int width, height;
    int tfMode, gateMode, quirkLumAlpha;
    // ---- Monte-Carlo ray-start offsets (USE_MC_OFFSET) and user clip planes ----
    const float *mcOffsets;          // [height][width], fp16-rounded values in [0,1]; null = off
    int nClip;                       // active clip planes
    double clipEq[3][4];             // glClipPlane equations (n / |n|, d), n.q + d >= 0 kept, q = position - centerD
    double clipN[3][3], clipDist[3]; // unit normal and distance of the cap polygon, n^.q = -(d - 0.0001); clipDist NaN = no cap
    double centerD[3];
    // ---- partition / outputs ----
    int rank, world, nBlocksX, nBlocksY, nLocalBlocks;
    int blockSkew;                   // per-row rotation of the block ids, see block_id()
    int partUnit;                    // the partition deals units of partUnit x partUnit blocks to the ranks, see block_id_u()
    float4 *tiles;
    unsigned long long *sampleCounter;
    unsigned int *blockCounter;
    unsigned int *samplesPerPixel;   // optional, block-major like tiles
    // ---- sample-parallel pipeline ----
    float4 *rayA, *rayB;             // [tile*32 + lane]: (pos0.xyz, n) and (dir.xyz, state)
    uint2  *tileRec;                 // [tile]: (first src row, max samples of the tile)
    float4 *rayCk;                   // march checkpoints: [(tileCk[tile] + j - 1) * 32 + lane] = the ray's position after 16 j steps (j >= 1)
    unsigned int *tileCk;            // [tile]: first checkpoint row of the tile
    unsigned int *ckAlloc;           // checkpoint rows allocated by the ray set-up
    float4 *src;                     // [(row + k) * 32 + lane]: shaded ray samples, w < 0 = gated off
    uint2  *items, *itemsNext;       // work items (tile, k) of the current / next depth window
    unsigned int *itemCount, *itemCountNext, *itemHead, *slotAlloc, *nMaxGlobal;
    unsigned int *tileLive;          // [tile]: longest ray of the tile that is still alive (0: every ray of the tile has finished)
    int win0, win1, win2;            // current window [win0, win1), next window [win1, win2) (the one whose items are being built)
    int emitItems;                   // composite_kernel emits the next window's items itself, tile-major (VV_OPT_DEPTH_MAJOR = 0)
    // depth-major item order: buckets = (band of block rows) x (chunk of 8 depths)
    unsigned int *bucketCount, *bucketBase, *bucketFill;
    int bandRows, nDepthChunks;
    // ---- view-aligned slicing (VV/slicing.cpp:42-114): unit view vector, covered depth, slice count ----
    int slicing;                     // 0 off; 1 FBO path (lic3d_slicing_fragment.glsl, two RGBA16F ping-pong targets); 2 without the FBO
                                     // (lic3d_slicingblend_fragment.glsl + (ONE_MINUS_DST_ALPHA, ONE) blending in the RGBA8 back buffer)
    float slV[3], slD;
    int slNum;
    double slCenter[3];
    // ---- LIC volume target ----
    float *licvol_out;
    int ow, oh, od, oz0, oz1, licvolFp16;
};

// ------------------------------------------------------------------------------------------------
// Sort-first partition.  Image block (bx, by) has id  by * nbx + (bx + skew * by) mod nbx ; rank (id mod world) owns it
// as its local block id / world.  The per-row rotation scatters one rank's blocks over x AND y: with plain row-major
// ids and nbx a multiple of world every rank would own whole block columns (measured 9 % load imbalance at 8 GPUs on
// cfg3; 0.2 % with the rotation).  skew = 0 for world 1.
__host__ __device__ __forceinline__ int block_skew_for(int world) { return world <= 1 ? 0 : 2 * ((382 * world / 1000) / 2) + 1; }
__host__ __device__ __forceinline__ int block_id(int nbx, int skew, int bx, int by) { return by * nbx + (bx + skew * by) % nbx; }
__host__ __device__ __forceinline__ void block_xy_of(int nbx, int skew, int b, int &bx, int &by)
{
    by = b / nbx;
    bx = b % nbx - (skew * by) % nbx;
    if (bx < 0) bx += nbx;
}

// The same deal with UNITS of U x U blocks: the unit (ux, uy) = (bx / U, by / U) has the unit id block_id(ceil(nbx / U), skew, ux, uy) and
// belongs to rank (unit id mod world); a block's id is ((unit id / world) U^2 + sub) world + rank with sub = (by mod U) U + bx mod U, so
// that, as before, id mod world is the owner and id / world the local block index (units of a rank back to back, blocks of a unit
// consecutive).  Blocks of a unit that lie outside the image (nbx or nby not a multiple of U) exist as ids and hold no pixel.
// Why units: a rank's block needs the volume its rays cross plus a margin of one streamline length on every side (cfg3: 6 voxels
// + 2 x 12); isolated 16-pixel blocks make every rank fetch ~25 x the volume it owns through L2, 2 x 2 units ~9 x, 4 x 4 units ~4 x.
__host__ __device__ __forceinline__ int block_id_u(int nbx, int skew, int world, int U, int bx, int by)
{
    if (U <= 1) return block_id(nbx, skew, bx, by);
    const int uid = block_id((nbx + U - 1) / U, skew, bx / U, by / U);
    return ((uid / world) * U * U + (by % U) * U + bx % U) * world + uid % world;
}
__host__ __device__ __forceinline__ void block_xy_of_u(int nbx, int skew, int world, int U, int b, int &bx, int &by)
{
    if (U <= 1) { block_xy_of(nbx, skew, b, bx, by); return; }
    const int lb = b / world, sub = lb % (U * U);
    int ux, uy;
    block_xy_of((nbx + U - 1) / U, skew, (lb / (U * U)) * world + b % world, ux, uy);
    bx = ux * U + sub % U;
    by = uy * U + sub / U;
}

__device__ __forceinline__ float lerpf(float a, float b, float f) { return fmaf(f, b - a, a); }

__device__ __forceinline__ void block_xy(const DevParams &P, int b, int &bx, int &by) { block_xy_of_u(P.nBlocksX, P.blockSkew, P.world, P.partUnit, b, bx, by); }

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }

// CLAMP_TO_EDGE + LINEAR on one axis: u = s n - 0.5 clamped to [0, n-1] (same value as clamping the
// two texel indices, GL 2.1 spec 3.8.8)
__device__ __forceinline__ void axis_clamp(float s, int n, int &i0, int &i1, float &f)
{
    float u = fmaf(s, (float)n, -0.5f);
    u = fminf(fmaxf(u, 0.0f), (float)(n - 1));
    i0 = __float2int_rd(u);             // one conversion-pipe op; the float floor comes back through I2FP
    f = u - __int2float_rn(i0);
    i1 = min(i0 + 1, n - 1);
}

// REPEAT + LINEAR: s' = s - floor(s); u = s' n - 0.5; cell index = floor(u) mod n
__device__ __forceinline__ void axis_repeat(float s, int n, int &i0, float &f)
{
    s = s - floorf(s);
    float u = fmaf(s, (float)n, -0.5f);
    i0 = __float2int_rd(u);
    f = u - __int2float_rn(i0);
    if (i0 < 0) i0 += n;
    if (i0 >= n) i0 -= n;
}

// The same two rules with the axis size passed as floats (loop-invariant conversions hoisted to the host) and
// without the neighbour index: the padded layouts make it i0 + 1 unconditionally.
__device__ __forceinline__ void axis_clamp_f(float s, float nf, float nm1f, int &i0, float &f)
{
    float u = fmaf(s, nf, -0.5f);
    u = fminf(fmaxf(u, 0.0f), nm1f);
    i0 = __float2int_rd(u);
    f = u - __int2float_rn(i0);
}
// i0 is left unwrapped in [-1, n-1]
__device__ __forceinline__ void axis_repeat_f(float s, float nf, int &i0, float &f)
{
    s = s - floorf(s);
    float u = fmaf(s, nf, -0.5f);
    i0 = __float2int_rd(u);
    f = u - __int2float_rn(i0);
}

__device__ __forceinline__ float4 ld_f4(const float4 *p) { return __ldg(p); }
// one 256-bit load (LDG.E.256 on sm_100): two consecutive uint4, 32-byte aligned
__device__ __forceinline__ void ld_u4x2(const uint4 *p, uint4 &a, uint4 &b)
{
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ uint4 ld_u4(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint2 ld_u2(const uint2 *p) { return __ldg(p); }

__device__ __forceinline__ float2 h2f(unsigned int w)
{
    __half2 h = *reinterpret_cast<__half2 *>(&w);
    return __half22float2(h);
}

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE binary32 lanes per issue slot ----
// The hot kernel is issue-bound (profiles/README.md), and ~half of its instructions are the sub + fma of the
// trilinear lerps, so two lerps per instruction is the main lever.  Each lane is exactly the scalar operation.
typedef unsigned long long pk2_t;
__device__ __forceinline__ pk2_t pk2(float lo, float hi) { pk2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(pk2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float lo2(pk2_t v) { float a, b; up2(v, a, b); return a; }
__device__ __forceinline__ float hi2(pk2_t v) { float a, b; up2(v, a, b); return b; }
__device__ __forceinline__ pk2_t fma2(pk2_t a, pk2_t b, pk2_t c) { pk2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2_t add2(pk2_t a, pk2_t b) { pk2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2_t sub2(pk2_t a, pk2_t b) { pk2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2_t mul2(pk2_t a, pk2_t b) { pk2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2_t bc2(float f) { return pk2(f, f); }
// a * b that ptxas cannot contract into a following add: it fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (unlike the scalar .rn
// forms, which it treats conservatively), which moves a rounding.  x * y + (+0) rounds exactly like x * y.
__device__ __forceinline__ pk2_t mul2_keep(pk2_t a, pk2_t b) { return fma2(a, b, pk2(0.0f, 0.0f)); }
// a + f (b - a) in both lanes: the same sub + fma as lerpf()
__device__ __forceinline__ pk2_t lerp2(pk2_t a, pk2_t b, pk2_t f) { return fma2(f, sub2(b, a), a); }
__device__ __forceinline__ pk2_t h2pk(unsigned int w) { float2 t = h2f(w); return pk2(t.x, t.y); }
__device__ __forceinline__ float hlo(unsigned int w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }

// Blackwell mixed-precision add (FHADD: f32 = f16 + f32) runs at the full FMA-pipe rate, whereas the fp16 -> fp32
// conversion HADD2.F32 runs at half rate (scripts/microbench/pipes.cu: 3.6 vs 1.9 warp-instr/clk/SM).  So a texel is
// converted with "h + 0" and the lerp's difference is formed as "h1 - t0" directly from the fp16 neighbour: one
// full-rate instruction each, no separate conversion.  Both are exact conversions followed by one fp32 rounding, i.e. the
// same bits as cvt + sub.
__device__ __forceinline__ unsigned short h_lo(unsigned int w) { return (unsigned short)(w & 0xffffu); }
__device__ __forceinline__ unsigned short h_hi(unsigned int w) { return (unsigned short)(w >> 16); }
__device__ __forceinline__ float fh_cvt(unsigned short h) { float d; asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(0.0f)); return d; }
__device__ __forceinline__ float fh_sub(unsigned short h, float c) { float d; asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c)); return d; }
// x-lerp of two channels held as half2 words w0 (texel x0) and w1 (texel x0+1)
__device__ __forceinline__ pk2_t xlerp_h2(unsigned int w0, unsigned int w1, pk2_t fx2)
{
    const float a0 = fh_cvt(h_lo(w0)), a1 = fh_cvt(h_hi(w0));
    return fma2(fx2, pk2(fh_sub(h_lo(w1), a0), fh_sub(h_hi(w1), a1)), pk2(a0, a1));
}

struct FieldVal { pk2_t rg; float b, a; };

// the four x-pairs of a trilinear cell: rows A = (y0,z0), B = (y1,z0), C = (y0,z1), D = (y1,z1), each { T[x0].rg, T[x0].ba,
// T[x0+1].rg, T[x0+1].ba } as fp16 words; idx = cell index in the padded array (DevParams::fRow / fPlane strides)
struct CellRaw { uint4 A, B, C, D; };
template <int LAYOUT>
__device__ __forceinline__ CellRaw load_cell_raw(const uint4 *field, unsigned int fRow, unsigned int fPlane, int idx)
{
    CellRaw c;
    if (LAYOUT == LAYOUT_QUAD) {
        struct alignas(32) Face { uint4 y0, y1; };
        const Face *F = reinterpret_cast<const Face *>(field) + idx;
        ld_u4x2(reinterpret_cast<const uint4 *>(F), c.A, c.B);
        ld_u4x2(reinterpret_cast<const uint4 *>(F + fPlane), c.C, c.D);
    } else {
        const uint4 *F = field + idx;
        c.A = __ldg(F); c.B = __ldg(F + fRow); c.C = __ldg(F + fPlane); c.D = __ldg(F + fPlane + fRow);
    }
    return c;
}

// ---- vector field: trilinear RGBA16F fetch, CLAMP_TO_EDGE (volumeSampler, VV/dataset.cpp:350-357) ----
template <int LAYOUT, bool ALPHA>
__device__ __forceinline__ FieldVal fetch_field_pk(const DevParams &P, float px, float py, float pz)
{
    FieldVal r;
    if (LAYOUT == LAYOUT_PAIR || LAYOUT == LAYOUT_QUAD) {
        int x0, y0, z0;
        float fx, fy, fz;
        axis_clamp_f(px, P.fnf[0], P.fnm1f[0], x0, fx);
        axis_clamp_f(py, P.fnf[1], P.fnm1f[1], y0, fy);
        axis_clamp_f(pz, P.fnf[2], P.fnm1f[2], z0, fz);
        // 32-bit cell indices (volumes up to 2^31 cells); the y / z neighbours are one padded row / plane further (at the
        // clamped edge the replicated texel, weight 0)
        const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz);
        // corner loads A = (y0,z0), B = (y1,z0), C = (y0,z1), D = (y1,z1); each holds texels x0 (.x,.y) and x0+1 (.z,.w)
        const CellRaw cr = load_cell_raw<LAYOUT>(P.field_pair, P.fRow, P.fPlane, z0 * (int)P.fPlane + y0 * (int)P.fRow + x0);
        const uint4 A = cr.A, B = cr.B, C = cr.C, D = cr.D;
        // x-lerp straight from the fp16 pair: t0 + fx (t1 - t0) with t0 = FHADD(h0, 0), t1 - t0 = FHADD(h1, -t0)
        const pk2_t rgA = xlerp_h2(A.x, A.z, fx2), rgB = xlerp_h2(B.x, B.z, fx2);
        const pk2_t rgC = xlerp_h2(C.x, C.z, fx2), rgD = xlerp_h2(D.x, D.z, fx2);
        r.rg = lerp2(lerp2(rgA, rgB, fy2), lerp2(rgC, rgD, fy2), fz2);
        if (ALPHA) {
            const pk2_t baA = xlerp_h2(A.y, A.w, fx2), baB = xlerp_h2(B.y, B.w, fx2);
            const pk2_t baC = xlerp_h2(C.y, C.w, fx2), baD = xlerp_h2(D.y, D.w, fx2);
            const pk2_t ba = lerp2(lerp2(baA, baB, fy2), lerp2(baC, baD, fy2), fz2);
            up2(ba, r.b, r.a);
        } else {
            // blue only: pack the two z planes into one register pair for the x and y lerps
            const float bA0 = fh_cvt(h_lo(A.y)), bC0 = fh_cvt(h_lo(C.y)), bB0 = fh_cvt(h_lo(B.y)), bD0 = fh_cvt(h_lo(D.y));
            const pk2_t bAC = fma2(fx2, pk2(fh_sub(h_lo(A.w), bA0), fh_sub(h_lo(C.w), bC0)), pk2(bA0, bC0));
            const pk2_t bBD = fma2(fx2, pk2(fh_sub(h_lo(B.w), bB0), fh_sub(h_lo(D.w), bD0)), pk2(bB0, bD0));
            const pk2_t by = lerp2(bAC, bBD, fy2);
            r.b = lerpf(lo2(by), hi2(by), fz);
            r.a = 0.0f;
        }
    } else {
        int x0, x1, y0, y1, z0, z1;
        float fx, fy, fz;
        axis_clamp(px, P.fnx, x0, x1, fx);
        axis_clamp(py, P.fny, y0, y1, fy);
        axis_clamp(pz, P.fnz, z0, z1, fz);
        const unsigned int row = (unsigned int)P.fnx;
        const unsigned int b00 = ((unsigned int)z0 * (unsigned int)P.fny + (unsigned int)y0) * row + (unsigned int)x0;
        const unsigned int dy = (unsigned int)(y1 - y0) * row;                       // 0 at the clamped edge
        const unsigned int dz = (unsigned int)(z1 - z0) * row * (unsigned int)P.fny;
        const unsigned int b10 = b00 + dy, b01 = b00 + dz, b11 = b01 + dy;
        const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz);
        const float4 *F = P.field_f4;
        const unsigned int dx = (unsigned int)(x1 - x0);
        const float4 t000 = ld_f4(F + b00), t100 = ld_f4(F + b00 + dx);
        const float4 t010 = ld_f4(F + b10), t110 = ld_f4(F + b10 + dx);
        const float4 t001 = ld_f4(F + b01), t101 = ld_f4(F + b01 + dx);
        const float4 t011 = ld_f4(F + b11), t111 = ld_f4(F + b11 + dx);
        const pk2_t rgA = lerp2(pk2(t000.x, t000.y), pk2(t100.x, t100.y), fx2), rgB = lerp2(pk2(t010.x, t010.y), pk2(t110.x, t110.y), fx2);
        const pk2_t rgC = lerp2(pk2(t001.x, t001.y), pk2(t101.x, t101.y), fx2), rgD = lerp2(pk2(t011.x, t011.y), pk2(t111.x, t111.y), fx2);
        r.rg = lerp2(lerp2(rgA, rgB, fy2), lerp2(rgC, rgD, fy2), fz2);
        const pk2_t baA = lerp2(pk2(t000.z, t000.w), pk2(t100.z, t100.w), fx2), baB = lerp2(pk2(t010.z, t010.w), pk2(t110.z, t110.w), fx2);
        const pk2_t baC = lerp2(pk2(t001.z, t001.w), pk2(t101.z, t101.w), fx2), baD = lerp2(pk2(t011.z, t011.w), pk2(t111.z, t111.w), fx2);
        const pk2_t ba = lerp2(lerp2(baA, baB, fy2), lerp2(baC, baD, fy2), fz2);
        up2(ba, r.b, r.a);
    }
    return r;
}

template <int LAYOUT, bool ALPHA>
__device__ __forceinline__ float4 fetch_field(const DevParams &P, float px, float py, float pz)
{
    const FieldVal v = fetch_field_pk<LAYOUT, ALPHA>(P, px, py, pz);
    float4 r;
    up2(v.rg, r.x, r.y);
    r.z = v.b;
    r.w = v.a;
    return r;
}

// ---- one trilinear cell of the x-pair field, widened once, evaluated at several positions -------------------------
// Heun's predictor position Pos2 = p + d1 and corrector position p + (d1 + d2)/2 differ by (d2 - d1)/2, a small fraction
// of a voxel wherever the field is smooth, so both almost always lie in the same trilinear cell: the 4 LDG.128, the address
// arithmetic and the 24 fp16 -> fp32 widenings (FHADD) are done once per cell and only the lerps are evaluated twice.
// The operations on the texel values are exactly those of fetch_field_pk<LAYOUT_PAIR, false> (same bits).
struct CellCoord { int idx; float fx, fy, fz; };
struct FieldCell {
    pk2_t rg0[4], rgd[4];   // corner rows A = (y0,z0), B = (y1,z0), C = (y0,z1), D = (y1,z1): (r,g) of texel x0, and texel x0+1 minus it
    pk2_t b0[2], bd[2];     // blue with the two z planes packed: [0] = rows (A, C), [1] = rows (B, D)
};

// compile-time options of the walk's coordinate arithmetic (template parameter XF of the sample kernel)
enum { XF_GUARD = 1,     // field coordinates without the CLAMP_TO_EDGE clamp: the layout's guard band replicates the edge texels, so
                         // t0 + f (t1 - t0) with t1 == t0 returns the edge value exactly as the clamped f = 0 does (host guarantees the
                         // band covers every position a walk can reach)
       XF_NSHARE = 2,    // gradient build, noise dimensions == field dimensions and noise_bf stored with the field's guard geometry:
                         // for a position inside [0,1)^3 REPEAT leaves the coordinate untouched, so the noise cell index and weights
                         // ARE the field's (same expression u = s n - 0.5, floor, fraction: same bits).  Used for ray samples whose
                         // whole walk stays inside [0,1)^3 (warp-uniform interior test, DevParams::walkReach); other warps take the
                         // REPEAT rule
       XF_SSHARE = 4 };  // scalar builds, scalar-volume dimensions == field dimensions: the band gate's cell index / weights are the field's

template <bool GUARD>
__device__ __forceinline__ CellCoord field_cell_coord(const DevParams &P, float px, float py, float pz)
{
    CellCoord c;
    int x0, y0, z0;
    if (GUARD) {
        float u = fmaf(px, P.fnf[0], -0.5f);
        x0 = __float2int_rd(u); c.fx = u - __int2float_rn(x0);
        u = fmaf(py, P.fnf[1], -0.5f);
        y0 = __float2int_rd(u); c.fy = u - __int2float_rn(y0);
        u = fmaf(pz, P.fnf[2], -0.5f);
        z0 = __float2int_rd(u); c.fz = u - __int2float_rn(z0);
    } else {
        axis_clamp_f(px, P.fnf[0], P.fnm1f[0], x0, c.fx);
        axis_clamp_f(py, P.fnf[1], P.fnm1f[1], y0, c.fy);
        axis_clamp_f(pz, P.fnf[2], P.fnm1f[2], z0, c.fz);
    }
    c.idx = z0 * (int)P.fPlane + y0 * (int)P.fRow + x0;
    return c;
}

__device__ __forceinline__ void widen_rg(unsigned int w0, unsigned int w1, pk2_t &t0, pk2_t &d)
{
    const float a0 = fh_cvt(h_lo(w0)), a1 = fh_cvt(h_hi(w0));
    t0 = pk2(a0, a1);
    d = pk2(fh_sub(h_lo(w1), a0), fh_sub(h_hi(w1), a1));
}

template <int LAYOUT>
__device__ __forceinline__ FieldCell load_field_cell(const DevParams &P, int idx)
{
    const CellRaw cr = load_cell_raw<LAYOUT>(P.field_pair, P.fRow, P.fPlane, idx);
    const uint4 A = cr.A, B = cr.B, C = cr.C, D = cr.D;
    FieldCell c;
    widen_rg(A.x, A.z, c.rg0[0], c.rgd[0]);
    widen_rg(B.x, B.z, c.rg0[1], c.rgd[1]);
    widen_rg(C.x, C.z, c.rg0[2], c.rgd[2]);
    widen_rg(D.x, D.z, c.rg0[3], c.rgd[3]);
    const float bA0 = fh_cvt(h_lo(A.y)), bC0 = fh_cvt(h_lo(C.y)), bB0 = fh_cvt(h_lo(B.y)), bD0 = fh_cvt(h_lo(D.y));
    c.b0[0] = pk2(bA0, bC0); c.bd[0] = pk2(fh_sub(h_lo(A.w), bA0), fh_sub(h_lo(C.w), bC0));
    c.b0[1] = pk2(bB0, bD0); c.bd[1] = pk2(fh_sub(h_lo(B.w), bB0), fh_sub(h_lo(D.w), bD0));
    return c;
}

__device__ __forceinline__ FieldVal eval_field_cell(const FieldCell &c, float fx, float fy, float fz)
{
    const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz);
    FieldVal r;
    const pk2_t rgA = fma2(fx2, c.rgd[0], c.rg0[0]), rgB = fma2(fx2, c.rgd[1], c.rg0[1]);
    const pk2_t rgC = fma2(fx2, c.rgd[2], c.rg0[2]), rgD = fma2(fx2, c.rgd[3], c.rg0[3]);
    r.rg = lerp2(lerp2(rgA, rgB, fy2), lerp2(rgC, rgD, fy2), fz2);
    const pk2_t bAC = fma2(fx2, c.bd[0], c.b0[0]), bBD = fma2(fx2, c.bd[1], c.b0[1]);
    const pk2_t by = lerp2(bAC, bBD, fy2);
    r.b = lerpf(lo2(by), hi2(by), fz);
    r.a = 0.0f;
    return r;
}

// byte k of w as the float 8388608 + byte (bits 0x4B0000bb): no I2F; the bias cancels in differences
constexpr float kByteBias = 8388608.0f;
__device__ __forceinline__ float byte_biased(unsigned int w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + k)); }
__device__ __forceinline__ float byte_f(unsigned int w, int k) { return byte_biased(w, k) - kByteBias; }
// x-lerp of two biased byte pairs: (a - bias) + f (b - a); b - a is exact on the biased values
__device__ __forceinline__ pk2_t lerp2_biased(pk2_t a, pk2_t b, pk2_t f) { return fma2(f, sub2(b, a), sub2(a, bc2(kByteBias))); }

// trilinear blend of the 8 corner bytes of one cell8 entry, result in [0,1]
__device__ __forceinline__ float cell8_blend(uint2 c, float fx, float fy, float fz)
{
    // register pairs hold the two z planes: (z0, z1)
    const pk2_t fx2 = bc2(fx);
    const pk2_t xa = lerp2_biased(pk2(byte_biased(c.x, 0), byte_biased(c.y, 0)), pk2(byte_biased(c.x, 1), byte_biased(c.y, 1)), fx2);   // y0
    const pk2_t xb = lerp2_biased(pk2(byte_biased(c.x, 2), byte_biased(c.y, 2)), pk2(byte_biased(c.x, 3), byte_biased(c.y, 3)), fx2);   // y1
    const pk2_t y = lerp2(xa, xb, bc2(fy));
    return lerpf(lo2(y), hi2(y), fz) * (1.0f / 255.0f);
}

// scalarSampler: LUMINANCE8, CLAMP_TO_EDGE (VV/dataset.cpp:1025-1038) -> .r
__device__ __forceinline__ float fetch_scalar(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_clamp_f(px, P.snf[0], P.snm1f[0], x0, fx);
    axis_clamp_f(py, P.snf[1], P.snm1f[1], y0, fy);
    axis_clamp_f(pz, P.snf[2], P.snm1f[2], z0, fz);
    uint2 c = ld_u2(P.scalar_cell + (((unsigned int)z0 * (unsigned int)P.sny + (unsigned int)y0) * (unsigned int)P.snx + (unsigned int)x0));
    return cell8_blend(c, fx, fy, fz);
}

// noiseSampler, LUMINANCE8 / alpha channel, REPEAT (VV/dataset.cpp:1283-1335) -> noise value
__device__ __forceinline__ float fetch_noise_scalar(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_repeat_f(px, P.nnf[0], x0, fx);
    axis_repeat_f(py, P.nnf[1], y0, fy);
    axis_repeat_f(pz, P.nnf[2], z0, fz);
    uint2 c = ld_u2(P.noise_cell + (z0 * P.ncPlane + y0 * P.ncRow + x0));   // signed: cell -1 is the wrapped border
    return cell8_blend(c, fx, fy, fz);
}

struct Rgba2 { pk2_t rg, ba; };

// bilinear blend inside one plane of the quad layout: texels (x0y0, x1y0, x0y1, x1y1), channels as (r,g) and (b,a) pairs
__device__ __forceinline__ Rgba2 quad_blend(uint4 q, pk2_t fx2, pk2_t fy2)
{
    Rgba2 r;
    const pk2_t rg0 = lerp2_biased(pk2(byte_biased(q.x, 0), byte_biased(q.x, 1)), pk2(byte_biased(q.y, 0), byte_biased(q.y, 1)), fx2);
    const pk2_t rg1 = lerp2_biased(pk2(byte_biased(q.z, 0), byte_biased(q.z, 1)), pk2(byte_biased(q.w, 0), byte_biased(q.w, 1)), fx2);
    const pk2_t ba0 = lerp2_biased(pk2(byte_biased(q.x, 2), byte_biased(q.x, 3)), pk2(byte_biased(q.y, 2), byte_biased(q.y, 3)), fx2);
    const pk2_t ba1 = lerp2_biased(pk2(byte_biased(q.z, 2), byte_biased(q.z, 3)), pk2(byte_biased(q.w, 2), byte_biased(q.w, 3)), fx2);
    r.rg = lerp2(rg0, rg1, fy2);
    r.ba = lerp2(ba0, ba1, fy2);
    return r;
}

// noiseSampler, RGBA8 (gradient.xyz, noise), REPEAT -> raw texel (freqSamplingGrad, inc_lic.glsl:61-68)
// two bf16 values of one word as an fp32 register pair: a bf16 is the upper half of the fp32 with the same value, so the
// widening is a byte permute / mask on the ALU pipe instead of an FHADD on the (binding) FMA pipe
__device__ __forceinline__ pk2_t bf2pk(unsigned int w)
{
    unsigned int lo, hi;
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(w));
    asm("prmt.b32 %0, %1, 0, 0x3244;" : "=r"(hi) : "r"(w));
    return pk2(__uint_as_float(lo), __uint_as_float(hi));
}

// pre-differenced bf16 layout, cell index idx: x-lerp = t0 + fx * d with d = t1 - t0 stored (exact: |d| <= 255 has 8 significant
// bits); the same fp32 values and operations as the fp16 x-pair path (FHADD forms t1 - t0 exactly as well).
// (An xy-quad form of this layout -- rows y and y+1 in one 32-byte sector, 2 x LDG.256 per tap, like the vector field's -- was
// measured and dropped: cfg3 11.71 -> 12.02 ms, the doubled noise footprint costs more L1 hits than the two loads save;
// profiles/r02/ab16_quad_noise_reload_then_eval.log.)
__device__ __forceinline__ Rgba2 blend_noise_bf(const DevParams &P, int idx, float fx, float fy, float fz)
{
    const uint4 *N = P.noise_bf + idx;
    const uint4 A = ld_u4(N), B = ld_u4(N + P.nbRow);
    const uint4 C = ld_u4(N + P.nbPlane), D = ld_u4(N + P.nbPlane + P.nbRow);
    const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz), k2 = bc2(1.0f / 255.0f);
    Rgba2 r;
    r.rg = mul2(lerp2(lerp2(fma2(fx2, bf2pk(A.z), bf2pk(A.x)), fma2(fx2, bf2pk(B.z), bf2pk(B.x)), fy2),
                      lerp2(fma2(fx2, bf2pk(C.z), bf2pk(C.x)), fma2(fx2, bf2pk(D.z), bf2pk(D.x)), fy2), fz2), k2);
    r.ba = mul2(lerp2(lerp2(fma2(fx2, bf2pk(A.w), bf2pk(A.y)), fma2(fx2, bf2pk(B.w), bf2pk(B.y)), fy2),
                      lerp2(fma2(fx2, bf2pk(C.w), bf2pk(C.y)), fma2(fx2, bf2pk(D.w), bf2pk(D.y)), fy2), fz2), k2);
    return r;
}

// NL: layout known at compile time (2 bf16 {t0, t1 - t0}, 1 fp16 x-pair, 0 u8 xy-quad) or -1 = chosen at run time
// load + trilinear blend of the RGBA noise cell (x0, y0, z0) (cell index in [-1, n-1] per axis) with the weights (fx, fy, fz)
template <int NL = -1>
__device__ __forceinline__ Rgba2 blend_noise_rgba(const DevParams &P, int x0, int y0, int z0, float fx, float fy, float fz)
{
    if (NL == 2 || (NL < 0 && P.noise_bf)) return blend_noise_bf(P, z0 * P.nbPlane + y0 * P.nbRow + x0, fx, fy, fz);
    if (NL == 1 || (NL < 0 && P.noise_pair)) {
        // fp16 x-pair layout (byte values 0..255 are exact in fp16): the same FHADD lerp as the vector field, no byte
        // decode; wrapped border rows / planes, so the neighbours are +npRow / +npPlane for every cell index in [-1, n-1]
        const uint4 *N = P.noise_pair + (z0 * P.npPlane + y0 * P.npRow + x0);
        const uint4 A = ld_u4(N), B = ld_u4(N + P.npRow);
        const uint4 C = ld_u4(N + P.npPlane), D = ld_u4(N + P.npPlane + P.npRow);
        const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz), k2 = bc2(1.0f / 255.0f);
        Rgba2 r;
        r.rg = mul2(lerp2(lerp2(xlerp_h2(A.x, A.z, fx2), xlerp_h2(B.x, B.z, fx2), fy2),
                          lerp2(xlerp_h2(C.x, C.z, fx2), xlerp_h2(D.x, D.z, fx2), fy2), fz2), k2);
        r.ba = mul2(lerp2(lerp2(xlerp_h2(A.y, A.w, fx2), xlerp_h2(B.y, B.w, fx2), fy2),
                          lerp2(xlerp_h2(C.y, C.w, fx2), xlerp_h2(D.y, D.w, fx2), fy2), fz2), k2);
        return r;
    }
    if (x0 < 0) x0 += P.nnx;
    if (y0 < 0) y0 += P.nny;
    if (z0 < 0) z0 += P.nnz;
    int z1 = z0 + 1;
    if (z1 >= P.nnz) z1 = 0;
    const unsigned int plane = (unsigned int)P.nny * (unsigned int)P.nnx;
    const unsigned int i0 = (unsigned int)y0 * (unsigned int)P.nnx + (unsigned int)x0;
    const uint4 a = ld_u4(P.noise_quad + ((unsigned int)z0 * plane + i0));
    const uint4 b = ld_u4(P.noise_quad + ((unsigned int)z1 * plane + i0));
    const pk2_t fx2 = bc2(fx), fy2 = bc2(fy), fz2 = bc2(fz), k2 = bc2(1.0f / 255.0f);
    const Rgba2 p0 = quad_blend(a, fx2, fy2), p1 = quad_blend(b, fx2, fy2);
    Rgba2 r;
    r.rg = mul2(lerp2(p0.rg, p1.rg, fz2), k2);
    r.ba = mul2(lerp2(p0.ba, p1.ba, fz2), k2);
    return r;
}

// noiseSampler, RGBA8 (gradient.xyz, noise), REPEAT -> raw texel (freqSamplingGrad, inc_lic.glsl:61-68)
template <int NL = -1>
__device__ __forceinline__ Rgba2 fetch_noise_rgba_pk(const DevParams &P, float px, float py, float pz)
{
    int x0, y0, z0;
    float fx, fy, fz;
    axis_repeat_f(px, P.nnf[0], x0, fx);
    axis_repeat_f(py, P.nnf[1], y0, fy);
    axis_repeat_f(pz, P.nnf[2], z0, fz);
    return blend_noise_rgba<NL>(P, x0, y0, z0, fx, fy, fz);
}

__device__ __forceinline__ float4 fetch_noise_rgba(const DevParams &P, float px, float py, float pz)
{
    const Rgba2 v = fetch_noise_rgba_pk(P, px, py, pz);
    float4 r;
    up2(v.rg, r.x, r.y);
    up2(v.ba, r.z, r.w);
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
