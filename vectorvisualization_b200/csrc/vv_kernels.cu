// vv_kernels.cu -- sm_100a kernels of the 3D-LIC hot path.
//
//   lic_raycast_kernel   K1  VV/shader/lic3d_fragment.glsl:5-99 + inc_lic.glsl:61-202 + inc_illum.glsl
//   lic_volume_kernel    K2  VV/shader/lic3d_volume_fragment.glsl:2-21
//   volume_raycast_kernel K3 VV/shader/raycast_lic3d_fragment.glsl:5-73
//   unblock_kernel       K5  tile buffer -> row-major frame (+ RGBA8 store, + background_fragment.glsl:9-16)
//
// One thread per ray; a warp owns an 8x4-pixel ray tile and a 256-thread CTA a 16x16-pixel block, so the 32
// rays of a warp sample neighbouring voxels at every march step and their 2x32 Heun streamline walks stay inside
// the same few L1 lines.  CTAs are persistent and pull blocks from an atomic queue (only ~20-70 % of the
// pixels hit the box and chord lengths vary 0..sqrt(3)).  Nothing here is a dense contraction: no tensor cores.
#include "vv_device.cuh"
#include "vv_kernels.h"

// CTA shape of lic_sample_kernel (warps pull work items independently; the CTA only shares the tables).  The register file
// allows 64 registers per thread at 4 x 256 threads per SM and 72 at 7 x 128 (7 instead of 8 warps per scheduler).  The walk
// loop of the gradient build wants ~130: at 64 it spills its four accumulators in every iteration, at 72 it does not --
// measured (profiles/r02/ab14_checkpoints_cta_shapes.log) 2.8 % faster on cfg3; the scalar builds (cfg2, cfg4) are 1.5 % / 3 %
// SLOWER at 7 x 128 and keep 4 x 256 (3 x 256 / 80 registers: cfg2 +5 %, cfg4 +8 %; 5 x 128 / 96 and 4 x 128 / 106
// registers: slower everywhere).  LIC_CTA_THREADS / LIC_MIN_CTAS force one shape for every build (A/B variants).
#if defined(LIC_CTA_THREADS) && defined(LIC_MIN_CTAS)
template <int ILLUM> struct LicShape { static constexpr int kThreads = LIC_CTA_THREADS, kMinCtas = LIC_MIN_CTAS; };
#else
template <int ILLUM> struct LicShape { static constexpr int kThreads = 256, kMinCtas = 4; };
template <> struct LicShape<1 /* ILLUM_GRADIENT */> { static constexpr int kThreads = 128, kMinCtas = 7; };
#endif

namespace vvb200 {

// ------------------------------------------------------------------------------------------------
// ray set-up: VV/shader/lic3d_fragment.glsl:12-21; the fragment position is the entry point of the pixel
// ray into [0,extent]^3 (front faces of the proxy cube, VV/renderer.cpp:682-736).  Evaluated with explicit
// round-to-nearest double intrinsics (no FMA contraction) so the host oracle computes the same bits.
__device__ __forceinline__ bool pixel_ray(const DevParams &P, int px, int py, float e[3])
{
    double ex = __dmul_rn(__dmul_rn(__dsub_rn(__ddiv_rn(__dmul_rn(2.0, (double)px + 0.5), (double)P.width), 1.0), P.tanHalf), P.aspect);
    double ey = __dmul_rn(__dsub_rn(__ddiv_rn(__dmul_rn(2.0, (double)py + 0.5), (double)P.height), 1.0), P.tanHalf);
    double ez = -1.0;
    double d[3], o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d[i] = __dadd_rn(__dadd_rn(__dmul_rn(P.rot[i], ex), __dmul_rn(P.rot[3 + i], ey)), __dmul_rn(P.rot[6 + i], ez));
        o[i] = P.camD[i];
    }
    double tn = -1e300, tf = 1e300, faceval = 0.0;
    int face = -1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double lo = 0.0, hi = P.extent[i];
        if (d[i] == 0.0) {
            if (o[i] < lo || o[i] > hi) return false;
            continue;
        }
        double t0 = __ddiv_rn(__dsub_rn(lo, o[i]), d[i]), t1 = __ddiv_rn(__dsub_rn(hi, o[i]), d[i]);
        double fv = lo;
        if (t0 > t1) { double t = t0; t0 = t1; t1 = t; fv = hi; }
        if (t0 > tn) { tn = t0; face = i; faceval = fv; }
        if (t1 < tf) tf = t1;
    }
    bool have = (tn < tf) && tn >= P.nearD && tn <= P.farD && face >= 0;   // view-volume clipping of the front faces
    double p[3] = {0.0, 0.0, 0.0};
    if (have) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            p[i] = __dadd_rn(o[i], __dmul_rn(tn, d[i]));
            if (i == face) p[i] = faceval;
            p[i] = fmin(fmax(p[i], 0.0), P.extent[i]);
        }
    }
    if (P.nClip > 0) {
        // User clip planes (Renderer::enableClipPlanes / drawClippedPolygon, VV/renderer.cpp:156-163, 1294-1309): GL clips
        // the cube's front faces; the cap polygons n^.q = -(d - 0.0001) are drawn afterwards in plane order, culled when
        // they face away, clipped by the other planes; without depth test or blending the last fragment under the pixel
        // wins.  Same double arithmetic, same order as the oracle.
        auto kept = [&](const double q[3], int skip) {
            for (int j = 0; j < P.nClip; ++j) {
                if (j == skip) continue;
                const double *c = P.clipEq[j];
                if (__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[0], q[0]), __dmul_rn(c[1], q[1])), __dmul_rn(c[2], q[2])), c[3]) < 0.0) return false;
            }
            return true;
        };
        if (have) {
            const double q[3] = {__dsub_rn(p[0], P.centerD[0]), __dsub_rn(p[1], P.centerD[1]), __dsub_rn(p[2], P.centerD[2])};
            have = kept(q, -1);
        }
        const double oc[3] = {__dsub_rn(o[0], P.centerD[0]), __dsub_rn(o[1], P.centerD[1]), __dsub_rn(o[2], P.centerD[2])};
        for (int i = 0; i < P.nClip; ++i) {
            const double *n = P.clipN[i];
            const double dist = P.clipDist[i];
            if (!(dist == dist)) continue;                                    // degenerate normal: no cap
            const double dn = __dadd_rn(__dadd_rn(__dmul_rn(n[0], d[0]), __dmul_rn(n[1], d[1])), __dmul_rn(n[2], d[2]));
            if (!(dn > 0.0)) continue;                                        // cap faces away from the viewer: culled
            const double on = __dadd_rn(__dadd_rn(__dmul_rn(n[0], oc[0]), __dmul_rn(n[1], oc[1])), __dmul_rn(n[2], oc[2]));
            const double t = __ddiv_rn(__dsub_rn(dist, on), dn);
            if (!(t >= P.nearD && t <= P.farD)) continue;
            const double q[3] = {__dadd_rn(oc[0], __dmul_rn(t, d[0])), __dadd_rn(oc[1], __dmul_rn(t, d[1])), __dadd_rn(oc[2], __dmul_rn(t, d[2]))};
            bool inside = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (fabs(q[k]) > __dmul_rn(0.5, P.extent[k])) inside = false;
            if (!inside || !kept(q, i)) continue;
            have = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) p[k] = fmin(fmax(__dadd_rn(q[k], __dmul_rn(0.5, P.extent[k])), 0.0), P.extent[k]);
        }
    }
    if (!have) return false;
#pragma unroll
    for (int i = 0; i < 3; ++i) e[i] = (float)p[i];
    return true;
}

// pos += dir * stepSize * texture2DRect(mcOffsetSampler, gl_FragCoord.xy).r  (lic3d_fragment.glsl:31-33, USE_MC_OFFSET)
__device__ __forceinline__ void mc_offset(const DevParams &P, int px, int py, f3 &pos, f3 dstep)
{
    if (!P.mcOffsets) return;
    const float r = P.mcOffsets[(size_t)py * P.width + px];
    pos.x = __fadd_rn(pos.x, __fmul_rn(dstep.x, r));
    pos.y = __fadd_rn(pos.y, __fmul_rn(dstep.y, r));
    pos.z = __fadd_rn(pos.z, __fmul_rn(dstep.z, r));
}

// GLSL normalize() as v / sqrt(dot(v,v)), IEEE round-to-nearest, uncontracted
__device__ __forceinline__ f3 normalize_rn(f3 v)
{
    float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
    return mk3(__fdiv_rn(v.x, l), __fdiv_rn(v.y, l), __fdiv_rn(v.z, l));
}
__device__ __forceinline__ f3 normalize3(f3 v)
{
    float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return mk3(v.x / l, v.y / l, v.z / l);
}
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// ------------------------------------------------------------------------------------------------
// inc_lic.glsl

// freqSampling, inc_lic.glsl:71-90 (scalar band gate + noise(pos * gradient.z).a)
template <bool NGATE>
__device__ __forceinline__ float noise_tap(const DevParams &P, f3 q)
{
    if (NGATE) {
        float s = fetch_scalar(P, q.x, q.y, q.z);
        if (!(s > 0.1f && s < 0.3f)) return 0.0f;
    }
    if (P.quirkLumAlpha) return 1.0f;   // Q7: GL_LUMINANCE noise read through .a
    return fetch_noise_scalar(P, q.x * P.freq, q.y * P.freq, q.z * P.freq);
}

// streamline walker: position (x,y packed, z) and the field sample at it (r,g packed, b, a)
struct Walker { pk2_t qxy; float qz; pk2_t vrg; float vb, va; CellCoord c; };   // c: field cell of the position (heun_step, XF sharing)

__device__ __forceinline__ Walker make_walker(f3 pos, float4 centre)
{
    Walker w;
    w.qxy = pk2(pos.x, pos.y); w.qz = pos.z;
    w.vrg = pk2(centre.x, centre.y); w.vb = centre.z; w.va = centre.w;
    return w;
}

#ifndef VV_SEQ_WALKS
#define VV_SEQ_WALKS 0    // experiment: 1 = backward walk, then forward walk (one fetch chain per thread, fewer live registers)
#endif
#ifndef VV_CELL_REUSE
#define VV_CELL_REUSE 1   // one cell load per Heun step, evaluated at the predictor and the corrector position (see FieldCell)
#endif

// one Heun step of singleLICstep (inc_lic.glsl:104-128); sh = dir * h
template <int LAYOUT, bool SOF, int XF = 0>
__device__ __forceinline__ void heun_step(const DevParams &P, Walker &w, float sh, float *dbg = nullptr)
{
    const float s1 = SOF ? w.va * sh : sh;    // licdir *= step.a (SPEED_OF_FLOW) then *= h
    const pk2_t s2 = bc2(s1), two = bc2(2.0f), mone = bc2(-1.0f);
    const pk2_t d1 = mul2_keep(fma2(two, w.vrg, mone), s2);               // licdir = (2 v - 1) * dir * h
    // (scalar z lane with explicit round-to-nearest intrinsics: nvcc must not contract the product into the position sums, the
    // packed x / y lanes are not contracted either -- every lane rounds exactly where the shader's statement sequence does)
    const float d1z = __fmul_rn(fmaf(2.0f, w.vb, -1.0f), s1);
    const float p2z = __fadd_rn(w.qz, d1z);
    const pk2_t p2 = add2(w.qxy, d1);                                      // Pos2 = newPos + licdir
    if constexpr (LAYOUT != LAYOUT_F4 && !SOF && VV_CELL_REUSE) {
        constexpr bool GUARD = (XF & XF_GUARD) != 0;
        const CellCoord c2 = field_cell_coord<GUARD>(P, lo2(p2), hi2(p2), p2z);
        const FieldCell cell = load_field_cell<LAYOUT>(P, c2.idx);
        const FieldVal v2 = eval_field_cell(cell, c2.fx, c2.fy, c2.fz);
        if (dbg) { dbg[0] = lo2(p2); dbg[1] = hi2(p2); dbg[2] = p2z; dbg[3] = lo2(v2.rg); dbg[4] = hi2(v2.rg); dbg[5] = v2.b; }   // diagnostic only
        const pk2_t d2 = mul2_keep(fma2(two, v2.rg, mone), s2);
        const float d2z = __fmul_rn(fmaf(2.0f, v2.b, -1.0f), s1);
        w.qxy = fma2(bc2(0.5f), add2(d1, d2), w.qxy);                      // newPos += 0.5 (licdir + licdir2)
        w.qz = fmaf(0.5f, __fadd_rn(d1z, d2z), w.qz);
        const CellCoord c = field_cell_coord<GUARD>(P, lo2(w.qxy), hi2(w.qxy), w.qz);
        // (measured alternative: the reload as a conditional region in front of ONE copy of the evaluation -- cfg3 +2 %, cfg2 +2.5 %,
        // cfg4 -1.7 %; profiles/r02/ab16_quad_noise_reload_then_eval.log)
        FieldVal v;
        if (c.idx == c2.idx) v = eval_field_cell(cell, c.fx, c.fy, c.fz);
        else v = eval_field_cell(load_field_cell<LAYOUT>(P, c.idx), c.fx, c.fy, c.fz);   // the corrector left the predictor's cell (rare)
        w.vrg = v.rg; w.vb = v.b; w.va = v.a;
        w.c = c;
    } else {
        static_assert(XF == 0 || (LAYOUT != LAYOUT_F4 && !SOF && VV_CELL_REUSE), "the coordinate fast paths live in the cell-reuse step");
        const FieldVal v2 = fetch_field_pk<LAYOUT, false>(P, lo2(p2), hi2(p2), p2z);
        const pk2_t d2 = mul2_keep(fma2(two, v2.rg, mone), s2);
        const float d2z = __fmul_rn(fmaf(2.0f, v2.b, -1.0f), s1);
        w.qxy = fma2(bc2(0.5f), add2(d1, d2), w.qxy);                      // newPos += 0.5 (licdir + licdir2)
        w.qz = fmaf(0.5f, __fadd_rn(d1z, d2z), w.qz);
        const FieldVal v = fetch_field_pk<LAYOUT, SOF>(P, lo2(w.qxy), hi2(w.qxy), w.qz);
        w.vrg = v.rg; w.vb = v.b; w.va = v.a;
    }
}

// (Measured and dropped: one Heun step of BOTH walkers with their predictor cells requested together -- pinned loads in front of
// the first walker's divergent reload region, which ptxas does not hoist the second walker's loads across, so that the two walks'
// memory round trips overlap instead of running one after the other.  Same arithmetic, same frames; the 32 registers of raw cells in
// flight push the loop into spills (gradient build 5 + 5, scalar builds 11 + 11 local accesses per double step) and the kernel is
// throughput-, not latency-limited: cfg3 +2.6 %, cfg2 +2.1 %, cfg4 +7.5 % slower; profiles/r02/ab23_paired_walker_loads.log.)

// computeLIC, inc_lic.glsl:152-202, scalar build.  The backward and forward walks are independent; they are
// advanced in the same loop iteration so that each thread keeps two dependent fetch chains in flight.
// nBwdEff / nFwdEff: the walk stops after the last step whose filter-kernel weight is non-zero (trailing zero
// weights contribute exactly 0 to the sum).
// STRAIGHT: the common part of the two walks as straight-line code (measured: 3 % faster in lic_sample_kernel on the
// scalar builds, neutral on the gradient build, 8 % slower in lic_volume_kernel, which therefore keeps the guarded loop)
template <int LAYOUT, bool NGATE, bool SOF, bool STRAIGHT = true, int XF = 0>
__device__ __forceinline__ float compute_lic_scalar(const DevParams &P, const float *s_kw, f3 pos, float4 centre)
{
    float acc0 = noise_tap<NGATE>(P, pos) * s_kw[0];
    float accB = 0.0f, accF = 0.0f;
    Walker wb = make_walker(pos, centre), wf = wb;
    const float *kwB = s_kw + 1, *kwF = s_kw + 1 + P.nBwd;
    const int nB = P.nBwdEff, nF = P.nFwdEff;
    if (STRAIGHT) {
        // common part of the two walks without per-direction guards, then the tail of the longer walk
        const int nMin = VV_SEQ_WALKS ? 0 : min(nB, nF);
        int k = 0;
        for (; k < nMin; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
            heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
            accB = fmaf(noise_tap<NGATE>(P, mk3(lo2(wb.qxy), hi2(wb.qxy), wb.qz)), kwB[k], accB);
            accF = fmaf(noise_tap<NGATE>(P, mk3(lo2(wf.qxy), hi2(wf.qxy), wf.qz)), kwF[k], accF);
        }
        for (; k < nB; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
            accB = fmaf(noise_tap<NGATE>(P, mk3(lo2(wb.qxy), hi2(wb.qxy), wb.qz)), kwB[k], accB);
        }
        for (k = nMin; k < nF; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
            accF = fmaf(noise_tap<NGATE>(P, mk3(lo2(wf.qxy), hi2(wf.qxy), wf.qz)), kwF[k], accF);
        }
    } else {
        const int n = max(nB, nF);
        for (int k = 0; k < n; ++k) {
            if (k < nB) {
                heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
                accB = fmaf(noise_tap<NGATE>(P, mk3(lo2(wb.qxy), hi2(wb.qxy), wb.qz)), kwB[k], accB);
            }
            if (k < nF) {
                heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
                accF = fmaf(noise_tap<NGATE>(P, mk3(lo2(wf.qxy), hi2(wf.qxy), wf.qz)), kwF[k], accF);
            }
        }
    }
    return (acc0 + accB) + accF;
}

// the RGBA-noise tap at a walker's position; fast: the walk stays inside [0,1)^3 and the noise shares the field's cell (XF_NSHARE)
template <int NL, int XF>
__device__ __forceinline__ Rgba2 noise_tap_rgba(const DevParams &P, const Walker &w, bool fast)
{
    if constexpr ((XF & XF_NSHARE) != 0) {
        static_assert(NL == 2, "the shared cell index addresses the bf16 layout");
        // only the coordinates differ between the two cases; one copy of the loads and the blend
        int idx = w.c.idx;
        float fx = w.c.fx, fy = w.c.fy, fz = w.c.fz;
        if (!fast) {
            int x0, y0, z0;
            axis_repeat_f(lo2(w.qxy), P.nnf[0], x0, fx);
            axis_repeat_f(hi2(w.qxy), P.nnf[1], y0, fy);
            axis_repeat_f(w.qz, P.nnf[2], z0, fz);
            idx = z0 * P.nbPlane + y0 * P.nbRow + x0;
        }
        return blend_noise_bf(P, idx, fx, fy, fz);
    } else {
        return fetch_noise_rgba_pk<NL>(P, lo2(w.qxy), hi2(w.qxy), w.qz);
    }
}

// computeLIC, USE_NOISE_GRADIENTS build: vec4 accumulation of raw RGBA noise texels (Q8)
// the two walks of compute_lic_grad; FAST (XF_NSHARE only): every tap takes the field's cell -- a separate instantiation of the
// loops, so that the hot one carries neither the REPEAT arithmetic nor the branch around it
template <int LAYOUT, bool SOF, bool STRAIGHT, int NL, int XF, bool FAST>
__device__ __forceinline__ void lic_grad_walks(const DevParams &P, const float *s_kw, f3 pos, float4 centre,
                                               pk2_t &accBrg, pk2_t &accBba, pk2_t &accFrg, pk2_t &accFba)
{
    Walker wb = make_walker(pos, centre), wf = wb;
    const float *kwB = s_kw + 1, *kwF = s_kw + 1 + P.nBwd;
    const int nB = P.nBwdEff, nF = P.nFwdEff;
    if (STRAIGHT) {
        const int nMin = VV_SEQ_WALKS ? 0 : min(nB, nF);
        int k = 0;
        for (; k < nMin; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
            heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
            const Rgba2 tb = noise_tap_rgba<NL, XF>(P, wb, FAST);
            const Rgba2 tf = noise_tap_rgba<NL, XF>(P, wf, FAST);
            const pk2_t wB = bc2(kwB[k]), wF = bc2(kwF[k]);
            accBrg = fma2(tb.rg, wB, accBrg);
            accBba = fma2(tb.ba, wB, accBba);
            accFrg = fma2(tf.rg, wF, accFrg);
            accFba = fma2(tf.ba, wF, accFba);
        }
        for (; k < nB; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
            const Rgba2 t = noise_tap_rgba<NL, XF>(P, wb, FAST);
            const pk2_t w = bc2(kwB[k]);
            accBrg = fma2(t.rg, w, accBrg);
            accBba = fma2(t.ba, w, accBba);
        }
        for (k = nMin; k < nF; ++k) {
            heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
            const Rgba2 t = noise_tap_rgba<NL, XF>(P, wf, FAST);
            const pk2_t w = bc2(kwF[k]);
            accFrg = fma2(t.rg, w, accFrg);
            accFba = fma2(t.ba, w, accFba);
        }
    } else {
        const int n = max(nB, nF);
        for (int k = 0; k < n; ++k) {
            if (k < nB) {
                heun_step<LAYOUT, SOF, XF>(P, wb, -P.h);
                const Rgba2 t = noise_tap_rgba<NL, XF>(P, wb, FAST);
                const pk2_t w = bc2(kwB[k]);
                accBrg = fma2(t.rg, w, accBrg);
                accBba = fma2(t.ba, w, accBba);
            }
            if (k < nF) {
                heun_step<LAYOUT, SOF, XF>(P, wf, P.h);
                const Rgba2 t = noise_tap_rgba<NL, XF>(P, wf, FAST);
                const pk2_t w = bc2(kwF[k]);
                accFrg = fma2(t.rg, w, accFrg);
                accFba = fma2(t.ba, w, accFba);
            }
        }
    }
}

template <int LAYOUT, bool SOF, bool STRAIGHT = true, int NL = -1, int XF = 0>
__device__ __forceinline__ float4 compute_lic_grad(const DevParams &P, const float *s_kw, f3 pos, float4 centre)
{
    const Rgba2 c = fetch_noise_rgba_pk<NL>(P, pos.x, pos.y, pos.z);
    const pk2_t w0 = bc2(s_kw[0]);
    pk2_t accBrg = pk2(0.f, 0.f), accBba = accBrg, accFrg = accBrg, accFba = accBrg;
    if constexpr ((XF & XF_NSHARE) != 0) {
        // warp-uniform: every active lane's walk stays inside [0,1)^3 (it starts at least (S + 2) h away from the faces)
        const float lo = fminf(fminf(pos.x, pos.y), pos.z), hi = fmaxf(fmaxf(pos.x, pos.y), pos.z);
        const bool fast = __all_sync(__activemask(), lo >= P.walkReach && hi < 1.0f - P.walkReach) != 0;
        if (fast) lic_grad_walks<LAYOUT, SOF, STRAIGHT, NL, XF, true>(P, s_kw, pos, centre, accBrg, accBba, accFrg, accFba);
        else lic_grad_walks<LAYOUT, SOF, STRAIGHT, NL, XF, false>(P, s_kw, pos, centre, accBrg, accBba, accFrg, accFba);
    } else {
        lic_grad_walks<LAYOUT, SOF, STRAIGHT, NL, XF, false>(P, s_kw, pos, centre, accBrg, accBba, accFrg, accFba);
    }
    const pk2_t rg = add2(fma2(c.rg, w0, accBrg), accFrg);
    const pk2_t ba = add2(fma2(c.ba, w0, accBba), accFba);
    float4 r;
    up2(rg, r.x, r.y);
    up2(ba, r.z, r.w);
    return r;
}

// ------------------------------------------------------------------------------------------------
// inc_illum.glsl

// color.a = 1 - pow(1 - color.a, alphaCorrection), inc_illum.glsl:38,169
__device__ __forceinline__ float opacity_correct(const DevParams &P, float a) { return 1.0f - powf(1.0f - a, P.alphaCorr); }

// illumLIC, inc_illum.glsl:158-172
__device__ __forceinline__ float4 illum_lic(const DevParams &P, const float *s_opac, float illum, float4 tf)
{
    float4 c;
    c.x = illum * tf.x * P.illumScale;
    c.y = illum * tf.y * P.illumScale;
    c.z = illum * tf.z * P.illumScale;
    c.w = opacity_correct(P, opac_lookup(s_opac, illum * 1.3f) * tf.w);
    return c;
}

// illumGradient, inc_illum.glsl:1-41 (LIGHT0: ambient 0, diffuse 1, specular 1; exponent VV/3DLIC.cpp:736)
__device__ __forceinline__ float4 illum_gradient(const DevParams &P, const float *s_opac, float4 illum, float4 tf, f3 pos, f3 dir)
{
    f3 L = normalize3(mk3(P.lightPos[0] - pos.x * P.scaleVolInv[0], P.lightPos[1] - pos.y * P.scaleVolInv[1],
                          P.lightPos[2] - pos.z * P.scaleVolInv[2]));
    f3 V = normalize3(mk3(-dir.x, -dir.y, -dir.z));
    f3 N = normalize3(mk3(-illum.x, -illum.y, -illum.z));
    float ln = dot3(L, N);
    f3 R = normalize3(mk3(2.0f * ln * N.x - L.x, 2.0f * ln * N.y - L.y, 2.0f * ln * N.z - L.z));
    float spec = powf(clamp01(dot3(R, V)), P.specExp) * illum.w;
    float diff = clamp01(ln) * P.illumScale;
    float k = diff + 0.3f;
    float4 c;
    c.x = (tf.x * illum.w * k + spec) * P.illumScale;
    c.y = (tf.y * illum.w * k + spec) * P.illumScale;
    c.z = (tf.z * illum.w * k + spec) * P.illumScale;
    c.w = opacity_correct(P, opac_lookup(s_opac, illum.w * 1.3f) * tf.w);
    return c;
}

__device__ __forceinline__ void illum2d_lookup(const DevParams &P, int which, int ch, float sx, float sy, float *out)
{
    int x0, x1, y0, y1;
    float fx, fy;
    axis_clamp(sx, P.illum_w, x0, x1, fx);
    axis_clamp(sy, P.illum_h, y0, y1, fy);
    const float *T = P.illum2d[which];
    for (int k = 0; k < ch; ++k) {
        float a = lerpf(__ldg(T + (y0 * P.illum_w + x0) * ch + k), __ldg(T + (y0 * P.illum_w + x1) * ch + k), fx);
        float b = lerpf(__ldg(T + (y1 * P.illum_w + x0) * ch + k), __ldg(T + (y1 * P.illum_w + x1) * ch + k), fx);
        out[k] = lerpf(a, b, fy);
    }
}

__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// illumMallo, inc_illum.glsl:46-112
__device__ __forceinline__ float4 illum_mallo(const DevParams &P, const float *s_opac, float illum, float4 tf, f3 pos, f3 dir, f3 tangent)
{
    f3 L = normalize3(mk3(P.lightPos[0] - pos.x * P.scaleVolInv[0], P.lightPos[1] - pos.y * P.scaleVolInv[1],
                          P.lightPos[2] - pos.z * P.scaleVolInv[2]));
    f3 V = normalize3(mk3(-dir.x, -dir.y, -dir.z));
    f3 T = normalize3(mk3(2.0f * tangent.x - 1.0f, 2.0f * tangent.y - 1.0f, 2.0f * tangent.z - 1.0f));
    f3 B = normalize3(cross3(T, V));
    f3 N = cross3(B, T);
    f3 H = normalize3(mk3(V.x + L.x, V.y + L.y, V.z + L.z));
    float ltx = dot3(L, N), lty = dot3(L, T), ltz = dot3(H, N), ltw = dot3(H, T);
    float tmpx = 1.0f / sqrtf(1.0f - lty * lty);
    float tmpy = 1.0f / sqrtf(1.0f - ltw * ltw);
    float nz = ltx * tmpx, nw = ltz * tmpy;     // lt.zw = lt.xz * tmp
    float cy = 0.5f * lty + 0.5f, cz = 0.5f * nz + 0.5f, cw = 0.5f * nw + 0.5f;
    float d1, s1;
    illum2d_lookup(P, 1, 1, cz, cy, &d1);
    illum2d_lookup(P, 2, 1, cz, cw, &s1);
    float specular = clamp01(s1 * powf(tmpy, -P.specExp));
    float diffuse = d1 * P.illumScale;
    float4 c;
    c.x = tf.x * illum * diffuse + specular;
    c.y = tf.y * illum * diffuse + specular;
    c.z = tf.z * illum * diffuse + specular;
    c.w = opacity_correct(P, opac_lookup(s_opac, illum * 1.3f) * tf.w);
    return c;
}

// illumZoeckler, inc_illum.glsl:117-154 (Q17: .rg of a LUMINANCE_ALPHA texture = (L, L))
__device__ __forceinline__ float4 illum_zoeckler(const DevParams &P, const float *s_opac, float illum, float4 tf, f3 pos, f3 dir, f3 tangent)
{
    f3 L = normalize3(mk3(P.lightPos[0] - pos.x * P.scaleVolInv[0], P.lightPos[1] - pos.y * P.scaleVolInv[1],
                          P.lightPos[2] - pos.z * P.scaleVolInv[2]));
    f3 V = normalize3(dir);
    f3 T = normalize3(mk3(2.0f * tangent.x - 1.0f, 2.0f * tangent.y - 1.0f, 2.0f * tangent.z - 1.0f));
    float la[2];
    illum2d_lookup(P, 0, 2, 0.5f * dot3(L, T) + 0.5f, 0.5f * dot3(V, T) + 0.5f, la);
    float sr = la[0] * P.illumScale, sg = la[0];
    float4 c;
    c.x = tf.x * illum * sr + 0.9f * sg;
    c.y = tf.y * illum * sr + 0.9f * sg;
    c.z = tf.z * illum * sr + 0.9f * sg;
    c.w = opacity_correct(P, opac_lookup(s_opac, illum * 1.3f) * tf.w);
    return c;
}

__device__ __forceinline__ float tf_index(const DevParams &P, float4 v, float scalar)
{
    switch (P.tfMode) {
    case TF_A: return v.w;
    case TF_R: return v.x;
    case TF_LENGTH: return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    case TF_SCALAR: return scalar;
    default: return v.z;
    }
}

// ------------------------------------------------------------------------------------------------
struct SharedTables {
    float4 tf[256];
    float opac[256];
    int block;
    float kw[1];      // 1 + nBwd + nFwd weights; the dynamic shared-memory size is table_bytes(P)
};

// shared memory actually needed: the smaller it is, the more of the 256 KB L1/shared array is left to the L1 cache
static size_t table_bytes(const DevParams &P) { return offsetof(SharedTables, kw) + sizeof(float) * (size_t)(1 + P.nBwd + P.nFwd); }

__device__ __forceinline__ void load_tables(const DevParams &P, SharedTables &S)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        S.tf[i] = P.tf_rgba[i];
        S.opac[i] = P.tf_opac[i];
    }
    const int nk = 1 + P.nBwd + P.nFwd;
    for (int i = threadIdx.x; i < nk; i += blockDim.x) S.kw[i] = P.kw[i];
    __syncthreads();
}

// one ray sample of lic3d_fragment.glsl:44-81: vector fetch, TF, gate, computeLIC, illumination.
// Returns false when the LIC gate skips the sample (src keeps its previous value in the shader).
template <int LAYOUT, int ILLUM, bool NGATE, bool SOF, int NL = -1, int XF = 0>
__device__ __forceinline__ bool shade_sample(const DevParams &P, const SharedTables &S, f3 pos, f3 dir, float4 &src)
{
    constexpr bool GRAD = (ILLUM == ILLUM_GRADIENT);   // ILLUM_GRADIENT => USE_NOISE_GRADIENTS, inc_header.glsl:17-19
    float4 vd = fetch_field<LAYOUT, true>(P, pos.x, pos.y, pos.z);                  // :44
    float sc = 0.0f;
    if (P.tfMode == TF_SCALAR) sc = fetch_scalar(P, pos.x, pos.y, pos.z);            // :52
    float4 tf = tf_lookup(S.tf, tf_index(P, vd, sc));                               // :54
    // gate :59-61 (scalarData.g > -0.0001 is always true for a LUMINANCE8 texture)
    if (P.gateMode == GATE_TF_ALPHA && !(tf.w > 0.05f)) return false;
    if (GRAD) {
        float4 il = compute_lic_grad<LAYOUT, SOF, true, NL, XF>(P, S.kw, pos, vd);                // :64
        il.w *= P.licScale;                                                         // :67
        src = illum_gradient(P, S.opac, il, tf, pos, dir);
    } else {
        float il = compute_lic_scalar<LAYOUT, NGATE, SOF, true, XF>(P, S.kw, pos, vd) * P.licScale;
        if (ILLUM == ILLUM_MALLO) src = illum_mallo(P, S.opac, il, tf, pos, dir, mk3(vd.x, vd.y, vd.z));
        else if (ILLUM == ILLUM_ZOECKLER) src = illum_zoeckler(P, S.opac, il, tf, pos, dir, mk3(vd.x, vd.y, vd.z));
        else src = illum_lic(P, S.opac, il, tf);
    }
    return true;
}

// GL float -> UNORM8 -> float of the RGBA8 back buffer (GL 2.1 spec 2.14.9), uncontracted like the oracle's
__device__ __forceinline__ float unorm8_rn(float v) { return floorf(__fadd_rn(__fmul_rn(clamp01(v), 255.0f), 0.5f)); }
// glBlendFunc(GL_ONE_MINUS_DST_ALPHA, GL_ONE) in an RGBA8 frame buffer: dst holds the byte values 0..255 as floats
__device__ __forceinline__ void blend8(float4 &dst, float4 src)
{
    const float k = __fsub_rn(1.0f, __fdiv_rn(dst.w, 255.0f));
    dst.x = unorm8_rn(__fadd_rn(__fmul_rn(src.x, k), __fdiv_rn(dst.x, 255.0f)));
    dst.y = unorm8_rn(__fadd_rn(__fmul_rn(src.y, k), __fdiv_rn(dst.y, 255.0f)));
    dst.z = unorm8_rn(__fadd_rn(__fmul_rn(src.z, k), __fdiv_rn(dst.z, 255.0f)));
    dst.w = unorm8_rn(__fadd_rn(__fmul_rn(src.w, k), __fdiv_rn(dst.w, 255.0f)));
}
__device__ __forceinline__ float half_round(float v) { return __half2float(__float2half_rn(v)); }

// lic3d_fragment.glsl:83-84: src.rgb *= src.a; dest = clamp((1 - dest.a) src + dest, 0, 1)
__device__ __forceinline__ void composite(float4 &dest, float4 src)
{
    const float k = 1.0f - dest.w;
    dest.x = clamp01(fmaf(k, src.x * src.w, dest.x));
    dest.y = clamp01(fmaf(k, src.y * src.w, dest.y));
    dest.z = clamp01(fmaf(k, src.z * src.w, dest.z));
    dest.w = clamp01(fmaf(k, src.w, dest.w));
}

// ray set-up in texture space (lic3d_fragment.glsl:12-21), uncontracted fp32
__device__ __forceinline__ void ray_setup(const DevParams &P, const float e[3], f3 &pos, f3 &dir, f3 &dstep)
{
    pos = mk3(__fmul_rn(e[0], P.scaleVol[0]), __fmul_rn(e[1], P.scaleVol[1]), __fmul_rn(e[2], P.scaleVol[2]));
    f3 gd = normalize_rn(mk3(__fsub_rn(e[0], P.camera[0]), __fsub_rn(e[1], P.camera[1]), __fsub_rn(e[2], P.camera[2])));
    dir = mk3(__fmul_rn(gd.x, P.scaleVol[0]), __fmul_rn(gd.y, P.scaleVol[1]), __fmul_rn(gd.z, P.scaleVol[2]));
    dstep = mk3(__fmul_rn(dir.x, P.stepSize), __fmul_rn(dir.y, P.stepSize), __fmul_rn(dir.z, P.stepSize));
}
__device__ __forceinline__ bool outside_box(const DevParams &P, f3 pos)
{
    // any(clamp(pos, 0, texMax) - pos), lic3d_fragment.glsl:91
    return pos.x < 0.0f || pos.x > P.texMax[0] || pos.y < 0.0f || pos.y > P.texMax[1] || pos.z < 0.0f || pos.z > P.texMax[2];
}

// K1 (per-ray variant) ------------------------------------------------------------------------------------
// One thread marches one ray front to back.  Kept as the literal form of the fragment shader and as a cross-check
// of the sample-parallel pipeline below (VV_OPT_RAYCAST_MODE = 0); it under-fills the GPU for small images.
template <int LAYOUT, int ILLUM, bool NGATE, bool SOF>
__global__ void __launch_bounds__(256) lic_raycast_kernel(const __grid_constant__ DevParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &S = *reinterpret_cast<SharedTables *>(smem_raw);
    load_tables(P, S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp & 1) * 8 + (lane & 7);
    const int ly = (warp >> 1) * 4 + (lane >> 3);

    for (;;) {
        if (threadIdx.x == 0) S.block = (int)atomicAdd(P.blockCounter, 1u);
        __syncthreads();
        const int lb = S.block;
        __syncthreads();
        if (lb >= P.nLocalBlocks) break;
        const int b = P.rank + lb * P.world;
        int bx, by;
        block_xy(P, b, bx, by);
        const int px = bx * kBlockDim + lx;
        const int py = by * kBlockDim + ly;

        float4 dest = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int nsamples = 0;
        float e[3];
        if (px < P.width && py < P.height && pixel_ray(P, px, py, e)) {
            f3 pos, dir, dstep;
            ray_setup(P, e, pos, dir, dstep);
            mc_offset(P, px, py, pos, dstep);
            const unsigned int maxSamples = (unsigned int)P.numIter * (unsigned int)P.numIter;   // nested loops :38-40
            float src_a = 0.0f;
            for (;;) {
                ++nsamples;
                float4 src;
                if (shade_sample<LAYOUT, ILLUM, NGATE, SOF>(P, S, pos, dir, src)) {
                    composite(dest, src);
                    src_a = src.w;
                }
                // :88-93
                pos.x = __fadd_rn(pos.x, dstep.x);
                pos.y = __fadd_rn(pos.y, dstep.y);
                pos.z = __fadd_rn(pos.z, dstep.z);
                if (outside_box(P, pos) || (src_a > 0.95f) || nsamples >= maxSamples) break;   // Q4: src.a
            }
        }
        const int o = lb * kBlockPixels + ly * kBlockDim + lx;
        P.tiles[o] = dest;
        if (P.samplesPerPixel) P.samplesPerPixel[o] = nsamples;
        if (P.sampleCounter) {
            unsigned int tot = __reduce_add_sync(0xffffffffu, nsamples);
            if (lane == 0 && tot) atomicAdd(P.sampleCounter, (unsigned long long)tot);
        }
    }
}

// K1 (sample-parallel pipeline) ----------------------------------------------------------------------------
// The LIC integral of a ray sample does not depend on any other sample; only the compositing and the
// `src.a > 0.95` termination are sequential along a ray.  So the frame is computed in three kernels:
//   ray_setup_kernel     one warp per 8x4 ray tile: entry point, direction, sample count n (march without shading),
//                        slot allocation in the src buffer, work items (tile, k) of the first depth window
//   lic_sample_kernel    persistent warps pull items (tile, k) from a queue; the 32 lanes shade sample k of the 32
//                        rays of the tile (the dominant kernel: perfectly balanced, full occupancy at any image size)
//   composite_kernel     one thread per ray: front-to-back over-operator in sample order with the shader's clamp and
//                        early termination; emits the next window's items for rays still alive
// Depth windows bound the speculative work past an early termination (4, 8, 16, ... samples); when the transfer
// function cannot produce src.a > 0.95 a single window covers the whole ray.

__device__ __forceinline__ int tile_pixel(const DevParams &P, int lt, int lane, int &px, int &py)
{
    // local tile lt = local block * 8 + sub-tile; sub-tiles 8x4 pixels arranged 2 x 4 inside the 16x16 block
    const int lb = lt >> 3, sub = lt & 7;
    const int b = P.rank + lb * P.world;
    const int lx = (sub & 1) * 8 + (lane & 7), ly = (sub >> 1) * 4 + (lane >> 3);
    int bx, by;
    block_xy(P, b, bx, by);
    px = bx * kBlockDim + lx;
    py = by * kBlockDim + ly;
    return lb * kBlockPixels + ly * kBlockDim + lx;
}

__global__ void __launch_bounds__(256) ray_setup_kernel(const __grid_constant__ DevParams P)
{
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        int px, py;
        const int o = tile_pixel(P, lt, lane, px, py);
        int n = 0;
        f3 pos = mk3(0, 0, 0), dir = mk3(0, 0, 0), dstep;
        float e[3];
        if (px < P.width && py < P.height && pixel_ray(P, px, py, e)) {
            ray_setup(P, e, pos, dir, dstep);
            mc_offset(P, px, py, pos, dstep);
            const int maxSamples = (int)min((unsigned int)P.numIter * (unsigned int)P.numIter, 0x7fffffffu);   // numIter <= 32768
            f3 q = pos;
            for (;;) {                                  // the march of lic3d_fragment.glsl:38-95 without the shading
                ++n;
                q.x = __fadd_rn(q.x, dstep.x); q.y = __fadd_rn(q.y, dstep.y); q.z = __fadd_rn(q.z, dstep.z);
                if (outside_box(P, q) || n >= maxSamples) break;
            }
        }
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        unsigned int base = 0;
        if (lane == 0 && nmax > 0) base = atomicAdd(P.slotAlloc, (unsigned int)nmax);
        base = __shfl_sync(0xffffffffu, base, 0);
        const int ray = lt * 32 + lane;
        P.rayA[ray] = make_float4(pos.x, pos.y, pos.z, __int_as_float(n));
        P.rayB[ray] = make_float4(dir.x, dir.y, dir.z, __int_as_float(n > 0 ? 0 : -1));   // state: consumed samples, < 0 = finished
        if (lane == 0) {
            // rows of march checkpoints (ray_checkpoint_kernel): samples kCkStride, 2 kCkStride, ... < nmax
            P.tileCk[lt] = nmax > kCkStride ? atomicAdd(P.ckAlloc, (unsigned int)((nmax - 1) / kCkStride)) : 0u;
            P.tileRec[lt] = make_uint2(base, (unsigned int)nmax);
            P.tileLive[lt] = (unsigned int)nmax;
            if (nmax > 0) atomicMax(P.nMaxGlobal, (unsigned int)nmax);
        }
        P.tiles[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.samplesPerPixel) P.samplesPerPixel[o] = 0;
        // the work items of the first window are built afterwards (item_bucket_kernel, or composite_kernel run on the empty
        // window [0,0)), once the host has sized the src / item buffers from slotAlloc
    }
}

// Position checkpoints of the march.  The shader accumulates pos += dir * stepSize with one rounding per step, so sample k's
// position is only reachable by k dependent additions from the entry point; a warp that shades sample k of its tile would spend
// k x 3 FADD on it (cfg3: ~60 on average, cfg4 ~120).  After the set-up has sized the buffers this kernel repeats the march once
// per ray and keeps every kCkStride-th position, so the sample kernel adds at most kCkStride - 1 steps.  Same additions in the
// same order: same bits.  (1 byte per ray sample; runs only when the view changed.)
__global__ void __launch_bounds__(256) ray_checkpoint_kernel(const __grid_constant__ DevParams P)
{
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        const int nmax = (int)P.tileRec[lt].y;
        if (nmax <= kCkStride) continue;
        const int ray = lt * 32 + lane;
        const float4 A = P.rayA[ray], B = P.rayB[ray];
        const f3 dstep = mk3(__fmul_rn(B.x, P.stepSize), __fmul_rn(B.y, P.stepSize), __fmul_rn(B.z, P.stepSize));
        f3 q = mk3(A.x, A.y, A.z);
        float4 *ck = P.rayCk + (size_t)P.tileCk[lt] * 32 + lane;
        const int rows = (nmax - 1) / kCkStride;
        for (int j = 0; j < rows; ++j) {
#pragma unroll
            for (int i = 0; i < kCkStride; ++i) {
                q.x = __fadd_rn(q.x, dstep.x); q.y = __fadd_rn(q.y, dstep.y); q.z = __fadd_rn(q.z, dstep.z);
            }
            ck[(size_t)j * 32] = make_float4(q.x, q.y, q.z, 0.0f);     // rays shorter than the tile's longest keep marching: never read
        }
    }
}

// The same view as the previous frame (camera, frame size, sample spacing, partition ... unchanged): the ray records, the src
// rows and the first window's item list of that frame are still valid -- only what a frame consumes is put back: ray state,
// tile accumulators, per-frame counters.  (ray_setup_kernel would recompute exactly the same records.)
__global__ void __launch_bounds__(256) ray_reset_kernel(const __grid_constant__ DevParams P, unsigned int *counters)
{
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        // [0,1] ray samples, [2] block queue, [4],[5] item counts of the later windows, [6] item queue head; [3] src rows,
        // [7] longest ray and [8] the first window's item count are kept
        const int t = threadIdx.x;
        if (t != 3 && t != 7) counters[t] = 0u;
    }
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        int px, py;
        const int o = tile_pixel(P, lt, lane, px, py);
        const int ray = lt * 32 + lane;
        const int n = __float_as_int(P.rayA[ray].w);
        P.rayB[ray].w = __int_as_float(n > 0 ? 0 : -1);
        if (lane == 0) P.tileLive[lt] = P.tileRec[lt].y;
        P.tiles[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.samplesPerPixel) P.samplesPerPixel[o] = 0;
    }
}

// ---- view-aligned slicing (VOLIC_SLICING): VV/slicing.cpp:42-114, VV/renderer.cpp:1123-1267 ---------------------
// The slice polygons are planes v.(p - center) = d_i clipped to the box; under a pixel the fragment of slice i is the
// intersection of the pixel ray with that plane.  Same double arithmetic, same order as the oracle.
struct PixelDir { double o[3], d[3]; };

__device__ __forceinline__ PixelDir pixel_dir(const DevParams &P, int px, int py)
{
    PixelDir r;
    const double ex = __dmul_rn(__dmul_rn(__dsub_rn(__ddiv_rn(__dmul_rn(2.0, (double)px + 0.5), (double)P.width), 1.0), P.tanHalf), P.aspect);
    const double ey = __dmul_rn(__dsub_rn(__ddiv_rn(__dmul_rn(2.0, (double)py + 0.5), (double)P.height), 1.0), P.tanHalf);
    const double ez = -1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.d[i] = __dadd_rn(__dadd_rn(__dmul_rn(P.rot[i], ex), __dmul_rn(P.rot[3 + i], ey)), __dmul_rn(P.rot[6 + i], ez));
        r.o[i] = P.camD[i];
    }
    return r;
}

// plane offset of slice i: -0.5 d + (i + 0.5) d / numSlices in float (VV/slicing.cpp:114)
__device__ __forceinline__ float slice_offset(const DevParams &P, int slice)
{
    return __fadd_rn(__fmul_rn(-0.5f, P.slD), __fdiv_rn(__fmul_rn(__fadd_rn((float)slice, 0.5f), P.slD), (float)P.slNum));
}

__device__ __forceinline__ bool slice_fragment(const DevParams &P, const PixelDir &r, int slice, float g[3])
{
    const double v0 = (double)P.slV[0], v1 = (double)P.slV[1], v2 = (double)P.slV[2];
    const double a = __dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(r.o[0], P.slCenter[0]), v0), __dmul_rn(__dsub_rn(r.o[1], P.slCenter[1]), v1)),
                               __dmul_rn(__dsub_rn(r.o[2], P.slCenter[2]), v2));
    const double b = __dadd_rn(__dadd_rn(__dmul_rn(r.d[0], v0), __dmul_rn(r.d[1], v1)), __dmul_rn(r.d[2], v2));
    if (b == 0.0) return false;
    const double t = __ddiv_rn(__dsub_rn((double)slice_offset(P, slice), a), b);
    if (!(t >= P.nearD && t <= P.farD)) return false;
    double pd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double p = __dadd_rn(r.o[i], __dmul_rn(t, r.d[i]));
        if (p < 0.0 || p > P.extent[i]) return false;
        pd[i] = p;
        g[i] = (float)p;
    }
    // user clip planes clip the slice polygons (VV/renderer.cpp:156-163)
    for (int j = 0; j < P.nClip; ++j) {
        const double *c = P.clipEq[j];
        const double q0 = __dsub_rn(pd[0], P.slCenter[0]), q1 = __dsub_rn(pd[1], P.slCenter[1]), q2 = __dsub_rn(pd[2], P.slCenter[2]);
        if (__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[0], q0), __dmul_rn(c[1], q1)), __dmul_rn(c[2], q2)), c[3]) < 0.0) return false;
    }
    return true;
}

// set-up for slicing: per pixel the first slice with a fragment and the span up to the last one
__global__ void __launch_bounds__(256) slice_setup_kernel(const __grid_constant__ DevParams P)
{
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        int px, py;
        const int o = tile_pixel(P, lt, lane, px, py);
        int first = -1, last = -1;
        if (px < P.width && py < P.height) {
            const PixelDir r = pixel_dir(P, px, py);
            float g[3];
            for (int i = 0; i < P.slNum; ++i)
                if (slice_fragment(P, r, i, g)) { if (first < 0) first = i; last = i; }
        }
        // FBO path: Renderer::sliceVolume swaps its two image textures before EVERY slice and a slice writes only the pixels its polygon
        // covers (VV/renderer.cpp:1201-1225); the frame is the target of the last slice N - 1.  A pixel's fragments are consecutive
        // slices (ray / convex region), so along them the value read back is the value just written -- except that the LAST fragment
        // lands in the texture that is not displayed when its slice has the other parity than N - 1: it is dropped from the frame
        // (and still counted as a ray sample if the frame buffer it read was not opaque, like every fragment that did work).
        int dropped = 0;
        if (P.slicing == 1 && last >= 0 && ((last ^ (P.slNum - 1)) & 1)) { dropped = 1; --last; }
        // a tile's items address slices first_tile + k, so that the 32 lanes of an item shade the same slice
        const int tfirst = __reduce_min_sync(0xffffffffu, first < 0 ? 0x7fffffff : first);
        const int tlast = __reduce_max_sync(0xffffffffu, (last >= first) ? last : -1);
        const int nmax = (tlast >= 0 && tfirst != 0x7fffffff) ? tlast - tfirst + 1 : 0;
        const int n = (first >= 0 && last >= first) ? last - tfirst + 1 : 0;       // per ray: samples up to its last displayed fragment
        unsigned int base = 0;
        if (lane == 0 && nmax > 0) base = atomicAdd(P.slotAlloc, (unsigned int)nmax);
        base = __shfl_sync(0xffffffffu, base, 0);
        const int ray = lt * 32 + lane;
        P.rayA[ray] = make_float4(__int_as_float(tfirst), __int_as_float(dropped), 0.f, __int_as_float(n));
        P.rayB[ray] = make_float4(0.f, 0.f, 0.f, __int_as_float(n > 0 ? 0 : -1));
        if (lane == 0) {
            P.tileRec[lt] = make_uint2(base, (unsigned int)nmax);
            P.tileLive[lt] = (unsigned int)nmax;
            if (nmax > 0) atomicMax(P.nMaxGlobal, (unsigned int)nmax);
        }
        // a pixel without (displayed) fragments: the cleared target; without the FBO the white plane that is blended over everything
        // at the end (VV/renderer.cpp:1236-1255) gives (1,1,1,1).  A dropped single fragment still counts as a ray sample: it read
        // the cleared (transparent) target and did its work.
        const float bg = (P.slicing == 2 && n == 0) ? 1.0f : 0.0f;
        P.tiles[o] = make_float4(bg, bg, bg, bg);
        const unsigned int lone = (n == 0 && dropped) ? 1u : 0u;
        if (P.samplesPerPixel) P.samplesPerPixel[o] = lone;
        if (P.sampleCounter) {
            const unsigned int tot = __reduce_add_sync(0xffffffffu, lone);
            if (lane == 0 && tot) atomicAdd(P.sampleCounter, (unsigned long long)tot);
        }
    }
}

// NL: RGBA-noise layout of the gradient build as a compile-time parameter (the walk loop holds one sampler, not both)
// XF: coordinate fast paths (XF_GUARD | XF_NSHARE | XF_SSHARE) the host found applicable to this frame
template <int LAYOUT, int ILLUM, bool NGATE, bool SOF, int NL, int XF>
__global__ void __launch_bounds__(LicShape<ILLUM>::kThreads, LicShape<ILLUM>::kMinCtas) lic_sample_kernel(const __grid_constant__ DevParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &S = *reinterpret_cast<SharedTables *>(smem_raw);
    const unsigned int nItems = *P.itemCount;
    if (nItems == 0) return;
    load_tables(P, S);
    const int lane = threadIdx.x & 31;
    // Every warp pops single items from ONE global queue and works on its own (no barrier in the loop).  The list is bucket-major
    // (band, chunk of 8 depths), so the items in flight at any moment -- the last ~4700 pops -- are one thin slab of the volume: that
    // tight window is what keeps the gathers in L2.  Measured alternatives: a CTA pops a chunk of 8..64 consecutive items and its
    // warps walk it together behind a barrier (round 1: L1 hit rate unchanged, issue utilisation 72 % -> 64 % from the barrier);
    // CTA-affine hand-out without a barrier (chunks of 4..32 consecutive items dealt round-robin to the CTAs, each CTA popping from
    // its own cursor in shared memory so that its warps shade neighbouring depths of one ray tile, last 10 % through the global
    // queue): cfg3 +12 %, cfg2 +10 %, cfg4 +18 % SLOWER -- the CTAs drift apart and the window in L2 widens
    // (profiles/r02/ab18_cta_affine_items.log).
    for (;;) {
      {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(P.itemHead, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nItems) break;
        const uint2 it = P.items[i];
        const int ray = (int)it.x * 32 + lane;
        const int k = (int)it.y;
        const float4 A = P.rayA[ray];
        const int n = __float_as_int(A.w);
        if (k >= n) continue;
        const float4 B = P.rayB[ray];
        if (__float_as_int(B.w) < 0) continue;            // ray finished in an earlier window
        f3 dir, pos;
        float4 src;
        bool have = true;
        if (P.slicing) {
            // fragment of slice (first slice of the tile + k) under this pixel: lic3d_slicing_fragment.glsl:17-25
            int px, py;
            tile_pixel(P, (int)it.x, lane, px, py);
            const PixelDir r = pixel_dir(P, px, py);
            float g[3];
            have = slice_fragment(P, r, __float_as_int(A.x) + k, g);
            if (have) {
                f3 dstep;
                ray_setup(P, g, pos, dir, dstep);
                mc_offset(P, px, py, pos, dstep);     // lic3d_slicing_fragment.glsl:31-33
            }
        } else {
            dir = mk3(B.x, B.y, B.z);
            const f3 dstep = mk3(__fmul_rn(dir.x, P.stepSize), __fmul_rn(dir.y, P.stepSize), __fmul_rn(dir.z, P.stepSize));
            // pos += dir * stepSize, k times, as the shader accumulates it: from the last checkpoint of the march at or before k
            f3 q = mk3(A.x, A.y, A.z);
            const int ckRow = k / kCkStride;               // warp-uniform
            if (ckRow > 0) {
                const float4 c = P.rayCk[((size_t)P.tileCk[it.x] + ckRow - 1) * 32 + lane];
                q = mk3(c.x, c.y, c.z);
            }
            for (int j = ckRow * kCkStride; j < k; ++j) {
                q.x = __fadd_rn(q.x, dstep.x); q.y = __fadd_rn(q.y, dstep.y); q.z = __fadd_rn(q.z, dstep.z);
            }
            pos = q;
        }
        if (!have) src = make_float4(0.f, 0.f, 0.f, -2.0f);                                                               // no fragment
        else if (!shade_sample<LAYOUT, ILLUM, NGATE, SOF, NL, XF>(P, S, pos, dir, src)) src = make_float4(0.f, 0.f, 0.f, -1.0f);   // gated
        const uint2 tr = P.tileRec[it.x];
        P.src[((size_t)tr.x + k) * 32 + lane] = src;
      }
    }
}

// (Measured and dropped: compositing fused into lic_sample_kernel for single-window frames -- a per-tile counter of shaded items, the
// warp that shades a tile's last item blends the tile's rays right there (fence + atomic per item, samples read past L1, the blend
// as a called function so that the walk loops' register allocation does not see it).  Frames identical, racecheck clean, one launch
// less -- and no faster: cfg3 11.87 -> 11.92 ms, cfg2 +0.8 %, cfg4 +2.2 %, a rank's 1/8 share 1.580 -> 1.583 ms: the per-item
// synchronisation added to the hot kernel costs what the 72 us launch saves.  profiles/r02/ab25_fused_compositing.log.)
__global__ void __launch_bounds__(256) composite_kernel(const __grid_constant__ DevParams P)
{
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        const uint2 tr = P.tileRec[lt];
        const int nmax = (int)tr.y;
        if (nmax <= P.win0) continue;                      // warp-uniform: nothing of this tile in the window
        if (P.win0 > 0 && P.tileLive[lt] == 0u) continue;  // ... or every ray of the tile finished in an earlier window (no ray record is read)
        const int ray = lt * 32 + lane;
        const float4 A = P.rayA[ray];
        float4 B = P.rayB[ray];
        const int n = __float_as_int(A.w);
        int state = __float_as_int(B.w);
        int px, py;
        const int o = tile_pixel(P, lt, lane, px, py);
        unsigned int consumed = 0;
        if (state >= 0) {
            float4 dest = P.tiles[o];
            if (P.slicing == 2) { dest.x = unorm8_rn(dest.x); dest.y = unorm8_rn(dest.y); dest.z = unorm8_rn(dest.z); dest.w = unorm8_rn(dest.w); }   // bytes
            const int kend = min(n, P.win1);
            bool done = false;
            // The blend is a serial chain but the loads are not: fetch 8 samples ahead so that one ray keeps 8 requests in
            // flight (the kernel is latency-bound: a 222-sample ray is 222 dependent round trips otherwise).
            constexpr int kAhead = 8;
            for (int k0 = P.win0; k0 < kend && !done; k0 += kAhead) {
                float4 buf[kAhead];
#pragma unroll
                for (int j = 0; j < kAhead; ++j) buf[j] = P.src[((size_t)tr.x + min(k0 + j, kend - 1)) * 32 + lane];
#pragma unroll
                for (int j = 0; j < kAhead; ++j) {
                    if (done || k0 + j >= kend) continue;
                    const float4 s = buf[j];
                    if (P.slicing == 2) {
                        // without the FBO (lic3d_slicingblend_fragment.glsl:5-67): every fragment's colour -- src.rgb * src.a, src.a,
                        // clamped by the GL -- is blended into the RGBA8 back buffer, no skip
                        if (s.w < 0.0f) continue;                         // no fragment, or gated off: colour 0 leaves the buffer as it is
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        composite(v, s);                                  // clamp((1 - 0) (src.rgb * src.a, src.a) + 0)
                        if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) ++consumed;
                        blend8(dest, v);
                        continue;
                    }
                    if (P.slicing) {
                        // lic3d_slicing_fragment.glsl:14: a fragment only works while the frame buffer has dest.a < 0.95
                        if (s.w == -2.0f) continue;                       // no fragment of this slice under the pixel
                        if (!(dest.w < 0.95f)) { done = true; continue; }
                        ++consumed;
                        if (s.w >= 0.0f) {
                            composite(dest, s);
                            // the target is GL_RGBA16F_ARB (VV/renderer.cpp:566-606): the next slice reads this rounded to fp16
                            dest.x = half_round(dest.x); dest.y = half_round(dest.y); dest.z = half_round(dest.z); dest.w = half_round(dest.w);
                        }
                        continue;
                    }
                    ++consumed;
                    if (s.w >= 0.0f) {
                        composite(dest, s);
                        if (s.w > 0.95f) done = true;                     // early ray termination on src.a (Q4)
                    }
                }
            }
            if (kend >= n) {
                // the fragment that went into the other ping-pong target (slice_setup_kernel): it did its work if what it read was not opaque
                if (P.slicing == 1 && !done && __float_as_int(A.y) && dest.w < 0.95f) ++consumed;
                done = true;
            }
            if (P.slicing == 2) {
                if (done) blend8(dest, make_float4(1.f, 1.f, 1.f, 1.f));       // the screen-filling white plane, VV/renderer.cpp:1236-1255
                dest.x = __fdiv_rn(dest.x, 255.0f); dest.y = __fdiv_rn(dest.y, 255.0f); dest.z = __fdiv_rn(dest.z, 255.0f); dest.w = __fdiv_rn(dest.w, 255.0f);
            }
            state += (int)consumed;
            P.tiles[o] = dest;
            if (P.samplesPerPixel) P.samplesPerPixel[o] = (unsigned int)state;
            if (done) state = -1 - state;
            B.w = __int_as_float(state);
            P.rayB[ray] = B;
        }
        if (P.sampleCounter) {
            unsigned int tot = __reduce_add_sync(0xffffffffu, consumed);
            if (lane == 0 && tot) atomicAdd(P.sampleCounter, (unsigned long long)tot);
        }
        // what is left of the tile for the next window: the longest ray still alive
        const int live = __reduce_max_sync(0xffffffffu, state >= 0 ? n : 0);
        if (lane == 0) P.tileLive[lt] = (unsigned int)live;
        const int cnt = P.emitItems ? min(live, P.win2) - P.win1 : 0;
        if (cnt > 0) {
            unsigned int ib = 0;
            if (lane == 0) ib = atomicAdd(P.itemCountNext, (unsigned int)cnt);
            ib = __shfl_sync(0xffffffffu, ib, 0);
            for (int k = lane; k < cnt; k += 32) P.itemsNext[ib + k] = make_uint2((unsigned int)lt, (unsigned int)(P.win1 + k));
        }
    }
}

// ---- depth-major work-item order --------------------------------------------------------------------------------
// The item list decides which voxels are live in L2 at any time.  Tile-major order (all depths of a tile, then the
// next tile) re-reads every voxel from DRAM once per row of tiles whose streamlines reach it (measured 6.9 GB per cfg3
// frame against 0.55 GB of compulsory bytes).  Here items are bucketed by (band of block rows, chunk of 8 depths) and the
// list is bucket-major: the frame is swept band by band, front to back, so the live set is a slab of the band a few
// streamline lengths thick and each voxel comes from DRAM about once per band.
// pass 0 counts the items of every bucket, bucket_scan_kernel turns counts into offsets, pass 1 writes the items.
__global__ void __launch_bounds__(256) item_bucket_kernel(const __grid_constant__ DevParams P, int pass)
{
    const int lane = threadIdx.x & 31;
    const int nTiles = P.nLocalBlocks * 8;
    // items of the window [win1, win2) for the tiles that still have live rays reaching into it, in chunks of 8 depths from win1
    for (int lt = blockIdx.x * 8 + (threadIdx.x >> 5); lt < nTiles; lt += gridDim.x * 8) {
        const int nmax = min((int)P.tileLive[lt], P.win2);
        if (nmax <= P.win1) continue;
        int bbx, bby;
        block_xy(P, P.rank + (lt >> 3) * P.world, bbx, bby);
        const int band = bby / P.bandRows;
        const int nchunks = (nmax - P.win1 + 7) >> 3;
        for (int c = lane; c < nchunks; c += 32) {
            const int k0 = P.win1 + 8 * c;
            const int cnt = min(8, nmax - k0);
            const int bucket = band * P.nDepthChunks + c;
            if (pass == 0) {
                atomicAdd(P.bucketCount + bucket, (unsigned int)cnt);
            } else {
                const unsigned int at = P.bucketBase[bucket] + atomicAdd(P.bucketFill + bucket, (unsigned int)cnt);
                for (int j = 0; j < cnt; ++j) P.itemsNext[at + j] = make_uint2((unsigned int)lt, (unsigned int)(k0 + j));
            }
        }
    }
}

// exclusive scan of the bucket counts (a few thousand entries: one CTA), total -> item count of the window
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const __grid_constant__ DevParams P, int nBuckets)
{
    __shared__ unsigned int s_part[1024];
    const int t = threadIdx.x;
    const int per = (nBuckets + 1023) / 1024;
    unsigned int sum = 0;
    for (int i = t * per; i < min(nBuckets, (t + 1) * per); ++i) sum += P.bucketCount[i];
    s_part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned int v = (t >= off) ? s_part[t - off] : 0u;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    unsigned int run = s_part[t] - sum;   // exclusive prefix of this thread's segment
    for (int i = t * per; i < min(nBuckets, (t + 1) * per); ++i) {
        P.bucketBase[i] = run;
        run += P.bucketCount[i];
    }
    if (t == 1023) *P.itemCountNext = s_part[1023];
}

// K3 ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(256) volume_raycast_kernel(const __grid_constant__ DevParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &S = *reinterpret_cast<SharedTables *>(smem_raw);
    load_tables(P, S);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp & 1) * 8 + (lane & 7);
    const int ly = (warp >> 1) * 4 + (lane >> 3);
    for (;;) {
        if (threadIdx.x == 0) S.block = (int)atomicAdd(P.blockCounter, 1u);
        __syncthreads();
        const int lb = S.block;
        __syncthreads();
        if (lb >= P.nLocalBlocks) break;
        const int b = P.rank + lb * P.world;
        int bx, by;
        block_xy(P, b, bx, by);
        const int px = bx * kBlockDim + lx;
        const int py = by * kBlockDim + ly;
        float4 dest = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int nsamples = 0;
        float e[3];
        if (px < P.width && py < P.height && pixel_ray(P, px, py, e)) {
            f3 pos = mk3(__fmul_rn(e[0], P.scaleVol[0]), __fmul_rn(e[1], P.scaleVol[1]), __fmul_rn(e[2], P.scaleVol[2]));
            f3 gd = normalize_rn(mk3(__fsub_rn(e[0], P.camera[0]), __fsub_rn(e[1], P.camera[1]), __fsub_rn(e[2], P.camera[2])));
            f3 dstep = mk3(__fmul_rn(__fmul_rn(gd.x, P.scaleVol[0]), P.stepSize), __fmul_rn(__fmul_rn(gd.y, P.scaleVol[1]), P.stepSize),
                           __fmul_rn(__fmul_rn(gd.z, P.scaleVol[2]), P.stepSize));
            const unsigned int maxSamples = (unsigned int)P.numIter * (unsigned int)P.numIter;
            for (;;) {
                ++nsamples;
                float4 vd = fetch_field<LAYOUT, false>(P, pos.x, pos.y, pos.z);                 // :40
                float lic = fetch_licvol(P, pos.x, pos.y, pos.z);                               // :41
                float4 tf = tf_lookup(S.tf, vd.z);                                              // :47
                float4 src = illum_lic(P, S.opac, lic, tf);                                     // :51
                const float k = 1.0f - dest.w;
                dest.x = clamp01(fmaf(k, src.x * src.w, dest.x));
                dest.y = clamp01(fmaf(k, src.y * src.w, dest.y));
                dest.z = clamp01(fmaf(k, src.z * src.w, dest.z));
                dest.w = clamp01(fmaf(k, src.w, dest.w));
                pos.x = __fadd_rn(pos.x, dstep.x);
                pos.y = __fadd_rn(pos.y, dstep.y);
                pos.z = __fadd_rn(pos.z, dstep.z);
                const bool outside = pos.x < 0.0f || pos.x > P.texMax[0] || pos.y < 0.0f || pos.y > P.texMax[1] ||
                                     pos.z < 0.0f || pos.z > P.texMax[2] || (dest.w > 0.95f);   // :65 dest.a
                if (outside || nsamples >= maxSamples) break;
            }
        }
        const int o = lb * kBlockPixels + ly * kBlockDim + lx;
        P.tiles[o] = dest;
        if (P.samplesPerPixel) P.samplesPerPixel[o] = nsamples;
        if (P.sampleCounter) {
            unsigned int tot = __reduce_add_sync(0xffffffffu, nsamples);
            if (lane == 0 && tot) atomicAdd(P.sampleCounter, (unsigned long long)tot);
        }
    }
}

// K2 ------------------------------------------------------------------------------------------------
// one thread per voxel of the target; CTA = 8x8x4 voxels (warp = 8x4 voxels of one z-layer)
#ifndef LICVOL_MIN_CTAS
#define LICVOL_MIN_CTAS 4   // resident CTAs per SM lic_volume_kernel is compiled for (64 registers; the kernel is latency-bound on large chaotic
                           // fields: 512^3 curl noise 170.5 -> 156.4 ms against 3 CTAs / 80 registers, 256^3 tornado 9.76 -> 10.0 ms; 5 CTAs / 48
                           // registers spill: 264 ms)
#endif
// (Measured and dropped: an L2 prefetch of the cell the next Heun step's predictor is expected in -- position + 2 x the current
// displacement, no destination registers -- made this kernel 36 % (512^3) to 41 % (1024^3) SLOWER on the curl-noise fields and 6 %
// slower on the 256^3 tornado: it is bound by DRAM gather bandwidth, not by the number of loads in flight, and every mispredicted
// line is paid for; profiles/r02/licvol22_l2_prefetch.log.)
template <int LAYOUT, bool GRAD, bool NGATE, bool SOF>
__global__ void __launch_bounds__(256, LICVOL_MIN_CTAS) lic_volume_kernel(const __grid_constant__ DevParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &S = *reinterpret_cast<SharedTables *>(smem_raw);
    load_tables(P, S);
    const int bxn = (P.ow + 7) / 8, byn = (P.oh + 7) / 8, bzn = (P.oz1 - P.oz0 + 3) / 4;
    const long long nblocks = (long long)bxn * byn * bzn;
    const int t = threadIdx.x;
    const int tx = t & 7, ty = (t >> 3) & 7, tz = t >> 6;
    for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int bx = (int)(b % bxn), by = (int)((b / bxn) % byn), bz = (int)(b / ((long long)bxn * byn));
        const int x = bx * 8 + tx, y = by * 8 + ty, z = P.oz0 + bz * 4 + tz;
        if (x >= P.ow || y >= P.oh || z >= P.oz1) continue;
        // fragment of the full-screen quad of layer z: ((x+.5)/w, (y+.5)/h, (z+.5)/d), VV/renderer.cpp:1353-1358
        f3 g = mk3(__fdiv_rn((float)x + 0.5f, (float)P.ow), __fdiv_rn((float)y + 0.5f, (float)P.oh), __fdiv_rn((float)z + 0.5f, (float)P.od));
        f3 pos = mk3(__fmul_rn(g.x, P.scaleVol[0]), __fmul_rn(g.y, P.scaleVol[1]), __fmul_rn(g.z, P.scaleVol[2]));
        float4 vd = fetch_field<LAYOUT, true>(P, pos.x, pos.y, pos.z);
        float r;
        if (GRAD) r = compute_lic_grad<LAYOUT, SOF, false>(P, S.kw, pos, vd).x;
        else r = compute_lic_scalar<LAYOUT, NGATE, SOF, false>(P, S.kw, pos, vd);
        r *= P.licScale;
        if (P.licvolFp16) r = __half2float(__float2half_rn(r));
        P.licvol_out[((size_t)z * P.oh + y) * P.ow + x] = r;
    }
}

// diagnostic: one direction of the walk of compute_lic_scalar / compute_lic_grad from one position, step by step, with the device
// functions of the hot path: out[16 i ..] = (position.xyz, field sample rgb, noise tap (.a of the RGBA tap), kernel weight,
// predictor position Pos2.xyz, field sample at it rgb, 0, 0)
template <int LAYOUT, bool GRAD, int XF>
__global__ void debug_walk_kernel(const __grid_constant__ DevParams P, float px, float py, float pz, int dirSign, int nSteps, float *out)
{
    const f3 pos = mk3(px, py, pz);
    const float4 centre = fetch_field<LAYOUT, true>(P, pos.x, pos.y, pos.z);
    Walker w = make_walker(pos, centre);
    const float *kw = P.kw + 1 + (dirSign < 0 ? 0 : P.nBwd);
    for (int k = 0; k < nSteps; ++k) {
        float *o = out + 16 * k;
        heun_step<LAYOUT, false, XF>(P, w, dirSign < 0 ? -P.h : P.h, threadIdx.x == 0 ? o + 8 : nullptr);
        float tap;
        if (GRAD) tap = hi2(noise_tap_rgba<2, XF>(P, w, (XF & XF_NSHARE) != 0).ba);
        else tap = noise_tap<true>(P, mk3(lo2(w.qxy), hi2(w.qxy), w.qz));
        if (threadIdx.x == 0) {
            o[0] = lo2(w.qxy); o[1] = hi2(w.qxy); o[2] = w.qz;
            o[3] = lo2(w.vrg); o[4] = hi2(w.vrg); o[5] = w.vb;
            o[6] = tap; o[7] = kw[k];
        }
    }
}

template <int LAYOUT>
static cudaError_t launch_debug_walk_layout(const DevParams &P, bool grad, int xf, const float pos[3], int dirSign, int nSteps, float *out, cudaStream_t st)
{
    if (grad) {
        if (xf == 3) debug_walk_kernel<LAYOUT, true, 3><<<1, 32, 0, st>>>(P, pos[0], pos[1], pos[2], dirSign, nSteps, out);
        else if (xf == 1) debug_walk_kernel<LAYOUT, true, 1><<<1, 32, 0, st>>>(P, pos[0], pos[1], pos[2], dirSign, nSteps, out);
        else debug_walk_kernel<LAYOUT, true, 0><<<1, 32, 0, st>>>(P, pos[0], pos[1], pos[2], dirSign, nSteps, out);
    } else {
        if (xf == 1) debug_walk_kernel<LAYOUT, false, 1><<<1, 32, 0, st>>>(P, pos[0], pos[1], pos[2], dirSign, nSteps, out);
        else debug_walk_kernel<LAYOUT, false, 0><<<1, 32, 0, st>>>(P, pos[0], pos[1], pos[2], dirSign, nSteps, out);
    }
    return cudaGetLastError();
}

cudaError_t launch_debug_walk(const DevParams &P, int layout, bool grad, int xf, const float pos[3], int dirSign, int nSteps, float *out, cudaStream_t st)
{
    if (layout == LAYOUT_QUAD) return launch_debug_walk_layout<LAYOUT_QUAD>(P, grad, xf, pos, dirSign, nSteps, out, st);
    if (layout == LAYOUT_PAIR) return launch_debug_walk_layout<LAYOUT_PAIR>(P, grad, xf, pos, dirSign, nSteps, out, st);
    return cudaErrorNotSupported;
}

// K5 ------------------------------------------------------------------------------------------------
// tiles laid out [world][nLocalBlocksMax][256] -> row-major float frame, RGBA8 frame, and displayed RGBA8
__global__ void unblock_kernel(const float4 *__restrict__ tiles, int world, int blocksPerRank, int nBlocksX, int nBlocksY, int skew, int unit,
                               int width, int height, float4 *__restrict__ frame, uchar4 *__restrict__ frame8,
                               uchar4 *__restrict__ display8)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= width || py >= height) return;
    const int b = block_id_u(nBlocksX, skew, world, unit, px / kBlockDim, py / kBlockDim);
    const int r = b % world, lb = b / world;
    float4 c = tiles[((size_t)r * blocksPerRank + lb) * kBlockPixels + (py % kBlockDim) * kBlockDim + (px % kBlockDim)];
    const size_t o = (size_t)py * width + px;
    if (frame) frame[o] = c;
    if (frame8) {
        // GL float -> UNORM8 of the back buffer (VV/renderer.cpp:216-226)
        frame8[o] = make_uchar4((unsigned char)floorf(clamp01(c.x) * 255.0f + 0.5f), (unsigned char)floorf(clamp01(c.y) * 255.0f + 0.5f),
                                (unsigned char)floorf(clamp01(c.z) * 255.0f + 0.5f), (unsigned char)floorf(clamp01(c.w) * 255.0f + 0.5f));
    }
    if (display8) {
        // background_fragment.glsl:9-16: clamp((1 - a) * white + dest)
        const float k = 1.0f - c.w;
        display8[o] = make_uchar4((unsigned char)floorf(clamp01(k + c.x) * 255.0f + 0.5f), (unsigned char)floorf(clamp01(k + c.y) * 255.0f + 0.5f),
                                  (unsigned char)floorf(clamp01(k + c.z) * 255.0f + 0.5f), (unsigned char)floorf(clamp01(k + c.w) * 255.0f + 0.5f));
    }
}

// ------------------------------------------------------------------------------------------------
// launchers

template <class K>
static cudaError_t launch_with_tables(K kernel, const DevParams &P, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 256, smem, st>>>(P);
    return cudaGetLastError();
}

template <int LAYOUT, int ILLUM, bool NGATE>
static cudaError_t launch_raycast_sof(const DevParams &P, bool sof, int grid, size_t smem, cudaStream_t st)
{
    if (sof) return launch_with_tables(lic_raycast_kernel<LAYOUT, ILLUM, NGATE, true>, P, grid, smem, st);
    return launch_with_tables(lic_raycast_kernel<LAYOUT, ILLUM, NGATE, false>, P, grid, smem, st);
}

template <int LAYOUT>
static cudaError_t launch_raycast_layout(const DevParams &P, int illum, bool ngate, bool sof, int grid, size_t smem, cudaStream_t st)
{
    switch (illum) {
    case ILLUM_GRADIENT: return launch_raycast_sof<LAYOUT, ILLUM_GRADIENT, false>(P, sof, grid, smem, st);
    case ILLUM_MALLO:
        return ngate ? launch_raycast_sof<LAYOUT, ILLUM_MALLO, true>(P, sof, grid, smem, st)
                     : launch_raycast_sof<LAYOUT, ILLUM_MALLO, false>(P, sof, grid, smem, st);
    case ILLUM_ZOECKLER:
        return ngate ? launch_raycast_sof<LAYOUT, ILLUM_ZOECKLER, true>(P, sof, grid, smem, st)
                     : launch_raycast_sof<LAYOUT, ILLUM_ZOECKLER, false>(P, sof, grid, smem, st);
    default:
        return ngate ? launch_raycast_sof<LAYOUT, ILLUM_NONE, true>(P, sof, grid, smem, st)
                     : launch_raycast_sof<LAYOUT, ILLUM_NONE, false>(P, sof, grid, smem, st);
    }
}

size_t shared_table_bytes() { return sizeof(SharedTables); }

cudaError_t launch_ray_setup(const DevParams &P, int grid, cudaStream_t st)
{
    if (P.slicing) slice_setup_kernel<<<grid, 256, 0, st>>>(P);
    else ray_setup_kernel<<<grid, 256, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_ray_checkpoints(const DevParams &P, int grid, cudaStream_t st)
{
    ray_checkpoint_kernel<<<grid, 256, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_ray_reset(const DevParams &P, unsigned int *counters, int grid, cudaStream_t st)
{
    ray_reset_kernel<<<grid, 256, 0, st>>>(P, counters);
    return cudaGetLastError();
}

cudaError_t launch_item_buckets(const DevParams &P, int nBuckets, int grid, cudaStream_t st)
{
    item_bucket_kernel<<<grid, 256, 0, st>>>(P, 0);
    bucket_scan_kernel<<<1, 1024, 0, st>>>(P, nBuckets);
    item_bucket_kernel<<<grid, 256, 0, st>>>(P, 1);
    return cudaGetLastError();
}

cudaError_t launch_composite(const DevParams &P, int grid, cudaStream_t st)
{
    composite_kernel<<<grid, 256, 0, st>>>(P);
    return cudaGetLastError();
}

// Ask for the smallest shared-memory carve-out that still holds the resident CTAs' tables, so the rest of the unified
// 256 KB array serves as L1 (the driver's default picked 102 KB of shared memory for 3 x 14 KB; ncu
// launch__shared_mem_config_size).  The gathers of this kernel live on L1 hits.
template <class K>
static cudaError_t prefer_l1(K kernel, size_t smem, int *occ_out = nullptr, int threads = 256)
{
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    const size_t need = (size_t)occ * (smem + 1024);                 // + 1 KB per CTA reserved by the driver
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if (occ_out) *occ_out = occ;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

template <int THREADS, class K>
static cudaError_t launch_sample_kernel(K kernel, const DevParams &P, int grid, size_t smem, cudaStream_t st)
{
    int dev = 0, sms = 148;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = prefer_l1(kernel, smem, &occ, THREADS);
    if (e != cudaSuccess) return e;
    kernel<<<grid > 0 ? grid : sms * occ, THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

// which coordinate fast paths apply (decided per frame on the host, see fill_params): only the hot layouts get them
template <int LAYOUT, int ILLUM, bool NGATE, bool SOF, int NL>
static cudaError_t launch_sample_xf(const DevParams &P, int grid, size_t smem, cudaStream_t st)
{
    if constexpr (LAYOUT != LAYOUT_F4 && !SOF && VV_CELL_REUSE && (NL == 2 || ILLUM != ILLUM_GRADIENT)) {
        if (P.fGuard > 1 && P.guardOk) {
            if constexpr (ILLUM == ILLUM_GRADIENT) {
                if (P.noiseShared) return launch_sample_kernel<LicShape<ILLUM>::kThreads>(lic_sample_kernel<LAYOUT, ILLUM, NGATE, SOF, NL, XF_GUARD | XF_NSHARE>, P, grid, smem, st);
            }
            return launch_sample_kernel<LicShape<ILLUM>::kThreads>(lic_sample_kernel<LAYOUT, ILLUM, NGATE, SOF, NL, XF_GUARD>, P, grid, smem, st);
        }
    }
    return launch_sample_kernel<LicShape<ILLUM>::kThreads>(lic_sample_kernel<LAYOUT, ILLUM, NGATE, SOF, NL, 0>, P, grid, smem, st);
}

template <int LAYOUT, int ILLUM, bool NGATE>
static cudaError_t launch_sample_sof(const DevParams &P, bool sof, int grid, size_t smem, cudaStream_t st)
{
    // only the gradient build samples the RGBA noise; each of its layouts (2 bf16 diff, 1 fp16 pair, 0 u8 quads) is an instantiation
    if (ILLUM == ILLUM_GRADIENT && !P.noise_bf) {
        if (P.noise_pair) {
            constexpr int NL = (ILLUM == ILLUM_GRADIENT) ? 1 : 2;
            if (sof) return launch_sample_xf<LAYOUT, ILLUM, NGATE, true, NL>(P, grid, smem, st);
            return launch_sample_xf<LAYOUT, ILLUM, NGATE, false, NL>(P, grid, smem, st);
        }
        constexpr int NL = (ILLUM == ILLUM_GRADIENT) ? 0 : 2;
        if (sof) return launch_sample_xf<LAYOUT, ILLUM, NGATE, true, NL>(P, grid, smem, st);
        return launch_sample_xf<LAYOUT, ILLUM, NGATE, false, NL>(P, grid, smem, st);
    }
    if (sof) return launch_sample_xf<LAYOUT, ILLUM, NGATE, true, 2>(P, grid, smem, st);
    return launch_sample_xf<LAYOUT, ILLUM, NGATE, false, 2>(P, grid, smem, st);
}

template <int LAYOUT>
static cudaError_t launch_sample_layout(const DevParams &P, int illum, bool ngate, bool sof, int grid, size_t smem, cudaStream_t st)
{
    switch (illum) {
    case ILLUM_GRADIENT: return launch_sample_sof<LAYOUT, ILLUM_GRADIENT, false>(P, sof, grid, smem, st);
    case ILLUM_MALLO:
        return ngate ? launch_sample_sof<LAYOUT, ILLUM_MALLO, true>(P, sof, grid, smem, st)
                     : launch_sample_sof<LAYOUT, ILLUM_MALLO, false>(P, sof, grid, smem, st);
    case ILLUM_ZOECKLER:
        return ngate ? launch_sample_sof<LAYOUT, ILLUM_ZOECKLER, true>(P, sof, grid, smem, st)
                     : launch_sample_sof<LAYOUT, ILLUM_ZOECKLER, false>(P, sof, grid, smem, st);
    default:
        return ngate ? launch_sample_sof<LAYOUT, ILLUM_NONE, true>(P, sof, grid, smem, st)
                     : launch_sample_sof<LAYOUT, ILLUM_NONE, false>(P, sof, grid, smem, st);
    }
}

cudaError_t launch_lic_sample(const DevParams &P, int layout, int illum, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st)
{
    const size_t smem = table_bytes(P);
    if (layout == LAYOUT_PAIR) return launch_sample_layout<LAYOUT_PAIR>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
    if (layout == LAYOUT_QUAD) return launch_sample_layout<LAYOUT_QUAD>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
    return launch_sample_layout<LAYOUT_F4>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
}

cudaError_t launch_lic_raycast(const DevParams &P, int layout, int illum, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st)
{
    const size_t smem = table_bytes(P);
    if (layout == LAYOUT_PAIR) return launch_raycast_layout<LAYOUT_PAIR>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
    if (layout == LAYOUT_QUAD) return launch_raycast_layout<LAYOUT_QUAD>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
    return launch_raycast_layout<LAYOUT_F4>(P, illum, noise_gate, speed_of_flow, grid, smem, st);
}

cudaError_t launch_volume_raycast(const DevParams &P, int layout, int grid, cudaStream_t st)
{
    const size_t smem = table_bytes(P);
    if (layout == LAYOUT_PAIR) return launch_with_tables(volume_raycast_kernel<LAYOUT_PAIR>, P, grid, smem, st);
    if (layout == LAYOUT_QUAD) return launch_with_tables(volume_raycast_kernel<LAYOUT_QUAD>, P, grid, smem, st);
    return launch_with_tables(volume_raycast_kernel<LAYOUT_F4>, P, grid, smem, st);
}

template <class K>
static cudaError_t launch_licvol_kernel(K kernel, const DevParams &P, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = prefer_l1(kernel, smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 256, smem, st>>>(P);
    return cudaGetLastError();
}

template <int LAYOUT, bool GRAD, bool NGATE>
static cudaError_t launch_licvol_sof(const DevParams &P, bool sof, int grid, size_t smem, cudaStream_t st)
{
    if (sof) return launch_licvol_kernel(lic_volume_kernel<LAYOUT, GRAD, NGATE, true>, P, grid, smem, st);
    return launch_licvol_kernel(lic_volume_kernel<LAYOUT, GRAD, NGATE, false>, P, grid, smem, st);
}

cudaError_t launch_lic_volume(const DevParams &P, int layout, bool grad, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st)
{
    const size_t smem = table_bytes(P);
    if (layout == LAYOUT_PAIR) {
        if (grad) return launch_licvol_sof<LAYOUT_PAIR, true, false>(P, speed_of_flow, grid, smem, st);
        return noise_gate ? launch_licvol_sof<LAYOUT_PAIR, false, true>(P, speed_of_flow, grid, smem, st)
                          : launch_licvol_sof<LAYOUT_PAIR, false, false>(P, speed_of_flow, grid, smem, st);
    }
    if (layout == LAYOUT_QUAD) {
        if (grad) return launch_licvol_sof<LAYOUT_QUAD, true, false>(P, speed_of_flow, grid, smem, st);
        return noise_gate ? launch_licvol_sof<LAYOUT_QUAD, false, true>(P, speed_of_flow, grid, smem, st)
                          : launch_licvol_sof<LAYOUT_QUAD, false, false>(P, speed_of_flow, grid, smem, st);
    }
    if (grad) return launch_licvol_sof<LAYOUT_F4, true, false>(P, speed_of_flow, grid, smem, st);
    return noise_gate ? launch_licvol_sof<LAYOUT_F4, false, true>(P, speed_of_flow, grid, smem, st)
                      : launch_licvol_sof<LAYOUT_F4, false, false>(P, speed_of_flow, grid, smem, st);
}

// ---- peer-to-peer tile exchange (multi-GPU, one process per GPU on one NVLink / NVSwitch node) ----------------------
// Replaces "all_gather of the tile buffers" by stores into the peers' memory: every rank writes its finished tiles
// into slot `rank` of every rank's gather buffer (P2P stores over NVLink), fences, and bumps an arrival counter in each
// peer; the consumer waits until `world` arrivals of the frame are in before it un-blocks.  Two gather buffers
// alternate by frame parity: a rank can only be one frame ahead of the slowest peer (its own wait needs that peer's
// arrival, which is stream-ordered after the peer's previous un-block), so the buffer it overwrites is no longer read.
__global__ void __launch_bounds__(256) scatter_tiles_kernel(P2PArgs a)
{
    // blockIdx.y = destination rank: a CTA streams contiguous 16-byte stores to ONE peer (the local re-reads of the tiles
    // hit L2), so that many independent NVLink write streams are in flight per SM
    const int j = blockIdx.y;
    float4 *dst = a.peerTiles[j] + (size_t)a.rank * a.n;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < a.n; i += 4 * stride) {
        const float4 v0 = a.tiles[i], v1 = a.tiles[i + stride], v2 = a.tiles[i + 2 * stride], v3 = a.tiles[i + 3 * stride];
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
    }
    for (; i < a.n; i += stride) dst[i] = a.tiles[i];
    __threadfence_system();                       // this thread's stores are visible system-wide before the counter moves
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(a.doneCounter, 1u);
        if (t == gridDim.x * gridDim.y - 1) {     // last CTA: all tiles of this rank have landed everywhere
            *a.doneCounter = 0;
            __threadfence_system();
            for (int k = 0; k < a.world; ++k) atomicAdd_system(a.peerFlags[k], 1u);
        }
    }
}

// one thread spins until `target` arrivals are in (wrap-safe compare); gives up after ~17 s (2^35 cycles) and raises
// *err instead of hanging the stream when a peer never arrives
__global__ void wait_arrivals_kernel(volatile unsigned int *flag, unsigned int target, unsigned int *err)
{
    const long long t0 = clock64();
    while ((int)(*flag - target) < 0) {
        if (clock64() - t0 > (1LL << 35)) { *err = 1u; break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

cudaError_t launch_scatter_tiles(const P2PArgs &a, int grid, cudaStream_t st)
{
    scatter_tiles_kernel<<<dim3(grid, a.world), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_wait_arrivals(unsigned int *flag, unsigned int target, unsigned int *err, cudaStream_t st)
{
    wait_arrivals_kernel<<<1, 1, 0, st>>>(flag, target, err);
    return cudaGetLastError();
}

cudaError_t launch_unblock(const float4 *tiles, int world, int blocksPerRank, int nBlocksX, int nBlocksY, int skew, int unit, int width, int height,
                           float4 *frame, uchar4 *frame8, uchar4 *display8, cudaStream_t st)
{
    dim3 blk(16, 16), grd((width + 15) / 16, (height + 15) / 16);
    unblock_kernel<<<grd, blk, 0, st>>>(tiles, world, blocksPerRank, nBlocksX, nBlocksY, skew, unit, width, height, frame, frame8, display8);
    return cudaGetLastError();
}

} // namespace vvb200
