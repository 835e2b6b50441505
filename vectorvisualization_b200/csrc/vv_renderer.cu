// vv_renderer.cu -- host side of the hot path: the headless counterpart of `class Renderer`
// (VV/renderer.h:28-299, VV/renderer.cpp) and of the dataset classes that feed it (VV/dataset.cpp), behind the
// C ABI declared in include/vv_c_api.h.  Everything that touches voxels or pixels runs in the CUDA kernels of
// vv_kernels.cu / vv_preprocess.cu; this file owns device memory, derives the parameter block the way
// Renderer::setRenderVolParams does (VV/renderer.cpp:925-996) and sequences launches on one stream.
#include <cmath>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "vv_device.cuh"
#include "vv_host.h"
#include "vv_kernels.h"

namespace vvb200 {
const std::string &last_error_string();
}

using namespace vvb200;

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(VV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)

static const int kNumCounters = 16;

// What the ray records, the src rows and the first window's item list of a frame depend on: compared frame to frame (memcmp of
// a zero-initialised struct) so that an unchanged view neither recomputes them nor reads their sizes back to the host.
struct GeomKey {
    double camD[3], rot[9], tanHalf, aspect, extent[3], nearD, farD, clipEq[3][4], clipN[3][3], clipDist[3], centerD[3], slCenter[3];
    float stepSize, scaleVol[3], texMax[3], camera[3], slV[3], slD;
    int width, height, numIter, nClip, rank, world, nLocalBlocks, blockSkew, partUnit, slicing, slNum, depthMajor, bandRows, firstWindow, windowGrowth, sampleMap;
    unsigned long long mcVersion;
    const void *mcOffsets;
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

struct VVRenderer {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = 148;
    bool inited = false;

    // ---- volume geometry (VolumeData, VV/dataset.cpp:144-176) ----
    int size[3] = {0, 0, 0};
    float slice_dist[3] = {1, 1, 1};
    float extent[3] = {1, 1, 1}, scale[3] = {1, 1, 1}, scale_inv[3] = {1, 1, 1}, center[3] = {.5f, .5f, .5f};

    // ---- vector field ----
    DevBuf<uint8_t> raw0, raw1;        // raw time steps (FLOAT3 or UCHAR3)
    bool have_next = false, field_u8 = false, have_field = false;
    int interp_index = 0, interp_size = 10;   // VV/3DLIC.cpp:705 setInterpolateSize(10)
    bool have_dat = false;                    // field loaded by vv_load_dat: vv_idle can step through its time steps
    VVDatInfo dat;
    VVTimeCursor cursor;
    DevBuf<float4> pack_tmp;
    DevBuf<unsigned int> maxbits;
    DevBuf<uint4> field_pair;
    int field_guard = 1, field_gx = 0, field_row = 0;   // guard cells of the x-pair layout (see pack_field_pass2), x offset, row stride
    float walk_reach = 0.0f;                            // (S + 2) h: how far a walk can get from its ray sample, texture coordinates
    int guard_wanted = 1;                               // guard the current LIC parameters need for the unclamped walk (XF_GUARD)
    DevBuf<float4> field_f4;
    int field_layout = LAYOUT_PAIR;        // layout of the packed field (resolved from field_layout_req by resolve_field_layout)
    int field_layout_req = 3;              // VV_OPT_FIELD_LAYOUT: 0 float4, 1 x-pair, 2 xy-quad, 3 (default) = by size / technique
    bool field_dirty = false;

    // ---- scalar volume ----
    DevBuf<uint2> scalar_cell;
    int sdim[3] = {0, 0, 0};
    bool have_scalar = false;

    // ---- noise ----
    DevBuf<uint8_t> noise_raw;
    DevBuf<uint2> noise_cell;
    DevBuf<uchar4> noise_rgba;
    DevBuf<uint4> noise_quad, noise_pair, noise_bf;
    int nbf_guard = 1, nbf_gx = 1, nbf_row = 0;   // geometry of noise_bf: its own wrapped border, or the vector field's guard geometry
    int noise_layout = 2;                  // RGBA noise: 2 = bf16 {t0, t1 - t0} (default), 1 = fp16 x-pair, 0 = u8 xy-quad (check layouts)
    DevBuf<float> grad_tmp, grad_filter;
    int ndim[3] = {0, 0, 0};
    bool have_noise = false, noise_has_grad = false;

    // ---- filter kernel (LICFilter) / transfer function (TransferEdit) ----
    std::vector<uint8_t> filter;
    float inv_filter_area = 0.5f;
    uint8_t tf[256 * 5];
    DevBuf<float4> tf_rgba;
    DevBuf<float> tf_opac, kw;
    bool tables_dirty = true;
    int n_bwd_eff = 0, n_fwd_eff = 0;      // LIC steps up to the last non-zero kernel weight

    // ---- parameters ----
    VVLicParams lp;
    float cam_quat[4] = {0, 0, 0, 1}, cam_pos[3] = {0, 0, 0}, cam_dist = 4.0f, fovy = 35.0f, near_clip = 0.1f, far_clip = 50.0f;
    float window_aspect = 0.0f;            // Camera::setWindow (VV/transform.h:79-80); 0 = frame width / frame height
    float light_quat[4] = {0, 0, 0, 1}, light_dist = 1.0f;   // VV/3DLIC.cpp:681
    float light_pos[3] = {0.5f, 0.5f, 1.5f};
    int technique = VV_VOLIC_RAYCAST;
    int illum_mode = ILLUM_NONE;
    bool speed_of_flow = false, lowres = false, float_target = false;
    // TF index / LIC gate are hard-coded per shader in the reference (Q5, Q6): [0] ray-cast program (.b, always),
    // [1] slicing program (.a, tfData.a > 0.05); vv_set_option changes the entry of the current technique
    int tf_modes[2] = {TF_B, TF_A}, gate_modes[2] = {GATE_ALWAYS, GATE_TF_ALPHA};
    int noise_gate = 1, quirk_scalevolinv = 1, quirk_lum_alpha = 0;
    int licvol_fp16 = 1, count_samples = 1, licvol_size = 0, sample_map = 0;
    DevBuf<unsigned int> sample_tiles;
    float spec_exp = 40.0f;
    // MC ray-start offsets (USE_MC_OFFSET + Renderer::updateMCOffsetTex) and user clip planes
    bool use_mc = false;
    DevBuf<float> mc_offsets;
    int mc_w = 0, mc_h = 0;
    // peer-to-peer frame exchange (vv_p2p_*): [256 B header: arrival counters of the two parities][tiles parity 0][tiles parity 1]
    unsigned char *p2p_base = nullptr;
    size_t p2p_tile_bytes = 0;
    int p2p_world = 0;
    void *p2p_peer[kMaxPeers] = {};
    bool p2p_opened[kMaxPeers] = {};
    unsigned int p2p_epoch = 0;
    DevBuf<unsigned int> p2p_scratch;      // [0] done counter of the scatter kernel, [1] time-out flag of the wait kernel
    // screenshot / recording (Renderer::renderFBO tail, VV/renderer.cpp:1478-1513; keys VV/3DLIC.cpp:262-270)
    bool screenshot = false, recording = false, animation_on = false;
    int frames = 0;
    std::string snapshot_dir = "snapshotOut", snapshot_name = "snapshot.png", last_snapshot;
    bool clip_active[3] = {false, false, false};
    double clip_eq[3][4] = {{0, 0, -1, 0}, {0, 0, -1, 0}, {0, 0, -1, 0}};   // ClipPlane ctor, VV/transform.cpp:240-254

    // ---- frame ----
    int width = 0, height = 0;
    int rank = 0, world = 1;
    int nbx = 0, nby = 0, n_local_blocks = 0, blocks_per_rank = 0;
    DevBuf<float4> tiles, frame;
    DevBuf<uchar4> frame8, display8;
    DevBuf<unsigned long long> counters;   // as unsigned int[16]: [0,1] ray samples (u64), [2] block queue, [3] src rows,
                                           // [4],[5] item counts (ping-pong), [6] item queue head, [7] max samples per ray,
                                           // [8] first window's item count, [9] march-checkpoint rows
    unsigned int *host_counters = nullptr; // pinned
    // sample-parallel pipeline state
    DevBuf<float4> rayA, rayB, src;
    DevBuf<uint2> tileRec, items[3];       // items[0]: first depth window (kept while the view is unchanged), [1], [2]: later windows
    DevBuf<unsigned int> tileLive, tileCk;
    DevBuf<float4> rayCk;                  // march checkpoints (ray_checkpoint_kernel), sized from counter [9]
    // the view the ray records / src rows / first-window items on the device were computed for (GeomKey), and their sizes
    GeomKey geom_key = {};
    bool geom_valid = false;
    size_t geom_rows = 0;
    int geom_nmax = 0;
    const void *geom_rayA = nullptr;
    unsigned long long mc_version = 0;
    int raycast_mode = 1;                  // 1: sample-parallel pipeline (default), 0: one thread per ray
    int lic_ctas_per_sm = 0;               // 0: as many as are resident (occupancy query)
    int xf_enable = 1;                     // coordinate fast paths of the walk (XF_GUARD / XF_NSHARE): 0 = the clamping samplers (check)
    int first_window = 2, window_growth = 300;    // depth windows of early-termination frames: first length, growth in percent: 2, 6, 18, 54 ...
                                                  // (measured, profiles/r02/ab27_window_schedules_any_length.log: a surface-like frame such as
                                                  // cfg3o pays for every speculative sample -- 1.12 ms with windows 8, 16, 32 ..., 0.40 ms with
                                                  // 2, 6, 18 ...; cfg1, whose rays live ~17 samples, is flat: 1.63 vs 1.65 ms)
    int part_unit = 1;                     // VV_OPT_PARTITION_UNIT: the sort-first partition deals units of this many x this many blocks
    int depth_major = 1;                   // 1: bucket work items by (band, depth chunk) for L2 locality; 0: tile-major
    int band_rows = 4;                     // block rows per band (4 x 16 = 64 pixel rows)
    DevBuf<unsigned int> buckets;
    bool frame_valid = false;
    int launches = 0;

    // ---- LIC volume ----
    DevBuf<float> licvol;
    int ldim[3] = {0, 0, 0};
    int slab_z0 = 0, slab_z1 = -1;
    bool licvol_valid = false;

    // ---- illumination tables (Illumination, VV/illumination.cpp) ----
    DevBuf<float> illum_tab[3];
    int illum_w = 0, illum_h = 0;
};

// ------------------------------------------------------------------------------------------------ helpers

static void volume_geometry(VVRenderer *r)
{
    // VectorDataSet::loadData, VV/dataset.cpp:144-176
    float volSize[3], maxVolSize = 0;
    int maxTexSize = 0;
    for (int i = 0; i < 3; ++i) {
        volSize[i] = r->size[i] * r->slice_dist[i];
        if (volSize[i] > maxVolSize) maxVolSize = volSize[i];
        if (r->size[i] > maxTexSize) maxTexSize = r->size[i];
    }
    for (int i = 0; i < 3; ++i) {
        r->scale[i] = maxTexSize / (r->size[i] * r->slice_dist[i]);
        r->scale_inv[i] = r->size[i] * r->slice_dist[i] / maxTexSize;
        r->extent[i] = r->size[i] * r->slice_dist[i] / maxVolSize;
        r->center[i] = r->extent[i] / 2.0f;
    }
}

// Quaternion_getAngleAxis, VV/mmath.cpp:213-232
static void quat_angle_axis(const float q[4], float *angle, float axis[3])
{
    double d = std::sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2]);
    if (d > 1e-6) {
        d = 1.f / d;
        axis[0] = (float)(q[0] * d); axis[1] = (float)(q[1] * d); axis[2] = (float)(q[2] * d);
        *angle = (1.0 - std::fabs(q[3]) > 1e-6) ? 2.f * (float)std::acos(q[3]) : 0.0f;
    } else {
        axis[0] = 0.f; axis[1] = 0.f; axis[2] = 1.f; *angle = 0.f;
    }
}

// glRotatef(angle, axis) as a row-major 3x3 (GL 2.1 spec 2.11.2)
static void gl_rotation(float angle_rad, const float axis[3], float R[9])
{
    const double deg = (double)(float)(angle_rad * 180.0 / M_PI);   // VV/camera.cpp:67
    const double a = deg * M_PI / 180.0;
    double x = axis[0], y = axis[1], z = axis[2];
    const double n = std::sqrt(x * x + y * y + z * z);
    if (n > 0) { x /= n; y /= n; z /= n; }
    const double c = std::cos(a), s = std::sin(a), t = 1.0 - c;
    const double M[9] = {t * x * x + c,     t * x * y - s * z, t * x * z + s * y,
                         t * x * y + s * z, t * y * y + c,     t * y * z - s * x,
                         t * x * z - s * y, t * y * z + s * x, t * z * z + c};
    for (int i = 0; i < 9; ++i) R[i] = (float)M[i];
}

// Renderer::updateLightPos, VV/renderer.cpp:431-466
static void update_light(VVRenderer *r)
{
    // Transform::getPosition: q * (0,0,dist)
    const double x = r->light_quat[0], y = r->light_quat[1], z = r->light_quat[2], w = r->light_quat[3];
    const double v[3] = {0.0, 0.0, (double)r->light_dist};
    const double tx = 2.0 * (y * v[2] - z * v[1]), ty = 2.0 * (z * v[0] - x * v[2]), tz = 2.0 * (x * v[1] - y * v[0]);
    const float lp[3] = {(float)(v[0] + w * tx + (y * tz - z * ty)), (float)(v[1] + w * ty + (z * tx - x * tz)),
                         (float)(v[2] + w * tz + (x * ty - y * tx))};
    float angle, axis[3], Rm[9];
    quat_angle_axis(r->cam_quat, &angle, axis);
    gl_rotation(-angle, axis, Rm);
    // M = T(center) R(-angle) T(cam_pos)   (no T(0,0,dist): the TODO at renderer.cpp:448)
    const double l[3] = {(double)lp[0] + r->cam_pos[0], (double)lp[1] + r->cam_pos[1], (double)lp[2] + r->cam_pos[2]};
    for (int i = 0; i < 3; ++i)
        r->light_pos[i] = (float)(Rm[3 * i] * l[0] + Rm[3 * i + 1] * l[1] + Rm[3 * i + 2] * l[2] + r->center[i]);
}

static int part_unit(const VVRenderer *r) { return r->world > 1 ? r->part_unit : 1; }

static int ensure_frame(VVRenderer *r)
{
    r->nbx = (r->width + kBlockDim - 1) / kBlockDim;
    r->nby = (r->height + kBlockDim - 1) / kBlockDim;
    // units of U x U blocks are dealt to the ranks (block_id_u); a single GPU keeps plain blocks
    const int U = part_unit(r);
    const int nunits = ((r->nbx + U - 1) / U) * ((r->nby + U - 1) / U);
    r->blocks_per_rank = (nunits + r->world - 1) / r->world * U * U;
    r->n_local_blocks = std::max(0, (nunits - r->rank + r->world - 1) / r->world) * U * U;
    CU(r->tiles.ensure((size_t)r->blocks_per_rank * kBlockPixels));
    const size_t npx = (size_t)r->width * r->height;
    CU(r->frame.ensure(npx));
    CU(r->frame8.ensure(npx));
    CU(r->display8.ensure(npx));
    CU(r->counters.ensure(kNumCounters / 2));
    return VV_OK;
}

// LICFilter texture semantics: LUMINANCE8, LINEAR, GL_CLAMP with border 0 (VV/dataset.cpp:1489-1499)
static float kernel_lookup(const VVRenderer *r, float s)
{
    const int n = (int)r->filter.size();
    s = std::fmin(std::fmax(s, 0.0f), 1.0f);
    const float u = s * (float)n - 0.5f;
    const float fl = std::floor(u);
    const float f = u - fl;
    const int i0 = (int)fl, i1 = i0 + 1;
    const float t0 = (i0 < 0 || i0 >= n) ? 0.0f : (float)r->filter[i0] / 255.0f;
    const float t1 = (i1 < 0 || i1 >= n) ? 0.0f : (float)r->filter[i1] / 255.0f;
    return (1.0f - f) * t0 + f * t1;
}

struct Uniforms {
    float stepSize, gradient[3], licParams[3], licKernel[3], alphaCorrection;
    int nFwd, nBwd;
};

// Renderer::setRenderVolParams, VV/renderer.cpp:947-995
static Uniforms derive_uniforms(const VVRenderer *r)
{
    Uniforms u;
    const VVLicParams &p = r->lp;
    if (r->lowres) {
        u.stepSize = 2.0f * p.stepSizeVol;
        u.gradient[0] = p.gradientScale; u.gradient[1] = p.illumScale; u.gradient[2] = 0.7f * p.freqScale;
        u.licParams[0] = 15.0f; u.licParams[1] = 15.0f; u.licParams[2] = 1.0f / 64.0f;
        u.licKernel[0] = 0.5f / 15.0f; u.licKernel[1] = 0.5f / 15.0f; u.licKernel[2] = r->inv_filter_area / (30.0f);
        u.alphaCorrection = 2.0f * p.stepSizeVol * 128.0f;
    } else {
        u.stepSize = p.stepSizeVol;
        u.gradient[0] = p.gradientScale; u.gradient[1] = p.illumScale; u.gradient[2] = p.freqScale;
        u.licParams[0] = (float)p.stepsForward; u.licParams[1] = (float)p.stepsBackward; u.licParams[2] = p.stepSizeLIC;
        u.licKernel[0] = 0.5f / p.stepsForward; u.licKernel[1] = 0.5f / p.stepsBackward;
        u.licKernel[2] = r->inv_filter_area / (p.stepsForward + p.stepsBackward);
        u.alphaCorrection = p.stepSizeVol * 128.0f;
    }
    u.nFwd = (int)u.licParams[0];   // int(licParams.x), inc_lic.glsl:192
    u.nBwd = (int)u.licParams[1];
    return u;
}

static int upload_tables(VVRenderer *r, const Uniforms &u)
{
    if (r->filter.empty()) return fail(VV_ERR_STATE, "no LIC filter kernel set");
    if (u.nFwd < 0 || u.nBwd < 0 || u.nFwd > kMaxLicSteps || u.nBwd > kMaxLicSteps)
        return fail(VV_ERR_INVALID, "LIC steps out of range (0..1024 per direction)");
    // per-step filter-kernel weights: texture1D(licKernelSampler, kernelOffset).r with the offsets of
    // computeLIC (inc_lic.glsl:156,172,182,196)
    std::vector<float> kw(1 + u.nBwd + u.nFwd);
    kw[0] = kernel_lookup(r, 0.5f);
    float off = 0.5f;
    for (int i = 0; i < u.nBwd; ++i) { off -= u.licKernel[1]; kw[1 + i] = kernel_lookup(r, off); }
    off = 0.5f;
    for (int i = 0; i < u.nFwd; ++i) { off += u.licKernel[0]; kw[1 + u.nBwd + i] = kernel_lookup(r, off); }
    r->n_bwd_eff = u.nBwd;
    while (r->n_bwd_eff > 0 && kw[r->n_bwd_eff] == 0.0f) --r->n_bwd_eff;
    r->n_fwd_eff = u.nFwd;
    while (r->n_fwd_eff > 0 && kw[u.nBwd + r->n_fwd_eff] == 0.0f) --r->n_fwd_eff;
    CU(r->kw.ensure(2 * kMaxLicSteps + 1));
    CU(cudaMemcpyAsync(r->kw.p, kw.data(), kw.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    // TF textures: RGBA8 and LUMINANCE_ALPHA8 (.a = LIC opacity), VV/transferEdit.cpp:505-540
    std::vector<float> rgba(256 * 4), opac(256);
    for (int i = 0; i < 256; ++i) {
        for (int k = 0; k < 4; ++k) rgba[4 * i + k] = (float)r->tf[5 * i + k] / 255.0f;
        opac[i] = (float)r->tf[5 * i + 4] / 255.0f;
    }
    CU(r->tf_rgba.ensure(256));
    CU(r->tf_opac.ensure(256));
    CU(cudaMemcpyAsync(r->tf_rgba.p, rgba.data(), rgba.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    CU(cudaMemcpyAsync(r->tf_opac.p, opac.data(), opac.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    CU(cudaStreamSynchronize(r->stream));   // host vectors go out of scope
    return VV_OK;
}

static int pack_field(VVRenderer *r)
{
    if (!r->have_field) return fail(VV_ERR_STATE, "no vector field set");
    const size_t n = (size_t)r->size[0] * r->size[1] * r->size[2];
    CU(r->pack_tmp.ensure(n));
    CU(r->maxbits.ensure(1));
    uint4 *pair = nullptr;
    float4 *f4 = nullptr;
    const bool cells = r->field_layout == LAYOUT_PAIR || r->field_layout == LAYOUT_QUAD;
    const size_t per_cell = r->field_layout == LAYOUT_QUAD ? 2 : 1;     // uint4 per cell
    if (cells) {
        // guard cells: what the unclamped walk needs (guard_wanted, from the LIC parameters), as long as the padded array stays
        // addressable with signed 32-bit element offsets; else the minimum and the clamping samplers
        int g = std::max(1, r->guard_wanted);
        auto padded = [&](int gg) {
            const size_t gx = gg > 1 ? (size_t)((gg + 7) & ~7) : 0;
            return (((size_t)r->size[0] + 2 * gx + 7) & ~(size_t)7) * ((size_t)r->size[1] + 2 * gg) * ((size_t)r->size[2] + 2 * gg);
        };
        if (padded(g) >= ((size_t)1 << 31)) g = 1;
        if (padded(g) >= ((size_t)1 << 31)) return fail(VV_ERR_INVALID, "vector field too large for 32-bit element offsets");
        r->field_guard = g;
        r->field_gx = g > 1 ? ((g + 7) & ~7) : 0;               // x guard in whole 128-byte lines: cell x = 0 stays line-aligned
        r->field_row = (r->size[0] + 2 * r->field_gx + 7) & ~7;
        CU(r->field_pair.ensure(padded(g) * per_cell));
        pair = r->field_pair.p;
    }
    else { CU(r->field_f4.ensure(n)); f4 = r->field_f4.p; }
    const float frac = (float)r->interp_index / r->interp_size;   // VV/dataset.cpp:590
    CU(launch_pack_field(r->raw0.p, r->have_next ? r->raw1.p : nullptr, r->field_u8 ? 1 : 0, r->size[0], r->size[1], r->size[2],
                         frac, r->pack_tmp.p, r->maxbits.p, pair, r->field_guard, r->field_gx, r->field_row, r->field_layout == LAYOUT_QUAD ? 1 : 0, f4, r->stream));
    r->field_dirty = false;
    return VV_OK;
}

// VV_OPT_FIELD_LAYOUT = 3: the xy-quad layout (32 B/voxel, one 256-bit load per cell face) wherever it fits comfortably: measured
// against the x-pair layout (16 B/voxel), frames and volumes bit-identical -- lic_sample_kernel 2.5 % faster on cfg3 and cfg2, 2 % on
// cfg4 (512^3) (profiles/r02/ab13_quad_layout.log, ab15_split_loops_per_build_shape.log); lic_volume_kernel equal at 512^3 and 9 %
// faster at 1024^3, where the kernel is DRAM-bound and a cell face is one fully used 32-byte sector instead of two half-used ones
// (profiles/r02/licvol16_layouts.log, licvol17_layouts_1024.log).  Above 48 GiB of packed field (~1150^3) the x-pair layout.
static void resolve_field_layout(VVRenderer *r)
{
    int want = r->field_layout_req;
    if (want == 3) {
        const size_t n = (size_t)r->size[0] * r->size[1] * r->size[2];
        want = (n * 32 <= ((size_t)48 << 30)) ? LAYOUT_QUAD : LAYOUT_PAIR;
    }
    if (want != r->field_layout) {
        r->field_layout = want;
        r->field_pair.release();
        r->field_f4.release();
        r->field_dirty = true;
    }
}

// raycast_program: the LIC ray-cast program, the only one where scaleVolInv can be an active uniform (Q1)
static int fill_params(VVRenderer *r, DevParams &P, bool need_frame, bool raycast_program = true)
{
    if (!r->have_field) return fail(VV_ERR_STATE, "no vector field set (vv_set_vector_field / vv_load_dat)");
    const Uniforms u = derive_uniforms(r);
    resolve_field_layout(r);
    // Guard band of the unclamped walk (XF_GUARD): a Heun step moves a position by at most h per axis in texture coordinates
    // (|2 v - 1| <= 1 per component for texels in [0,1]) and its predictor looks one more step ahead, so a walk of S steps stays
    // within (S + 1) h of its ray sample.  Ray samples lie in [0, max(texMax, extent * scaleVol)] per axis -- inside [0,1] for a
    // cubic volume, beyond it for anisotropic ones (texMax = maxTexSize / maxVolSize, and Q1 puts scaleInv into scaleVol).
    // + 3 cells of slack (rounding, the +1 neighbour).
    {
        double reach = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double pos_max = std::max((double)(r->extent[i] * r->scale[i]), (double)(r->extent[i] * r->scale_inv[i]));
            const double walk = (double)(std::max(u.nFwd, u.nBwd) + 1) * (double)(u.licParams[2] * 0.3f);
            reach = std::max(reach, (walk + std::max(0.0, pos_max - 1.0)) * (double)r->size[i]);
        }
        const int need = (reach < 60.0) ? (((int)std::ceil(reach) + 3 + 3) & ~3) : 1;      // capped: beyond 64 cells the clamping samplers run
        r->guard_wanted = std::max(need, 1);
        r->walk_reach = (float)((double)(std::max(u.nFwd, u.nBwd) + 2) * (double)(u.licParams[2] * 0.3f));
        if (r->field_layout != LAYOUT_F4 && r->field_pair.p && !r->field_dirty && r->field_guard < r->guard_wanted && need > 1)
            r->field_dirty = true;      // the LIC parameters outgrew the packed guard band: pack again with the larger one
    }
    if (r->field_dirty || (r->field_layout != LAYOUT_F4 ? !r->field_pair.p : !r->field_f4.p)) {
        int rc = pack_field(r);
        if (rc) return rc;
    }
    const bool grad = (r->illum_mode == ILLUM_GRADIENT);
    if (!r->have_noise) return fail(VV_ERR_STATE, "no noise set (vv_set_noise / vv_load_noise / vv_generate_white_noise)");
    if (grad && !r->noise_has_grad) return fail(VV_ERR_STATE, "ILLUM_GRADIENT needs noise gradients (-g)");
    if (!grad && r->noise_gate && !r->have_scalar) return fail(VV_ERR_STATE, "no scalar volume set (vv_set_scalar); the noise gate needs it");
    const int prog = (r->technique == VV_VOLIC_SLICING) ? 1 : 0;
    if (r->tf_modes[prog] == TF_SCALAR && !r->have_scalar) return fail(VV_ERR_STATE, "tf_mode scalar needs a scalar volume");
    if (need_frame && (r->width <= 0 || r->height <= 0)) return fail(VV_ERR_STATE, "vv_resize not called");

    if (r->tables_dirty) {
        int rc = upload_tables(r, u);
        if (rc) return rc;
        r->tables_dirty = false;
    }
    std::memset(&P, 0, sizeof(P));
    P.field_f4 = r->field_f4.p;
    P.fnx = r->size[0]; P.fny = r->size[1]; P.fnz = r->size[2];
    P.fRow = (unsigned int)r->field_row; P.fPlane = (unsigned int)r->field_row * (unsigned int)(r->size[1] + 2 * r->field_guard);
    P.field_pair = r->field_pair.p ? r->field_pair.p + ((size_t)r->field_guard * P.fPlane + (size_t)r->field_guard * P.fRow + r->field_gx) *
                                                           (r->field_layout == LAYOUT_QUAD ? 2 : 1) : nullptr;
    P.fGuard = r->field_guard;
    P.walkReach = r->walk_reach;
    P.guardOk = (r->field_guard > 1 && r->field_guard >= r->guard_wanted && r->xf_enable) ? 1 : 0;
    P.noiseSameDims = (r->ndim[0] == r->size[0] && r->ndim[1] == r->size[1] && r->ndim[2] == r->size[2]) ? 1 : 0;
    P.scalar_cell = r->scalar_cell.p; P.snx = r->sdim[0]; P.sny = r->sdim[1]; P.snz = r->sdim[2];
    P.nnx = r->ndim[0]; P.nny = r->ndim[1]; P.nnz = r->ndim[2];
    for (int k = 0; k < 3; ++k) {
        P.fnf[k] = (float)r->size[k]; P.fnm1f[k] = (float)(r->size[k] - 1);
        P.snf[k] = (float)r->sdim[k]; P.snm1f[k] = (float)(r->sdim[k] - 1);
        P.nnf[k] = (float)r->ndim[k];
    }
    // wrapped-border layouts: point at cell (0,0,0), i.e. one padded plane + row + element in
    P.ncRow = r->ndim[0] + 1; P.ncPlane = P.ncRow * (r->ndim[1] + 1);
    P.npRow = r->ndim[0] + 1; P.npPlane = P.npRow * (r->ndim[1] + 2);
    P.noise_cell = r->noise_cell.p ? r->noise_cell.p + (P.ncPlane + P.ncRow + 1) : nullptr;
    P.noise_quad = r->noise_quad.p;
    P.noise_pair = (r->noise_layout == 1 && r->noise_pair.p) ? r->noise_pair.p + (P.npPlane + P.npRow + 1) : nullptr;
    P.noise_bf = nullptr;
    P.noiseShared = 0;
    if (r->noise_layout == 2 && r->noise_has_grad && grad) {
        // The hot RGBA-noise layout takes the vector field's guard geometry when both volumes have the same dimensions: one cell
        // index then addresses both arrays (XF_NSHARE).  Built from the RGBA texels on first use and whenever that geometry moved.
        const bool share = P.guardOk && r->field_layout != LAYOUT_F4 && P.noiseSameDims && grad;
        const int g = share ? r->field_guard : 1, gx = share ? r->field_gx : 1, row = share ? r->field_row : r->ndim[0] + 1;
        if (g != r->nbf_guard || gx != r->nbf_gx || row != r->nbf_row || !r->noise_bf.p) {
            CU(r->noise_bf.ensure((size_t)row * (r->ndim[1] + 2 * g) * (r->ndim[2] + 2 * g)));
            CU(launch_build_noise_pair(r->noise_rgba.p, r->ndim[0], r->ndim[1], r->ndim[2], g, gx, row, r->noise_bf.p, 1, r->stream));
            r->nbf_guard = g; r->nbf_gx = gx; r->nbf_row = row;
        }
        P.nbRow = row; P.nbPlane = row * (r->ndim[1] + 2 * g);
        P.noise_bf = r->noise_bf.p + ((size_t)g * P.nbPlane + (size_t)g * P.nbRow + gx);
        P.noiseShared = share ? 1 : 0;
    }
    P.licvol = r->licvol.p; P.lnx = r->ldim[0]; P.lny = r->ldim[1]; P.lnz = r->ldim[2];
    P.tf_rgba = r->tf_rgba.p; P.tf_opac = r->tf_opac.p; P.kw = r->kw.p;
    for (int i = 0; i < 3; ++i) P.illum2d[i] = r->illum_tab[i].p;
    P.illum_w = r->illum_w; P.illum_h = r->illum_h;
    P.stepSize = u.stepSize; P.gradScale = u.gradient[0]; P.illumScale = u.gradient[1]; P.freq = u.gradient[2];
    P.h = u.licParams[2] * (0.0f * 0.5f + 0.3f);          // logEyeDist = 0 (Q3), inc_lic.glsl:114
    P.licScale = u.licKernel[2] * u.gradient[0];          // lic3d_fragment.glsl:67
    P.alphaCorr = u.alphaCorrection; P.specExp = r->spec_exp;
    P.numIter = r->lp.numIterations; P.nFwd = u.nFwd; P.nBwd = u.nBwd;
    P.nFwdEff = r->n_fwd_eff; P.nBwdEff = r->n_bwd_eff;
    // VV/renderer.cpp:934-944 incl. Q1
    const bool inv_active = raycast_program && (r->illum_mode != ILLUM_NONE);
    for (int i = 0; i < 3; ++i) {
        P.texMax[i] = r->extent[i] * r->scale[i];
        if (inv_active && r->quirk_scalevolinv) { P.scaleVol[i] = r->scale_inv[i]; P.scaleVolInv[i] = 0.0f; }
        else { P.scaleVol[i] = r->scale[i]; P.scaleVolInv[i] = r->scale_inv[i]; }
        P.lightPos[i] = r->light_pos[i];
        P.extent[i] = r->extent[i];
    }
    // view: Camera::setCamera (VV/camera.cpp:56-68) followed by glTranslatef(-center) (VV/renderer.cpp:146)
    float angle, axis[3], R[9];
    quat_angle_axis(r->cam_quat, &angle, axis);
    gl_rotation(angle, axis, R);
    const double t[3] = {-(double)r->cam_pos[0], -(double)r->cam_pos[1], (double)r->cam_dist - (double)r->cam_pos[2]};
    for (int i = 0; i < 3; ++i) {
        const double c = (double)r->center[i] + (double)R[i] * t[0] + (double)R[3 + i] * t[1] + (double)R[6 + i] * t[2];
        P.camera[i] = (float)c;
        P.camD[i] = (double)P.camera[i];
    }
    for (int i = 0; i < 9; ++i) P.rot[i] = (double)R[i];
    P.tanHalf = (double)(float)std::tan((double)r->fovy * M_PI / 360.0);
    P.nearD = (double)r->near_clip; P.farD = (double)r->far_clip;
    P.aspect = (r->window_aspect > 0.0f) ? (double)r->window_aspect : (r->height > 0) ? (double)((float)r->width / (float)r->height) : 1.0;
    P.width = r->width; P.height = r->height;
    P.slicing = 0;
    if (r->technique == VV_VOLIC_SLICING && need_frame) {
        // ViewSlicing::setupSlicing (VV/slicing.cpp:42-103) on the model-view of Renderer::updateSlices
        // (VV/renderer.cpp:1270-1292): m[2],m[6],m[10] = third row of R, m[14] = z translation, m[15] = 1
        const double tz = ((double)r->cam_pos[2] - (double)r->cam_dist) -
                          ((double)R[6] * r->center[0] + (double)R[7] * r->center[1] + (double)R[8] * r->center[2]);
        const float m14 = (float)tz;
        float inv = 1.0f / (m14 - 1.0f);
        float v[3] = {(R[6] - 0.0f) * inv, (R[7] - 0.0f) * inv, (R[8] - 0.0f) * inv};
        inv = 1.0f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
        const float xMax = 0.5f * r->extent[0], yMax = 0.5f * r->extent[1], zMax = 0.5f * r->extent[2];
        const float av[3] = {std::fabs(v[0]), std::fabs(v[1]), std::fabs(v[2])};
        const float dv[7] = {av[0] * xMax, av[1] * yMax, av[2] * zMax, av[0] * xMax + av[1] * yMax, av[0] * xMax + av[2] * zMax,
                             av[1] * yMax + av[2] * zMax, av[0] * xMax + av[1] * yMax + av[2] * zMax};
        float d = dv[6];
        for (int i = 0; i < 6; ++i) if (d < dv[i]) d = dv[i];
        d *= 2.0;
        // key 'F' (enableFBO): with the FBO the two RGBA16F ping-pong targets and lic3d_slicing_fragment.glsl; without it -- the
        // reference's start-up state -- lic3d_slicingblend_fragment.glsl blended into the RGBA8 back buffer (VV/renderer.cpp:1138-1160)
        P.slicing = r->float_target ? 1 : 2;
        P.slV[0] = v[0]; P.slV[1] = v[1]; P.slV[2] = v[2];
        P.slD = d;
        P.slNum = (int)(d / u.stepSize) + 1;
        for (int i = 0; i < 3; ++i) P.slCenter[i] = (double)r->center[i];
    }
    P.tfMode = r->tf_modes[prog]; P.gateMode = r->gate_modes[prog]; P.quirkLumAlpha = (r->quirk_lum_alpha && !r->noise_has_grad) ? 1 : 0;   // Q7 only bites GL_LUMINANCE noise
    if (r->use_mc && r->mc_offsets.p && need_frame) {
        if (r->mc_w != r->width || r->mc_h != r->height) return fail(VV_ERR_STATE, "MC offset texture size differs from the frame (vv_set_mc_offsets after vv_resize)");
        P.mcOffsets = r->mc_offsets.p;
    }
    for (int k = 0; k < 3; ++k) P.centerD[k] = (double)r->center[k];
    P.nClip = 0;
    for (int i = 0; i < 3; ++i) {
        if (!r->clip_active[i]) continue;
        const double *e = r->clip_eq[i];
        const int j = P.nClip++;
        // The reference normalises the plane's normal IN PLACE the first time its cap is drawn (drawClippedPolygon hands
        // ClipPlane::getNormal() to ViewSlicing::setupSingleSlice, VV/renderer.cpp:1301, VV/slicing.cpp:337-348; |n| <= VS_EPS
        // zeroes it), d untouched: from the second frame on the GL plane is (n / |n|, d).  That steady state is what runs here.
        const double len = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        for (int k = 0; k < 3; ++k) P.clipEq[j][k] = (len > 1e-8) ? e[k] / len : 0.0;
        P.clipEq[j][3] = e[3];
        if (len > 1e-8) {
            for (int k = 0; k < 3; ++k) P.clipN[j][k] = e[k] / len;
            P.clipDist[j] = -(e[3] - 0.0001);               // ClipPlane::drawSlice, VV/transform.cpp:432-444
        } else {
            P.clipDist[j] = std::nan("");
        }
    }
    P.rank = r->rank; P.world = r->world; P.nBlocksX = r->nbx; P.nBlocksY = r->nby; P.nLocalBlocks = r->n_local_blocks;
    P.blockSkew = block_skew_for(r->world);
    P.partUnit = part_unit(r);
    P.tiles = r->tiles.p;
    P.sampleCounter = r->count_samples ? r->counters.p : nullptr;
    P.blockCounter = r->counters.p ? reinterpret_cast<unsigned int *>(r->counters.p + 1) : nullptr;
    P.samplesPerPixel = nullptr;
    if (r->sample_map && need_frame) {
        CU(r->sample_tiles.ensure((size_t)r->blocks_per_rank * kBlockPixels));
        P.samplesPerPixel = r->sample_tiles.p;
    }
    P.licvolFp16 = r->licvol_fp16;
    return VV_OK;
}

static int persistent_grid(const VVRenderer *r, int work_items)
{
    int g = r->num_sms * 2;
    if (work_items > 0 && g > work_items) g = work_items;
    return g < 1 ? 1 : g;
}

static int run_unblock(VVRenderer *r, const float4 *tiles, int world, int blocks_per_rank)
{
    CU(launch_unblock(tiles, world, blocks_per_rank, r->nbx, r->nby, block_skew_for(world), world > 1 ? r->part_unit : 1, r->width, r->height, r->frame.p, r->frame8.p, r->display8.p, r->stream));
    ++r->launches;
    r->frame_valid = true;
    return VV_OK;
}

static int compute_lic_volume(VVRenderer *r)
{
    int n = r->licvol_size > 0 ? r->licvol_size : 0;
    int dims[3] = {n ? n : r->size[0], n ? n : r->size[1], n ? n : r->size[2]};
    DevParams P;
    int rc = fill_params(r, P, false, false);
    if (rc) return rc;
    const size_t cnt = (size_t)dims[0] * dims[1] * dims[2];
    if (r->licvol.n < cnt || r->ldim[0] != dims[0] || r->ldim[1] != dims[1] || r->ldim[2] != dims[2]) {
        CU(r->licvol.ensure(cnt));
        CU(cudaMemsetAsync(r->licvol.p, 0, cnt * sizeof(float), r->stream));
    }
    r->ldim[0] = dims[0]; r->ldim[1] = dims[1]; r->ldim[2] = dims[2];
    P.licvol_out = r->licvol.p;
    P.ow = dims[0]; P.oh = dims[1]; P.od = dims[2];
    P.oz0 = (r->slab_z1 >= 0) ? r->slab_z0 : 0;
    P.oz1 = (r->slab_z1 >= 0) ? r->slab_z1 : dims[2];
    if (P.oz0 < 0 || P.oz1 > dims[2] || P.oz0 > P.oz1) return fail(VV_ERR_INVALID, "LIC-volume slab out of range");
    const long long nblocks = (long long)((dims[0] + 7) / 8) * ((dims[1] + 7) / 8) * ((P.oz1 - P.oz0 + 3) / 4);
    if (nblocks > 0) {
        const int grid = (int)std::min<long long>(nblocks, (long long)r->num_sms * 16);
        const bool grad = (r->illum_mode == ILLUM_GRADIENT);
        CU(cudaEventRecord(r->ev0, r->stream));
        CU(launch_lic_volume(P, r->field_layout, grad, !grad && r->noise_gate, r->speed_of_flow, grid, r->stream));
        CU(cudaEventRecord(r->ev1, r->stream));
        ++r->launches;
    }
    r->licvol_valid = true;
    return VV_OK;
}

#ifdef VV_HAVE_ILLUM_TABLES
// CPU generation of the Zoeckler / Mallo tables is in vv_illum.cpp
namespace vvb200 {
void make_illum_tables(int w, int h, float spec_exp, std::vector<float> &zoeckler, std::vector<float> &mallo_diff, std::vector<float> &mallo_spec);
}

static int ensure_illum_tables(VVRenderer *r)
{
    if (r->illum_tab[0].p) return VV_OK;
    std::vector<float> z, md, ms;
    const int w = 256, h = 256;
    make_illum_tables(w, h, r->spec_exp, z, md, ms);
    CU(r->illum_tab[0].ensure(z.size()));
    CU(r->illum_tab[1].ensure(md.size()));
    CU(r->illum_tab[2].ensure(ms.size()));
    CU(cudaMemcpy(r->illum_tab[0].p, z.data(), z.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r->illum_tab[1].p, md.data(), md.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(r->illum_tab[2].p, ms.data(), ms.size() * sizeof(float), cudaMemcpyHostToDevice));
    r->illum_w = w; r->illum_h = h;
    return VV_OK;
}
#endif

// grad3: already quantised gradients [z][y][x][3] (the .grd cache), or null to compute them on the GPU
static int upload_noise(VVRenderer *r, const uint8_t *data, const int dims[3], int with_gradients, const uint8_t *grad3 = nullptr)
{
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    // the padded layouts are addressed with signed 32-bit element offsets
    if (n == 0 || (size_t)(dims[0] + 1) * (dims[1] + 2) * (dims[2] + 2) >= ((size_t)1 << 31)) return fail(VV_ERR_INVALID, "noise dimensions out of range");
    CU(r->noise_raw.ensure(n));
    CU(cudaMemcpyAsync(r->noise_raw.p, data, n, cudaMemcpyHostToDevice, r->stream));
    r->ndim[0] = dims[0]; r->ndim[1] = dims[1]; r->ndim[2] = dims[2];
    // REPEAT layouts carry a wrapped border (cell -1 on every axis), see vv_device.cuh
    CU(r->noise_cell.ensure((size_t)(dims[0] + 1) * (dims[1] + 1) * (dims[2] + 1)));
    CU(launch_build_cell8(r->noise_raw.p, 1, 0, dims[0], dims[1], dims[2], 1, 1, r->noise_cell.p, r->stream));
    r->noise_has_grad = false;
    if (with_gradients && grad3) {
        // NoiseDataSet::createTexture with stored gradients: pack (gradient.xyz, noise), VV/dataset.cpp:1269-1282
        std::vector<uint8_t> rgba(4 * n);
        for (size_t i = 0; i < n; ++i) {
            rgba[4 * i] = grad3[3 * i]; rgba[4 * i + 1] = grad3[3 * i + 1]; rgba[4 * i + 2] = grad3[3 * i + 2];
            rgba[4 * i + 3] = data[i];
        }
        CU(r->noise_rgba.ensure(n));
        CU(r->noise_quad.ensure(n));
        CU(cudaMemcpyAsync(r->noise_rgba.p, rgba.data(), 4 * n, cudaMemcpyHostToDevice, r->stream));
        CU(cudaStreamSynchronize(r->stream));   // rgba goes out of scope
    } else if (with_gradients) {
        // filter table of filterGradients, VV/gradient.cpp:392-409 (only the k,j,i in -2..0 corner is ever written; Q16)
        float filt[125];
        std::memset(filt, 0, sizeof(filt));
        const int fw = 2;
        float sum = 0.0f;
        for (int k = -fw; k < fw - 1; ++k)
            for (int j = -fw; j < fw - 1; ++j)
                for (int i = -fw; i < fw - 1; ++i)
                    sum += filt[((fw + k) * 5 + fw + j) * 5 + fw + i] = std::exp(-(float)(i * i + j * j + k * k) / 5.0f);
        for (int k = -fw; k < fw - 1; ++k)
            for (int j = -fw; j < fw - 1; ++j)
                for (int i = -fw; i < fw - 1; ++i) filt[((fw + k) * 5 + fw + j) * 5 + fw + i] /= sum;
        CU(r->grad_filter.ensure(125));
        CU(cudaMemcpyAsync(r->grad_filter.p, filt, sizeof(filt), cudaMemcpyHostToDevice, r->stream));
        CU(r->grad_tmp.ensure(3 * n));
        CU(r->noise_rgba.ensure(n));
        CU(r->noise_quad.ensure(n));
        // NoiseDataSet keeps sliceDist = 1 (never set for noise files, VV/dataset.cpp:1124-1173 / VolumeData ctor)
        const float sd[3] = {1.0f, 1.0f, 1.0f};
        CU(launch_noise_gradients(r->noise_raw.p, dims[0], dims[1], dims[2], sd, r->grad_filter.p, r->grad_tmp.p, r->noise_rgba.p, r->stream));
    }
    if (with_gradients) {
        CU(launch_build_quad(r->noise_rgba.p, dims[0], dims[1], dims[2], r->noise_quad.p, r->stream));
        CU(r->noise_pair.ensure((size_t)(dims[0] + 1) * (dims[1] + 2) * (dims[2] + 2)));
        CU(launch_build_noise_pair(r->noise_rgba.p, dims[0], dims[1], dims[2], 1, 1, dims[0] + 1, r->noise_pair.p, 0, r->stream));
        r->nbf_row = 0;                   // the hot bf16 layout is built by fill_params in the geometry the frame's field layout asks for
        r->noise_has_grad = true;
    }
    CU(cudaStreamSynchronize(r->stream));
    r->grad_tmp.release();
    r->have_noise = true;
    r->frame_valid = false;
    return VV_OK;
}

// Can a ray sample ever reach src.a > 0.95 (the shader's early-termination test, lic3d_fragment.glsl:91)?  src.a =
// 1 - (1 - opac * tf.a)^alphaCorrection with both factors linearly interpolated table entries, so the table maxima
// bound it.  If not, one depth window covers the whole ray and nothing is computed speculatively.
static bool termination_possible(const VVRenderer *r, const Uniforms &u)
{
    int amax = 0, omax = 0;
    for (int i = 0; i < 256; ++i) { amax = std::max<int>(amax, r->tf[5 * i + 3]); omax = std::max<int>(omax, r->tf[5 * i + 4]); }
    const double a = (amax / 255.0) * (omax / 255.0);
    const double corrected = 1.0 - std::pow(1.0 - a, (double)u.alphaCorrection);
    return corrected > 0.94;   // margin for fp32 rounding
}

static void make_geom_key(const VVRenderer *r, const DevParams &P, int first_window, GeomKey &k)
{
    std::memset(&k, 0, sizeof(k));
    std::memcpy(k.camD, P.camD, sizeof(k.camD)); std::memcpy(k.rot, P.rot, sizeof(k.rot));
    k.tanHalf = P.tanHalf; k.aspect = P.aspect; std::memcpy(k.extent, P.extent, sizeof(k.extent)); k.nearD = P.nearD; k.farD = P.farD;
    std::memcpy(k.clipEq, P.clipEq, sizeof(k.clipEq)); std::memcpy(k.clipN, P.clipN, sizeof(k.clipN));
    for (int i = 0; i < 3; ++i) k.clipDist[i] = (i < P.nClip && P.clipDist[i] == P.clipDist[i]) ? P.clipDist[i] : -1e300;   // NaN-free
    std::memcpy(k.centerD, P.centerD, sizeof(k.centerD)); std::memcpy(k.slCenter, P.slCenter, sizeof(k.slCenter));
    k.stepSize = P.stepSize;
    for (int i = 0; i < 3; ++i) { k.scaleVol[i] = P.scaleVol[i]; k.texMax[i] = P.texMax[i]; k.camera[i] = P.camera[i]; k.slV[i] = P.slV[i]; }
    k.slD = P.slD;
    k.width = P.width; k.height = P.height; k.numIter = P.numIter; k.nClip = P.nClip; k.rank = P.rank; k.world = P.world;
    k.nLocalBlocks = P.nLocalBlocks; k.blockSkew = P.blockSkew; k.partUnit = P.partUnit; k.slicing = P.slicing; k.slNum = P.slNum;
    k.depthMajor = r->depth_major; k.bandRows = r->band_rows; k.firstWindow = first_window; k.windowGrowth = r->window_growth; k.sampleMap = P.samplesPerPixel ? 1 : 0;
    k.mcVersion = r->mc_version; k.mcOffsets = P.mcOffsets;
}

// K1 as three kernels: ray_setup -> [work items -> lic_sample -> composite] per depth window (see vv_kernels.cu)
static int render_sample_parallel(VVRenderer *r, DevParams &P)
{
    const int nTiles = r->n_local_blocks * 8;
    if (nTiles == 0) return VV_OK;
    CU(r->rayA.ensure((size_t)nTiles * 32));
    CU(r->rayB.ensure((size_t)nTiles * 32));
    CU(r->tileRec.ensure((size_t)nTiles));
    CU(r->tileLive.ensure((size_t)nTiles));
    CU(r->tileCk.ensure((size_t)nTiles));
    unsigned int *cnt = reinterpret_cast<unsigned int *>(r->counters.p);
    P.rayA = r->rayA.p; P.rayB = r->rayB.p; P.tileRec = r->tileRec.p; P.tileLive = r->tileLive.p; P.tileCk = r->tileCk.p;
    P.slotAlloc = cnt + 3; P.nMaxGlobal = cnt + 7; P.itemHead = cnt + 6; P.ckAlloc = cnt + 9;
    P.rayCk = r->rayCk.p;
    const int setup_grid = std::max(1, std::min((nTiles + 7) / 8, r->num_sms * 8));
    // depth windows: whole ray at once when no sample can trigger the early termination (and without the FBO nothing stops a slice
    // from being blended), else 2, 6, 18, ... samples (VV_OPT_FIRST_WINDOW / VV_OPT_WINDOW_GROWTH): a window bounds the work done
    // speculatively past a termination
    const bool windowed = P.slicing == 1 || (!P.slicing && termination_possible(r, derive_uniforms(r)));
    const int first_window = windowed ? r->first_window : 0x7fffffff;
    GeomKey key;
    make_geom_key(r, P, first_window, key);
    const bool same_view = r->geom_valid && std::memcmp(&key, &r->geom_key, sizeof(key)) == 0 && r->rayA.p == r->geom_rayA;
    if (!same_view) {
        r->geom_valid = false;
        CU(cudaMemsetAsync(cnt, 0, kNumCounters * sizeof(unsigned int), r->stream));
        CU(launch_ray_setup(P, setup_grid, r->stream));
        ++r->launches;
        // size the src / item buffers from the allocation the set-up made (one small read-back, only when the view changed)
        if (!r->host_counters) CU(cudaMallocHost((void **)&r->host_counters, kNumCounters * sizeof(unsigned int)));
        CU(cudaMemcpyAsync(r->host_counters, cnt, kNumCounters * sizeof(unsigned int), cudaMemcpyDeviceToHost, r->stream));
        CU(cudaStreamSynchronize(r->stream));
        r->geom_rows = r->host_counters[3];
        r->geom_nmax = (int)r->host_counters[7];
        if (!P.slicing && r->host_counters[9] > 0) {
            // every kCkStride-th position of every ray's march (kept with the ray records while the view stays the same)
            CU(r->rayCk.ensure((size_t)r->host_counters[9] * 32));
            P.rayCk = r->rayCk.p;
            CU(launch_ray_checkpoints(P, setup_grid, r->stream));
            ++r->launches;
        }
    } else if (P.slicing) {
        // (the slicing set-up also counts ray samples and paints the white background: it runs again, and so does the item
        // construction below; only the sizes are known already)
        CU(cudaMemsetAsync(cnt, 0, kNumCounters * sizeof(unsigned int), r->stream));
        CU(launch_ray_setup(P, setup_grid, r->stream));
        ++r->launches;
    } else {
        CU(launch_ray_reset(P, cnt, setup_grid, r->stream));
        ++r->launches;
    }
    const size_t rows = r->geom_rows;
    const int nmax = r->geom_nmax;
    if (rows == 0 || nmax == 0) { r->geom_key = key; r->geom_rayA = r->rayA.p; r->geom_valid = true; return VV_OK; }
    if (rows * 512 > ((size_t)64 << 30)) return fail(VV_ERR_INVALID, "frame needs more than 64 GiB of ray-sample buffer (stepSizeVol too small for this frame size)");
    CU(r->src.ensure(rows * 32));
    CU(r->items[0].ensure(rows));
    P.src = r->src.p;
    std::vector<int> w;
    w.push_back(0);
    if (!windowed) w.push_back(nmax);
    else for (int len = first_window; w.back() < nmax; len = std::max(len + 1, len * r->window_growth / 100)) w.push_back(std::min(nmax, w.back() + len));
    w.push_back(w.back());   // sentinel: nothing after the last window
    if (w.size() > 3) { CU(r->items[1].ensure(rows)); CU(r->items[2].ensure(rows)); }
    const int comp_grid = setup_grid;
    const int lic_grid = r->num_sms * r->lic_ctas_per_sm;   // 0: launcher uses the occupancy of the instantiation
    const bool ngate = r->illum_mode != ILLUM_GRADIENT && r->noise_gate;
    const int nBands = (r->nby + r->band_rows - 1) / r->band_rows;
    P.bandRows = r->band_rows;
    // Order of a window's work items.  One window over the whole ray (and the long late windows of an early-termination frame): the
    // depth-major bucket sort, 3 small launches, which keeps the warps in flight inside one slab of the volume.  A SHORT window is a
    // thin slab in depth whatever the order of its items: there composite_kernel emits the next window's items itself, tile-major,
    // and the three launches are saved (measured: cfg1 -2.3 %, cfg3o -12.7 %, profiles/r02/ab29_tile_major_short_windows.log).
    constexpr int kTileMajorMax = 64;
    auto tile_major = [&](size_t p) { return !r->depth_major || (windowed && w[p + 1] - w[p] <= kTileMajorMax); };
    P.emitItems = 0;
    // work items of window p = [w[p], w[p+1]) into `items` / `count`, in (band, depth chunk)-major order (item_bucket_kernel)
    auto build_items = [&](size_t p, uint2 *items, unsigned int *count) -> int {
        P.win1 = w[p]; P.win2 = w[p + 1];
        P.nDepthChunks = (w[p + 1] - w[p] + 7) / 8;
        const int nBuckets = nBands * P.nDepthChunks;
        CU(r->buckets.ensure((size_t)3 * nBuckets));
        CU(cudaMemsetAsync(r->buckets.p, 0, (size_t)3 * nBuckets * sizeof(unsigned int), r->stream));
        P.bucketCount = r->buckets.p; P.bucketBase = r->buckets.p + nBuckets; P.bucketFill = r->buckets.p + 2 * nBuckets;
        P.itemsNext = items; P.itemCountNext = count;
        CU(launch_item_buckets(P, nBuckets, comp_grid, r->stream));
        r->launches += 3;
        return VV_OK;
    };
    if (!same_view || P.slicing) {
        // the first window's items live in their own buffer and counter ([8]): an unchanged view reuses them
        if (!tile_major(0)) {
            int rc = build_items(0, r->items[0].p, cnt + 8);
            if (rc) return rc;
        } else {
            // empty window [0,0): composite_kernel emits the work items of the first real window, tile-major
            P.win0 = 0; P.win1 = 0; P.win2 = w[1];
            P.itemsNext = r->items[0].p; P.itemCountNext = cnt + 8;
            P.sampleCounter = nullptr;
            P.emitItems = 1;
            CU(launch_composite(P, comp_grid, r->stream));
            P.emitItems = 0;
            ++r->launches;
        }
        r->geom_key = key; r->geom_rayA = r->rayA.p; r->geom_valid = true;
    }
    P.sampleCounter = r->count_samples ? r->counters.p : nullptr;
    CU(cudaEventRecord(r->ev0, r->stream));
    for (size_t p = 0; p + 2 < w.size(); ++p) {
        uint2 *items = (p == 0) ? r->items[0].p : r->items[1 + (p & 1)].p;
        unsigned int *count = (p == 0) ? cnt + 8 : cnt + 4 + (p & 1);
        uint2 *items_next = r->items[1 + ((p + 1) & 1)].p;
        unsigned int *count_next = cnt + 4 + ((p + 1) & 1);
        if (p > 0) {
            CU(cudaMemsetAsync(cnt + 6, 0, sizeof(unsigned int), r->stream));          // queue head
            if (!tile_major(p)) {
                int rc = build_items(p, items, count);
                if (rc) return rc;
            }
        }
        P.win0 = w[p]; P.win1 = w[p + 1]; P.win2 = w[p + 2];
        P.items = items; P.itemCount = count;
        P.itemsNext = items_next; P.itemCountNext = count_next;
        P.emitItems = (p + 3 < w.size() && tile_major(p + 1)) ? 1 : 0;      // composite_kernel emits the next window's items
        if (P.emitItems) CU(cudaMemsetAsync(count_next, 0, sizeof(unsigned int), r->stream));   // next window's item count
        CU(launch_lic_sample(P, r->field_layout, r->illum_mode, ngate, r->speed_of_flow, lic_grid, r->stream));
        if (p == 0) CU(cudaEventRecord(r->ev1, r->stream));   // first window = the bulk of the work (all of it in single-window mode)
        CU(launch_composite(P, comp_grid, r->stream));
        r->launches += 2;
    }
    return VV_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char *vv_version(void) { return "vectorvisualization_b200 0.1 (sm_100a)"; }

void vv_default_lic_params(VVLicParams *p)
{
    // LICParams::LICParams, VV/types.h:93-98
    p->stepSizeVol = 1.0f / 128.0f; p->gradientScale = 30.0f; p->illumScale = 1.0f; p->freqScale = 1.0f;
    p->numIterations = 255; p->stepsForward = 32; p->stepsBackward = 32; p->stepSizeLIC = 0.01f;
}

static int create_device_objects(VVRenderer *r)
{
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, r->device));
    r->num_sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
    r->stream = r->own_stream;
    CU(cudaEventCreate(&r->ev0));
    CU(cudaEventCreate(&r->ev1));
    return VV_OK;
}

int vv_create(VVRenderer **out, int cuda_device)
{
    if (!out) return fail(VV_ERR_INVALID, "vv_create: null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VV_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                     " (this library has no CPU fallback)");
    if (cuda_device < 0 || cuda_device >= count) return fail(VV_ERR_INVALID, "vv_create: bad device index");
    CU(cudaSetDevice(cuda_device));
    VVRenderer *r = new VVRenderer();
    r->device = cuda_device;
    const int rc = create_device_objects(r);
    if (rc != VV_OK) { vv_destroy(r); return rc; }       // (the error text set by fail() stays)
    vv_default_lic_params(&r->lp);
    // TransferEdit ctor, VV/transferEdit.cpp:76-82
    for (int i = 0; i < 256; ++i) {
        for (int ch = 0; ch < 3; ++ch) r->tf[5 * i + ch] = (unsigned char)i;
        for (int ch = 3; ch < 5; ++ch) r->tf[5 * i + ch] = (unsigned char)((i < 20) ? 0 : i - 20);
    }
    // LICFilter::createBoxFilter(256), VV/dataset.cpp:1405-1413 (3DLIC.cpp:729 fallback)
    r->filter.assign(256, 255);
    r->inv_filter_area = 0.5f;
    *out = r;
    return VV_OK;
}

static void p2p_close(VVRenderer *r);

void vv_destroy(VVRenderer *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    cudaDeviceSynchronize();
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    if (r->own_stream) cudaStreamDestroy(r->own_stream);
    if (r->host_counters) cudaFreeHost(r->host_counters);
    p2p_close(r);
    if (r->p2p_base) cudaFree(r->p2p_base);
    delete r;
}

// Renderer::loadGLSLShader(defines), VV/renderer.cpp:807-922; define strings from VV/3DLIC.cpp:416-436
int vv_load_glsl_shader(VVRenderer *r, const char *defines)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    int illum = ILLUM_NONE;
    bool sof = false, mc = false;
    if (defines) {
        if (std::strstr(defines, "ILLUM_GRADIENT")) illum = ILLUM_GRADIENT;
        else if (std::strstr(defines, "ILLUM_MALLO")) illum = ILLUM_MALLO;
        else if (std::strstr(defines, "ILLUM_ZOECKLER")) illum = ILLUM_ZOECKLER;
        if (std::strstr(defines, "SPEED_OF_FLOW")) sof = true;
        if (std::strstr(defines, "USE_MC_OFFSET")) mc = true;
        if (std::strstr(defines, "TIME_DEPENDENT")) return fail(VV_ERR_INVALID, "TIME_DEPENDENT builds are not supported");
    }
    if (illum == ILLUM_MALLO || illum == ILLUM_ZOECKLER) {
#ifdef VV_HAVE_ILLUM_TABLES
        CU(cudaSetDevice(r->device));
        int rc = ensure_illum_tables(r);
        if (rc) return rc;
#else
        return fail(VV_ERR_INVALID, "ILLUM_MALLO / ILLUM_ZOECKLER are not built into this library yet");
#endif
    }
    r->illum_mode = illum;
    r->speed_of_flow = sof;
    r->use_mc = mc;
    r->frame_valid = false;
    r->licvol_valid = false;
    return VV_OK;
}

// Renderer::updateMCOffsetTex, VV/renderer.cpp:636-679: one offset in [0,1] per pixel, GL_LUMINANCE16F rectangle texture,
// NEAREST.  Used by the ray-cast and slicing programs when they were built with USE_MC_OFFSET.
int vv_set_mc_offsets(VVRenderer *r, const float *offsets, int width, int height)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    CU(cudaSetDevice(r->device));
    r->frame_valid = false;
    ++r->mc_version;                       // the ray records of the previous frame started from the old offsets
    if (!offsets) { r->mc_offsets.release(); r->mc_w = r->mc_h = 0; return VV_OK; }
    if (width < 1 || height < 1) return fail(VV_ERR_INVALID, "vv_set_mc_offsets: bad size");
    const size_t n = (size_t)width * height;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __half2float(__float2half_rn(offsets[i]));   // the 16F upload rounding
    CU(r->mc_offsets.ensure(n));
    CU(cudaMemcpyAsync(r->mc_offsets.p, h.data(), n * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    r->mc_w = width; r->mc_h = height;
    return VV_OK;
}

int vv_update_mc_offset_tex(VVRenderer *r, int width, int height, uint32_t seed)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (width < 1 || height < 1) return fail(VV_ERR_INVALID, "vv_update_mc_offset_tex: bad size");
    // noise[i] = rand() / RAND_MAX there (unseeded); here mt19937(seed), u = draw / (2^32 - 1) in [0,1]
    std::vector<float> v((size_t)width * height);
    std::mt19937 gen(seed);
    for (size_t i = 0; i < v.size(); ++i) v[i] = (float)((double)gen() / 4294967295.0);
    return vv_set_mc_offsets(r, v.data(), width, height);
}

// ClipPlane::setNormal + setActive (VV/transform.cpp:296-315) for plane index 0..2 (GL_CLIP_PLANE0 + index, VV/3DLIC.cpp:763-781)
int vv_set_clip_plane(VVRenderer *r, int index, const double equation[4], int active)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (index < 0 || index > 2) return fail(VV_ERR_INVALID, "vv_set_clip_plane: index must be 0..2");
    if (equation) for (int k = 0; k < 4; ++k) r->clip_eq[index][k] = equation[k];
    r->clip_active[index] = active != 0;
    r->frame_valid = false;
    return VV_OK;
}

int vv_init(VVRenderer *r, const char *defines)
{
    int rc = vv_load_glsl_shader(r, defines);
    if (rc == VV_OK) r->inited = true;
    return rc;
}

int vv_resize(VVRenderer *r, int width, int height)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (width <= 0 || height <= 0 || width > 32768 || height > 32768) return fail(VV_ERR_INVALID, "vv_resize: bad size");
    CU(cudaSetDevice(r->device));
    r->width = width; r->height = height;
    r->frame_valid = false;
    return ensure_frame(r);
}

int vv_set_technique(VVRenderer *r, int technique)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (technique == VV_VOLIC_VOLUME) return fail(VV_ERR_INVALID, "VOLIC_VOLUME (raw vector-field DVR, F1) is out of scope");
    if (technique < VV_VOLIC_RAYCAST || technique > VV_VOLIC_VOLUMEANI) return fail(VV_ERR_INVALID, "unknown technique");
    r->technique = technique;
    r->frame_valid = false;
    return VV_OK;
}

int vv_update_lic_volume(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    CU(cudaSetDevice(r->device));
    r->launches = 0;
    return compute_lic_volume(r);
}

int vv_update_slices(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    return VV_OK;   // slice set-up is derived per frame from the camera (VV/slicing.cpp:42-103)
}

int vv_set_vector_field(VVRenderer *r, const void *data, const void *next, int dtype, const int dims[3], const float slice_dist[3])
{
    if (!r || !data || !dims) return fail(VV_ERR_INVALID, "vv_set_vector_field: null argument");
    if (dtype != VV_FLOAT && dtype != VV_UCHAR)
        return fail(VV_ERR_INVALID, "VectorData:  Only 8bit integer and 32bit float vectors are supported.");   // dataset.cpp:136-141
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || n > ((size_t)1 << 31)) return fail(VV_ERR_INVALID, "bad field dimensions");
    CU(cudaSetDevice(r->device));
    const size_t bytes = n * 3 * (dtype == VV_FLOAT ? 4 : 1);
    CU(r->raw0.ensure(bytes));
    CU(cudaMemcpyAsync(r->raw0.p, data, bytes, cudaMemcpyHostToDevice, r->stream));
    r->have_next = false;
    if (next && dtype == VV_FLOAT) {
        CU(r->raw1.ensure(bytes));
        CU(cudaMemcpyAsync(r->raw1.p, next, bytes, cudaMemcpyHostToDevice, r->stream));
        r->have_next = true;
    }
    CU(cudaStreamSynchronize(r->stream));
    for (int i = 0; i < 3; ++i) { r->size[i] = dims[i]; r->slice_dist[i] = slice_dist ? slice_dist[i] : 1.0f; }
    volume_geometry(r);
    update_light(r);
    r->field_u8 = (dtype == VV_UCHAR);
    r->have_field = true;
    r->have_dat = false;                       // (vv_load_dat sets it again)
    r->interp_index = 0;
    r->field_dirty = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    // re-pack now so the cost is not hidden in the first frame
    return pack_field(r);
}

int vv_set_time_interp(VVRenderer *r, int interp_index, int interp_size)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (interp_size <= 0 || interp_index < 0) return fail(VV_ERR_INVALID, "bad interpolation index");
    CU(cudaSetDevice(r->device));
    r->interp_index = interp_index; r->interp_size = interp_size;
    r->field_dirty = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    return r->have_field ? pack_field(r) : VV_OK;
}

int vv_set_scalar(VVRenderer *r, const void *data, int dtype, const int dims[3])
{
    if (!r || !data || !dims) return fail(VV_ERR_INVALID, "vv_set_scalar: null argument");
    if (dtype != VV_UCHAR && dtype != VV_FLOAT)
        return fail(VV_ERR_INVALID, "VolumeData:  Only 8bit integer and 32bit float scalar is supported.");   // dataset.cpp:873-878
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || n > ((size_t)1 << 31)) return fail(VV_ERR_INVALID, "bad scalar dimensions");
    CU(cudaSetDevice(r->device));
    DevBuf<uint8_t> raw;
    CU(raw.ensure(n));
    if (dtype == VV_FLOAT) {
        // GL_LUMINANCE internal format with GL_FLOAT source data: clamped to [0,1] and stored as UNORM8
        DevBuf<float> f;
        CU(f.ensure(n));
        CU(cudaMemcpyAsync(f.p, data, n * sizeof(float), cudaMemcpyHostToDevice, r->stream));
        CU(launch_float_to_unorm8(f.p, n, raw.p, r->stream));
        CU(cudaStreamSynchronize(r->stream));
    } else {
        CU(cudaMemcpyAsync(raw.p, data, n, cudaMemcpyHostToDevice, r->stream));
    }
    CU(r->scalar_cell.ensure(n));
    CU(launch_build_cell8(raw.p, 1, 0, dims[0], dims[1], dims[2], 0, 0, r->scalar_cell.p, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    r->sdim[0] = dims[0]; r->sdim[1] = dims[1]; r->sdim[2] = dims[2];
    r->have_scalar = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    return VV_OK;
}

int vv_set_noise(VVRenderer *r, const uint8_t *data, const int dims[3], int with_gradients)
{
    if (!r || !data || !dims) return fail(VV_ERR_INVALID, "vv_set_noise: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(VV_ERR_INVALID, "bad noise dimensions");
    CU(cudaSetDevice(r->device));
    r->licvol_valid = false;
    return upload_noise(r, data, dims, with_gradients);
}

int vv_set_noise_with_gradients(VVRenderer *r, const uint8_t *noise, const uint8_t *gradients3, const int dims[3])
{
    if (!r || !noise || !gradients3 || !dims) return fail(VV_ERR_INVALID, "vv_set_noise_with_gradients: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(VV_ERR_INVALID, "bad noise dimensions");
    CU(cudaSetDevice(r->device));
    r->licvol_valid = false;
    return upload_noise(r, noise, dims, 1, gradients3);
}

int vv_read_noise_gradients(VVRenderer *r, uint8_t *out, size_t out_bytes)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_noise_gradients: null argument");
    if (!r->have_noise || !r->noise_has_grad) return fail(VV_ERR_STATE, "the current noise has no gradients");
    const size_t n = (size_t)r->ndim[0] * r->ndim[1] * r->ndim[2];
    if (out_bytes < 3 * n) return fail(VV_ERR_INVALID, "output buffer too small");
    CU(cudaSetDevice(r->device));
    std::vector<uint8_t> rgba(4 * n);
    CU(cudaMemcpyAsync(rgba.data(), r->noise_rgba.p, 4 * n, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    for (size_t i = 0; i < n; ++i) { out[3 * i] = rgba[4 * i]; out[3 * i + 1] = rgba[4 * i + 1]; out[3 * i + 2] = rgba[4 * i + 2]; }
    return VV_OK;
}

int vv_generate_white_noise(VVRenderer *r, int n, uint32_t seed, float p, int with_gradients)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (n <= 0 || n > 1024) return fail(VV_ERR_INVALID, "noise size out of range");
    // NoiseDataSet::loadData fallback, VV/dataset.cpp:1142-1163: floor(0.6 rand()/(RAND_MAX+1) + 0.5) * 255, i.e. 255 with
    // probability 1/6.  rand() is unseeded there; here the stream is mt19937(seed), u = draw / 2^32, 255 iff u < p.
    std::vector<uint8_t> v((size_t)n * n * n);
    std::mt19937 gen(seed);
    for (size_t i = 0; i < v.size(); ++i) v[i] = ((double)gen() / 4294967296.0 < (double)p) ? 255 : 0;
    const int dims[3] = {n, n, n};
    return vv_set_noise(r, v.data(), dims, with_gradients);
}

int vv_set_filter(VVRenderer *r, const uint8_t *row, int width, int channels)
{
    if (!r || !row || width <= 0 || channels <= 0) return fail(VV_ERR_INVALID, "vv_set_filter: bad argument");
    // LICFilter::loadData, VV/dataset.cpp:1448-1461; calcFilterKernelInvArea :1503-1512
    const int fw = next_pow2(width);
    const int shift = (fw - width) / 2;
    r->filter.assign(fw, 0);
    for (int i = 0; i < width; ++i) r->filter[i + shift] = row[i * channels];
    float area = 0.0f;
    for (int i = 0; i < fw; ++i) area += r->filter[i];
    r->inv_filter_area = 0.5f * fw * 255.0f / area;
    r->tables_dirty = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    return VV_OK;
}

int vv_set_box_filter(VVRenderer *r, int width)
{
    if (!r || width <= 0) return fail(VV_ERR_INVALID, "vv_set_box_filter: bad argument");
    r->filter.assign(next_pow2(width), 255);   // VV/dataset.cpp:1405-1413
    r->inv_filter_area = 0.5f;
    r->tables_dirty = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    return VV_OK;
}

int vv_set_tf(VVRenderer *r, const uint8_t *tf)
{
    if (!r || !tf) return fail(VV_ERR_INVALID, "vv_set_tf: null argument");
    std::memcpy(r->tf, tf, 256 * 5);
    r->tables_dirty = true;
    r->frame_valid = false;
    return VV_OK;
}

int vv_set_default_tf(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    for (int i = 0; i < 256; ++i) {
        for (int ch = 0; ch < 3; ++ch) r->tf[5 * i + ch] = (unsigned char)i;
        for (int ch = 3; ch < 5; ++ch) r->tf[5 * i + ch] = (unsigned char)((i < 20) ? 0 : i - 20);
    }
    r->tables_dirty = true;
    r->frame_valid = false;
    return VV_OK;
}

int vv_set_lic_params(VVRenderer *r, const VVLicParams *p)
{
    if (!r || !p) return fail(VV_ERR_INVALID, "vv_set_lic_params: null argument");
    if (p->stepsForward < 0 || p->stepsBackward < 0 || p->stepsForward > kMaxLicSteps || p->stepsBackward > kMaxLicSteps)
        return fail(VV_ERR_INVALID, "LIC steps out of range (0..1024 per direction)");
    if (!(p->stepSizeVol > 0.0f)) return fail(VV_ERR_INVALID, "stepSizeVol must be > 0");
    // the shader's two nested loops give numIterations^2 samples per ray at most (lic3d_fragment.glsl:38-40); bounded so that the
    // product fits 32 bits in every kernel and the sample buffer stays finite
    if (p->numIterations <= 0 || p->numIterations > 32768) return fail(VV_ERR_INVALID, "numIterations must be in 1..32768");
    const bool steps_changed = p->stepsForward != r->lp.stepsForward || p->stepsBackward != r->lp.stepsBackward;
    r->lp = *p;
    if (steps_changed) r->tables_dirty = true;
    r->frame_valid = false;
    return VV_OK;
}

int vv_set_camera(VVRenderer *r, const float quat[4], const float pos[3], float dist, float fovy, float near_clip, float far_clip)
{
    if (!r || !quat || !pos) return fail(VV_ERR_INVALID, "vv_set_camera: null argument");
    for (int i = 0; i < 4; ++i) r->cam_quat[i] = quat[i];
    for (int i = 0; i < 3; ++i) r->cam_pos[i] = pos[i];
    r->cam_dist = dist; r->fovy = fovy; r->near_clip = near_clip; r->far_clip = far_clip;
    r->frame_valid = false;
    return VV_OK;
}

int vv_set_light(VVRenderer *r, const float quat[4], float dist)
{
    if (!r || !quat) return fail(VV_ERR_INVALID, "vv_set_light: null argument");
    for (int i = 0; i < 4; ++i) r->light_quat[i] = quat[i];
    r->light_dist = dist;
    r->frame_valid = false;
    return VV_OK;
}

int vv_update_light_pos(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    update_light(r);
    r->frame_valid = false;
    return VV_OK;
}

int vv_set_window(VVRenderer *r, int window_width, int window_height)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (window_width < 0 || window_height < 0 || (window_width == 0) != (window_height == 0)) return fail(VV_ERR_INVALID, "vv_set_window: bad size");
    r->window_aspect = window_width > 0 ? (float)window_width / window_height : 0.0f;   // _aspect = (float)_w/_h
    r->frame_valid = false;
    return VV_OK;
}

int vv_enable_lowres(VVRenderer *r, int enable)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    r->lowres = enable != 0;
    r->tables_dirty = true;
    r->frame_valid = false;
    return VV_OK;
}

int vv_enable_float_target(VVRenderer *r, int enable)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (r->technique == VV_VOLIC_SLICING && r->float_target != (enable != 0)) r->frame_valid = false;   // another slicing program runs
    r->float_target = enable != 0;
    return VV_OK;
}

int vv_set_option(VVRenderer *r, int option, int value)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    switch (option) {
    case VV_OPT_TF_MODE:
        if (value < TF_B || value > TF_SCALAR) return fail(VV_ERR_INVALID, "bad tf mode");
        r->tf_modes[r->technique == VV_VOLIC_SLICING ? 1 : 0] = value; break;
    case VV_OPT_GATE_MODE:
        if (value != GATE_ALWAYS && value != GATE_TF_ALPHA) return fail(VV_ERR_INVALID, "bad gate mode");
        r->gate_modes[r->technique == VV_VOLIC_SLICING ? 1 : 0] = value; break;
    case VV_OPT_NOISE_GATE: r->noise_gate = value != 0; break;
    case VV_OPT_QUIRK_SCALEVOLINV: r->quirk_scalevolinv = value != 0; break;
    case VV_OPT_QUIRK_LUMINANCE_ALPHA: r->quirk_lum_alpha = value != 0; break;
    case VV_OPT_LICVOL_FP16: r->licvol_fp16 = value != 0; break;
    case VV_OPT_FIELD_LAYOUT:
        if (value != LAYOUT_F4 && value != LAYOUT_PAIR && value != LAYOUT_QUAD && value != 3) return fail(VV_ERR_INVALID, "bad field layout");
        r->field_layout_req = value;
        resolve_field_layout(r);
        break;
    case VV_OPT_COUNT_SAMPLES: r->count_samples = value != 0; break;
    case VV_OPT_LICVOL_SIZE:
        if (value < 0 || value > 2048) return fail(VV_ERR_INVALID, "bad LIC volume size");
        r->licvol_size = value; break;
    case VV_OPT_SPEC_EXP: r->spec_exp = (float)value; break;
    case VV_OPT_SAMPLE_MAP: r->sample_map = value != 0; break;
    case VV_OPT_RAYCAST_MODE:
        if (value != 0 && value != 1) return fail(VV_ERR_INVALID, "bad raycast mode");
        r->raycast_mode = value; break;
    case VV_OPT_NOISE_LAYOUT:
        if (value < 0 || value > 2) return fail(VV_ERR_INVALID, "bad noise layout");
        r->noise_layout = value; break;
    case VV_OPT_DEPTH_MAJOR: r->depth_major = value != 0; break;
    case VV_OPT_PARTITION_UNIT:
        if (value != 1 && value != 2 && value != 4 && value != 8) return fail(VV_ERR_INVALID, "partition unit must be 1, 2, 4 or 8 blocks");
        if (value != r->part_unit) {
            r->part_unit = value;
            r->frame_valid = false;
            if (r->width > 0) { int rc = ensure_frame(r); if (rc) return rc; }
        }
        break;
    case VV_OPT_BAND_ROWS:
        if (value < 1 || value > 1024) return fail(VV_ERR_INVALID, "bad band rows");
        r->band_rows = value; break;
    case VV_OPT_FIRST_WINDOW:
        if (value < 1 || value > 4096) return fail(VV_ERR_INVALID, "first window must be 1..4096 samples");
        r->first_window = value; break;
    case VV_OPT_WINDOW_GROWTH:
        if (value < 100 || value > 400) return fail(VV_ERR_INVALID, "window growth must be 100..400 percent");
        r->window_growth = value; break;
    case VV_OPT_WALK_FAST_PATHS:
        if (value != 0 && value != 1) return fail(VV_ERR_INVALID, "bad walk fast-path switch");
        r->xf_enable = value; break;
    case VV_OPT_LIC_CTAS_PER_SM:
        if (value < 0 || value > 8) return fail(VV_ERR_INVALID, "bad CTAs per SM");
        r->lic_ctas_per_sm = value; break;
    default: return fail(VV_ERR_INVALID, "unknown option");
    }
    r->frame_valid = false;
    r->licvol_valid = false;
    return VV_OK;
}

// Renderer::render(update), VV/renderer.cpp:126-312: update == 0 re-presents the stored frame (:150-152, 228)
static int render_frame(VVRenderer *r, int update);

// the tail of Renderer::renderFBO, VV/renderer.cpp:1478-1513: while recording (or after a screenshot request) every
// displayed frame is written as "<dir>/<frames>_<name>" (animation / recording) or "<dir>/<dd-mm-YYYY HH-MM-SS> <name>"
static int save_snapshot(VVRenderer *r)
{
    char stamp[64];
    std::string path = r->snapshot_dir.empty() ? std::string() : r->snapshot_dir + "/";
    if (r->animation_on || r->recording) {
        std::snprintf(stamp, sizeof(stamp), "%d_", r->frames);
    } else {
        std::time_t t = std::time(nullptr);
        std::tm tmv;
        localtime_r(&t, &tmv);
        std::strftime(stamp, sizeof(stamp), "%d-%m-%Y %H-%M-%S ", &tmv);
    }
    path += stamp + r->snapshot_name;
    int rc = vv_save_png(r, path.c_str(), 0);   // saveTexture(_imgBufferTex0, 4, 15, 255.0): the stored RGBA8 frame
    if (rc) return rc;
    r->last_snapshot = path;
    if (r->screenshot) r->screenshot = false;
    if (r->recording) ++r->frames;
    return VV_OK;
}

int vv_render(VVRenderer *r, int update)
{
    int rc = render_frame(r, update);
    if (rc == VV_OK && r->world == 1 && (r->screenshot || r->recording)) rc = save_snapshot(r);
    return rc;
}

int vv_set_snapshot(VVRenderer *r, const char *dir, const char *file_name, int animation_on)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (dir) r->snapshot_dir = dir;
    if (file_name) r->snapshot_name = file_name;
    r->animation_on = animation_on != 0;
    return VV_OK;
}

int vv_screenshot(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    r->screenshot = true;
    return VV_OK;
}

int vv_switch_recording(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    r->recording = !r->recording;
    return r->recording ? 1 : 0;
}

const char *vv_last_snapshot_path(VVRenderer *r) { return r ? r->last_snapshot.c_str() : ""; }

static int render_frame(VVRenderer *r, int update)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    CU(cudaSetDevice(r->device));
    if (!update && r->frame_valid) return VV_OK;
    r->launches = 0;
    if (r->technique == VV_VOLIC_LICVOLUME || r->technique == VV_VOLIC_VOLUMEANI) {
        if (!r->licvol_valid || r->technique == VV_VOLIC_VOLUMEANI) {
            int rc = compute_lic_volume(r);
            if (rc) return rc;
        }
    }
    DevParams P;
    int rc = fill_params(r, P, true, r->technique == VV_VOLIC_RAYCAST || r->technique == VV_VOLIC_SLICING);
    if (rc) return rc;
    // (the sample-parallel pipeline manages its counters itself: some of them outlive the frame, see render_sample_parallel)
    const bool pipeline = r->technique == VV_VOLIC_SLICING || (r->technique == VV_VOLIC_RAYCAST && r->raycast_mode != 0);
    if (!pipeline) {
        CU(cudaMemsetAsync(r->counters.p, 0, kNumCounters * sizeof(unsigned int), r->stream));
        r->geom_valid = false;             // the other techniques use the same counters and tile buffers
    }
    const int grid = persistent_grid(r, r->n_local_blocks);
    switch (r->technique) {
    case VV_VOLIC_SLICING:
        rc = render_sample_parallel(r, P);
        if (rc) return rc;
        break;
    case VV_VOLIC_RAYCAST:
        if (r->raycast_mode == 0) {
            CU(cudaEventRecord(r->ev0, r->stream));
            CU(launch_lic_raycast(P, r->field_layout, r->illum_mode, r->illum_mode != ILLUM_GRADIENT && r->noise_gate, r->speed_of_flow, grid, r->stream));
            CU(cudaEventRecord(r->ev1, r->stream));
            ++r->launches;
        } else {
            rc = render_sample_parallel(r, P);
            if (rc) return rc;
        }
        break;
    case VV_VOLIC_LICVOLUME:
    case VV_VOLIC_VOLUMEANI:
        CU(cudaEventRecord(r->ev0, r->stream));
        CU(launch_volume_raycast(P, r->field_layout, grid, r->stream));
        CU(cudaEventRecord(r->ev1, r->stream));
        ++r->launches;
        break;
    default:
        return fail(VV_ERR_INVALID, "technique not implemented");
    }
    if (r->world == 1) return run_unblock(r, r->tiles.p, 1, r->blocks_per_rank);
    r->frame_valid = false;   // multi-GPU: the caller gathers tile buffers and calls vv_assemble_tiles
    return VV_OK;
}

static int check_frame(VVRenderer *r, size_t out_bytes, size_t px_bytes)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (!r->frame_valid) return fail(VV_ERR_STATE, "no frame rendered");
    if (out_bytes < (size_t)r->width * r->height * px_bytes) return fail(VV_ERR_INVALID, "output buffer too small");
    return VV_OK;
}

int vv_read_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes)
{
    int rc = check_frame(r, out_bytes, 4);
    if (rc) return rc;
    CU(cudaSetDevice(r->device));
    CU(cudaMemcpyAsync(out, r->frame8.p, (size_t)r->width * r->height * 4, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return VV_OK;
}

int vv_read_rgba32f(VVRenderer *r, float *out, size_t out_bytes)
{
    int rc = check_frame(r, out_bytes, 16);
    if (rc) return rc;
    CU(cudaSetDevice(r->device));
    CU(cudaMemcpyAsync(out, r->frame.p, (size_t)r->width * r->height * 16, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    if (r->float_target) {
        // enableFBO(true): the frame lives in a GL_RGBA16F_ARB texture (VV/renderer.cpp:562-606): what is read back is the
        // shader's result rounded to fp16
        const size_t n = (size_t)r->width * r->height * 4;
        for (size_t i = 0; i < n; ++i) out[i] = __half2float(__float2half_rn(out[i]));
    }
    return VV_OK;
}

int vv_read_display_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes)
{
    int rc = check_frame(r, out_bytes, 4);
    if (rc) return rc;
    CU(cudaSetDevice(r->device));
    CU(cudaMemcpyAsync(out, r->display8.p, (size_t)r->width * r->height * 4, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return VV_OK;
}

// The display pass over a WINDOW larger than the stored frame (low-res preset): background_fragment.glsl reads
// texture2DRect(imageFBOSampler, gl_FragCoord.xy * viewport.xy) with viewport = (renderWidth / winWidth, renderHeight / winHeight)
// as floats (VV/renderer.cpp:1438-1441) -- NEAREST, edge-clamped; the blend over white is per texel, so the displayed frame is
// up-scaled as it is read back.
int vv_read_display_window_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes, int window_width, int window_height)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_display_window_rgba8: null argument");
    if (window_width <= 0 || window_height <= 0) return fail(VV_ERR_INVALID, "vv_read_display_window_rgba8: bad window size");
    if (out_bytes < (size_t)window_width * window_height * 4) return fail(VV_ERR_INVALID, "output buffer too small");
    std::vector<uint8_t> img((size_t)r->width * r->height * 4);
    int rc = vv_read_display_rgba8(r, img.data(), img.size());
    if (rc) return rc;
    const float vx = static_cast<float>(r->width) / window_width, vy = static_cast<float>(r->height) / window_height;
    for (int y = 0; y < window_height; ++y) {
        int sy = (int)std::floor(((float)y + 0.5f) * vy);
        sy = sy < 0 ? 0 : (sy > r->height - 1 ? r->height - 1 : sy);
        for (int x = 0; x < window_width; ++x) {
            int sx = (int)std::floor(((float)x + 0.5f) * vx);
            sx = sx < 0 ? 0 : (sx > r->width - 1 ? r->width - 1 : sx);
            std::memcpy(out + 4 * ((size_t)y * window_width + x), &img[4 * ((size_t)sy * r->width + sx)], 4);
        }
    }
    return VV_OK;
}

int vv_read_lic_volume(VVRenderer *r, float *out, size_t out_bytes, int dims_out[3])
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_lic_volume: null argument");
    if (!r->licvol_valid) return fail(VV_ERR_STATE, "LIC volume not computed (vv_update_lic_volume)");
    const size_t n = (size_t)r->ldim[0] * r->ldim[1] * r->ldim[2];
    if (out_bytes < n * sizeof(float)) return fail(VV_ERR_INVALID, "output buffer too small");
    CU(cudaSetDevice(r->device));
    CU(cudaMemcpyAsync(out, r->licvol.p, n * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    if (dims_out) { dims_out[0] = r->ldim[0]; dims_out[1] = r->ldim[1]; dims_out[2] = r->ldim[2]; }
    return VV_OK;
}

static float half_bits_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            exp = 127 - 15 + 1;
            while (!(man & 0x400u)) { man <<= 1; --exp; }
            bits = sign | (exp << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

int vv_read_field_texture(VVRenderer *r, float *out, size_t out_bytes)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_field_texture: null argument");
    if (!r->have_field) return fail(VV_ERR_STATE, "no vector field set");
    CU(cudaSetDevice(r->device));
    if (r->field_dirty) { int rc = pack_field(r); if (rc) return rc; }
    const size_t n = (size_t)r->size[0] * r->size[1] * r->size[2];
    if (out_bytes < n * 16) return fail(VV_ERR_INVALID, "output buffer too small");
    if (r->field_layout == LAYOUT_F4) {
        CU(cudaMemcpyAsync(out, r->field_f4.p, n * 16, cudaMemcpyDeviceToHost, r->stream));
        CU(cudaStreamSynchronize(r->stream));
        return VV_OK;
    }
    // gather the unpadded [nz][ny][nx] texels out of the padded x-pair layout, one row at a time
    const size_t nx = r->size[0], ny = r->size[1], nz = r->size[2], G = r->field_guard, row = r->field_row, py = ny + 2 * G;
    const size_t pc = r->field_layout == LAYOUT_QUAD ? 2 : 1;             // uint4 per cell; the cell's own texel comes first in both
    std::vector<uint16_t> tmp(nx * ny * 8 * pc);
    for (size_t z = 0; z < nz; ++z) {
        CU(cudaMemcpy2DAsync(tmp.data(), nx * 16 * pc, r->field_pair.p + (((z + G) * py + G) * row + r->field_gx) * pc, row * 16 * pc, nx * 16 * pc, ny,
                             cudaMemcpyDeviceToHost, r->stream));
        CU(cudaStreamSynchronize(r->stream));
        for (size_t y = 0; y < ny; ++y)
            for (size_t x = 0; x < nx; ++x) {
                const size_t i = (z * ny + y) * nx + x, j = (y * nx + x) * pc;
                for (int k = 0; k < 4; ++k) out[4 * i + k] = half_bits_to_float(tmp[8 * j + k]);
            }
    }
    return VV_OK;
}

int vv_read_noise_texture(VVRenderer *r, uint8_t *out, size_t out_bytes, int *channels)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_noise_texture: null argument");
    if (!r->have_noise) return fail(VV_ERR_STATE, "no noise set");
    CU(cudaSetDevice(r->device));
    const size_t n = (size_t)r->ndim[0] * r->ndim[1] * r->ndim[2];
    const int ch = r->noise_has_grad ? 4 : 1;
    if (out_bytes < n * ch) return fail(VV_ERR_INVALID, "output buffer too small");
    CU(cudaMemcpyAsync(out, r->noise_has_grad ? (const void *)r->noise_rgba.p : (const void *)r->noise_raw.p, n * ch, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    if (channels) *channels = ch;
    return VV_OK;
}

int vv_read_sample_map(VVRenderer *r, uint32_t *out, size_t out_bytes)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_read_sample_map: null argument");
    if (!r->sample_map || !r->sample_tiles.p) return fail(VV_ERR_STATE, "VV_OPT_SAMPLE_MAP was not enabled before vv_render");
    if (r->world != 1) return fail(VV_ERR_STATE, "sample map is only available on an unpartitioned handle");
    const size_t npx = (size_t)r->width * r->height;
    if (out_bytes < npx * 4) return fail(VV_ERR_INVALID, "output buffer too small");
    CU(cudaSetDevice(r->device));
    std::vector<uint32_t> t((size_t)r->blocks_per_rank * kBlockPixels);
    CU(cudaMemcpyAsync(t.data(), r->sample_tiles.p, t.size() * 4, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    for (int y = 0; y < r->height; ++y)
        for (int x = 0; x < r->width; ++x) {
            const int b = (y / kBlockDim) * r->nbx + x / kBlockDim;
            out[(size_t)y * r->width + x] = t[(size_t)b * kBlockPixels + (y % kBlockDim) * kBlockDim + (x % kBlockDim)];
        }
    return VV_OK;
}

// Renderer::saveTexture(_imgBufferTex0, 4, 15, 255) / saveFrameBuffer, VV/renderer.cpp:340-428, 1500: PNG rows top-down
int vv_save_png(VVRenderer *r, const char *path, int displayed)
{
    if (!r || !path) return fail(VV_ERR_INVALID, "vv_save_png: null argument");
    std::vector<uint8_t> img((size_t)r->width * r->height * 4), flip(img.size());
    int rc;
    if (!displayed && r->float_target) {
        // saveTexture on a floating-point texture (VV/renderer.cpp:386-403): v = (int)(scale * texel) -- truncation, not the
        // GL's round-to-nearest -- clamped to 0..255, all four channels (mask 15, scale 255)
        std::vector<float> f((size_t)r->width * r->height * 4);
        rc = vv_read_rgba32f(r, f.data(), f.size() * sizeof(float));      // fp16-rounded, see there
        if (rc) return rc;
        for (size_t i = 0; i < f.size(); ++i) {
            int v = (int)(255.0f * f[i]);
            img[i] = (uint8_t)(v > 255 ? 255 : (v < 0 ? 0 : v));
        }
    } else {
        rc = displayed ? vv_read_display_rgba8(r, img.data(), img.size()) : vv_read_rgba8(r, img.data(), img.size());
        if (rc) return rc;
    }
    const size_t stride = (size_t)r->width * 4;
    for (int y = 0; y < r->height; ++y) std::memcpy(&flip[stride * y], &img[stride * (r->height - 1 - y)], stride);
    std::string err;
    if (!png_write_file(path, flip.data(), r->width, r->height, 4, err)) return fail(VV_ERR_IO, err);
    return VV_OK;
}

int vv_save_raw(VVRenderer *r, const char *path)
{
    if (!r || !path) return fail(VV_ERR_INVALID, "vv_save_raw: null argument");
    std::vector<float> img((size_t)r->width * r->height * 4);
    int rc = vv_read_rgba32f(r, img.data(), img.size() * sizeof(float));
    if (rc) return rc;
    FILE *fp = std::fopen(path, "wb");
    if (!fp) return fail(VV_ERR_IO, std::string("cannot write ") + path);
    const int32_t hdr[2] = {r->width, r->height};
    std::fwrite(hdr, sizeof(hdr), 1, fp);
    std::fwrite(img.data(), sizeof(float), img.size(), fp);
    std::fclose(fp);
    return VV_OK;
}

uint64_t vv_last_ray_samples(VVRenderer *r)
{
    if (!r || !r->counters.p) return 0;
    cudaSetDevice(r->device);
    unsigned long long v = 0;
    cudaMemcpyAsync(&v, r->counters.p, sizeof(v), cudaMemcpyDeviceToHost, r->stream);
    cudaStreamSynchronize(r->stream);
    return (uint64_t)v;
}

float vv_last_kernel_ms(VVRenderer *r)
{
    if (!r) return -1.0f;
    cudaSetDevice(r->device);
    float ms = -1.0f;
    if (cudaEventSynchronize(r->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, r->ev0, r->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int vv_last_launch_count(VVRenderer *r) { return r ? r->launches : 0; }
int vv_field_layout(VVRenderer *r) { if (!r) return -1; resolve_field_layout(r); return r->field_layout; }

int vv_synchronize(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    CU(cudaSetDevice(r->device));
    CU(cudaStreamSynchronize(r->stream));
    return VV_OK;
}

int vv_set_partition(VVRenderer *r, int rank, int world)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (world < 1 || rank < 0 || rank >= world) return fail(VV_ERR_INVALID, "bad partition");
    CU(cudaSetDevice(r->device));
    r->rank = rank; r->world = world;
    r->frame_valid = false;
    return (r->width > 0) ? ensure_frame(r) : VV_OK;
}

int vv_set_licvol_slab(VVRenderer *r, int z0, int z1)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    r->slab_z0 = z0; r->slab_z1 = z1;
    r->licvol_valid = false;
    return VV_OK;
}

int vv_get_tile_buffer(VVRenderer *r, void **dev_ptr, int *n_local_blocks, int *n_total_blocks)
{
    if (!r || !dev_ptr) return fail(VV_ERR_INVALID, "vv_get_tile_buffer: null argument");
    *dev_ptr = r->tiles.p;
    if (n_local_blocks) *n_local_blocks = r->blocks_per_rank;
    if (n_total_blocks) *n_total_blocks = r->nbx * r->nby;
    return VV_OK;
}

int vv_assemble_tiles(VVRenderer *r, const void *gathered_dev, int world)
{
    if (!r || !gathered_dev) return fail(VV_ERR_INVALID, "vv_assemble_tiles: null argument");
    if (world != r->world) return fail(VV_ERR_INVALID, "vv_assemble_tiles: world mismatch");
    CU(cudaSetDevice(r->device));
    return run_unblock(r, (const float4 *)gathered_dev, world, r->blocks_per_rank);
}

// ---- peer-to-peer frame exchange ---------------------------------------------------------------------------------
static void p2p_close(VVRenderer *r)
{
    for (int j = 0; j < kMaxPeers; ++j) {
        if (r->p2p_opened[j] && r->p2p_peer[j]) cudaIpcCloseMemHandle(r->p2p_peer[j]);
        r->p2p_peer[j] = nullptr;
        r->p2p_opened[j] = false;
    }
}

int vv_p2p_export(VVRenderer *r, void *ipc_handle_out, void **base_out)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (r->width <= 0 || r->height <= 0) return fail(VV_ERR_STATE, "vv_p2p_export: vv_resize / vv_set_partition first");
    if (r->world > kMaxPeers) return fail(VV_ERR_INVALID, "vv_p2p_export: more than 16 ranks");
    CU(cudaSetDevice(r->device));
    CU(cudaStreamSynchronize(r->stream));
    p2p_close(r);
    if (r->p2p_base) { CU(cudaFree(r->p2p_base)); r->p2p_base = nullptr; }
    r->p2p_tile_bytes = (size_t)r->world * r->blocks_per_rank * kBlockPixels * sizeof(float4);
    r->p2p_world = r->world;
    r->p2p_epoch = 0;
    CU(cudaMalloc((void **)&r->p2p_base, 256 + 2 * r->p2p_tile_bytes));   // plain cudaMalloc: IPC-exportable
    CU(cudaMemset(r->p2p_base, 0, 256 + 2 * r->p2p_tile_bytes));
    CU(r->p2p_scratch.ensure(2));
    CU(cudaMemset(r->p2p_scratch.p, 0, 2 * sizeof(unsigned int)));
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, r->p2p_base));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(ipc_handle_out, &h, sizeof(h));
    }
    if (base_out) *base_out = r->p2p_base;
    return VV_OK;
}

int vv_p2p_connect(VVRenderer *r, const void *ipc_handles, void *const *local_bases, int world)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (!r->p2p_base || world != r->p2p_world || world != r->world) return fail(VV_ERR_STATE, "vv_p2p_connect: vv_p2p_export first (same partition)");
    if (!ipc_handles && !local_bases) return fail(VV_ERR_INVALID, "vv_p2p_connect: no peer handles");
    CU(cudaSetDevice(r->device));
    p2p_close(r);
    for (int j = 0; j < world; ++j) {
        if (j == r->rank) { r->p2p_peer[j] = r->p2p_base; continue; }
        if (local_bases && local_bases[j]) { r->p2p_peer[j] = local_bases[j]; continue; }   // peers of the same process
        if (!ipc_handles) return fail(VV_ERR_INVALID, "vv_p2p_connect: missing handle");
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const unsigned char *)ipc_handles + 64 * (size_t)j, sizeof(h));
        void *ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        r->p2p_peer[j] = ptr;
        r->p2p_opened[j] = true;
    }
    return VV_OK;
}

int vv_p2p_render(VVRenderer *r)
{
    int rc = render_frame(r, 1);
    if (rc) return rc;
    if (r->world == 1) return VV_OK;
    if (!r->p2p_base || !r->p2p_peer[r->world - 1] || !r->p2p_peer[0]) return fail(VV_ERR_STATE, "vv_p2p_render: not connected (vv_p2p_export / vv_p2p_connect)");
    // the gather buffers (here and in every peer) were sized by vv_p2p_export for the frame size and partition of that moment: a
    // vv_resize / vv_set_partition since then would make the stores below land outside the peers' buffers
    if (r->world != r->p2p_world || (size_t)r->world * r->blocks_per_rank * kBlockPixels * sizeof(float4) != r->p2p_tile_bytes)
        return fail(VV_ERR_STATE, "vv_p2p_render: frame size or partition changed since vv_p2p_export (export and connect again on every rank)");
    const unsigned int parity = r->p2p_epoch & 1u;
    P2PArgs a;
    std::memset(&a, 0, sizeof(a));
    a.tiles = r->tiles.p;
    a.rank = r->rank; a.world = r->world;
    a.n = (size_t)r->blocks_per_rank * kBlockPixels;
    a.doneCounter = r->p2p_scratch.p;
    for (int j = 0; j < r->world; ++j) {
        unsigned char *b = (unsigned char *)r->p2p_peer[j];
        a.peerFlags[j] = (unsigned int *)(b + 128 * parity);
        a.peerTiles[j] = (float4 *)(b + 256 + parity * r->p2p_tile_bytes);
    }
    // grid.x CTAs per destination rank (grid.y = world): about 4 CTAs per SM in total
    const int grid = (int)std::min<size_t>((a.n + 1023) / 1024, (size_t)std::max(1, r->num_sms * 4 / r->world));
    CU(launch_scatter_tiles(a, std::max(grid, 1), r->stream));
    const unsigned int target = (unsigned int)r->world * (r->p2p_epoch / 2 + 1);
    CU(launch_wait_arrivals((unsigned int *)(r->p2p_base + 128 * parity), target, r->p2p_scratch.p + 1, r->stream));
    r->launches += 2;
    ++r->p2p_epoch;
    return run_unblock(r, (const float4 *)(r->p2p_base + 256 + parity * r->p2p_tile_bytes), r->world, r->blocks_per_rank);
}

// VV_ERR_STATE if a wait timed out (a peer never delivered its tiles); synchronises the stream
int vv_p2p_status(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (!r->p2p_scratch.p) return VV_OK;
    CU(cudaSetDevice(r->device));
    unsigned int e[2] = {0, 0};
    CU(cudaMemcpyAsync(e, r->p2p_scratch.p, sizeof(e), cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return e[1] ? fail(VV_ERR_STATE, "peer-to-peer exchange timed out: a rank did not deliver its tiles") : VV_OK;
}

int vv_p2p_disconnect(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    CU(cudaSetDevice(r->device));
    CU(cudaStreamSynchronize(r->stream));
    p2p_close(r);
    if (r->p2p_base) { CU(cudaFree(r->p2p_base)); r->p2p_base = nullptr; }
    r->p2p_world = 0;
    return VV_OK;
}

int vv_debug_walk(VVRenderer *r, const float pos[3], int dir_sign, int n_steps, int walk_variant, float *out, size_t out_bytes)
{
    if (!r || !pos || !out) return fail(VV_ERR_INVALID, "vv_debug_walk: null argument");
    if (n_steps < 1 || out_bytes < (size_t)n_steps * 16 * sizeof(float)) return fail(VV_ERR_INVALID, "vv_debug_walk: output buffer too small");
    CU(cudaSetDevice(r->device));
    DevParams P;
    int rc = fill_params(r, P, false, true);
    if (rc) return rc;
    if (r->field_layout == LAYOUT_F4) return fail(VV_ERR_STATE, "vv_debug_walk: fp16 cell layouts only (x-pair / xy-quad)");
    const bool grad = (r->illum_mode == ILLUM_GRADIENT);
    if (n_steps > (dir_sign < 0 ? P.nBwd : P.nFwd)) return fail(VV_ERR_INVALID, "vv_debug_walk: more steps than the LIC parameters have");
    if (walk_variant != 0 && !P.guardOk) return fail(VV_ERR_STATE, "vv_debug_walk: no guard band for this field / these LIC parameters");
    if (walk_variant == 3 && !(grad && P.noiseShared)) return fail(VV_ERR_STATE, "vv_debug_walk: the noise does not share the field's cells");
    if (walk_variant != 0 && walk_variant != 1 && walk_variant != 3) return fail(VV_ERR_INVALID, "vv_debug_walk: bad variant");
    if (grad && !P.noise_bf) return fail(VV_ERR_STATE, "vv_debug_walk: bf16 noise layout only");
    DevBuf<float> d;
    CU(d.ensure((size_t)n_steps * 16));
    CU(cudaMemsetAsync(d.p, 0, (size_t)n_steps * 16 * sizeof(float), r->stream));
    CU(launch_debug_walk(P, r->field_layout, grad, walk_variant, pos, dir_sign, n_steps, d.p, r->stream));
    CU(cudaMemcpyAsync(out, d.p, (size_t)n_steps * 16 * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return VV_OK;
}

int vv_get_lic_volume_ptr(VVRenderer *r, void **dev_ptr, int dims_out[3])
{
    if (!r || !dev_ptr) return fail(VV_ERR_INVALID, "vv_get_lic_volume_ptr: null argument");
    if (!r->licvol.p) return fail(VV_ERR_STATE, "LIC volume not allocated");
    *dev_ptr = r->licvol.p;
    if (dims_out) { dims_out[0] = r->ldim[0]; dims_out[1] = r->ldim[1]; dims_out[2] = r->ldim[2]; }
    r->licvol_valid = true;   // the caller may have filled the other slabs (all-gather)
    return VV_OK;
}

int vv_set_stream(VVRenderer *r, void *cuda_stream)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    r->stream = cuda_stream ? (cudaStream_t)cuda_stream : r->own_stream;
    return VV_OK;
}

// ---- loaders ------------------------------------------------------------------------------------
int vv_load_dat(VVRenderer *r, const char *dat_path)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    VVDatInfo info;
    int rc = parse_dat(dat_path, &info);
    if (rc) return rc;
    if (info.data_dim != 3) return fail(VV_ERR_IO, std::string("VectorData:  DAT file \"") + dat_path + "\" refers not to a vector data set.");
    if (info.data_type != VV_UCHAR && info.data_type != VV_FLOAT)
        return fail(VV_ERR_IO, "VectorData:  Only 8bit integer and 32bit float vectors are supported.");
    const size_t bytes = dat_bytes(&info);
    std::vector<uint8_t> a(bytes), b;
    rc = read_raw(&info, info.time_begin, a.data(), bytes);
    if (rc) return rc;
    const void *next = nullptr;
    if (info.time_end > info.time_begin && info.data_type == VV_FLOAT) {
        // VV/3DLIC.cpp:701-702: data = timestep cur, newData = the following one
        b.resize(bytes);
        rc = read_raw(&info, info.time_begin + 1, b.data(), bytes);
        if (rc) return rc;
        next = b.data();
    }
    rc = vv_set_vector_field(r, a.data(), next, info.data_type, info.resolution, info.slice_thickness);
    if (rc) return rc;
    r->dat = info;
    r->have_dat = true;
    vv_time_cursor_init(&r->cursor, info.time_begin, info.time_end, r->interp_size);
    // init() has run createTextureIterp + checkInterpolateStage once by the time the first idle tick comes (VV/3DLIC.cpp:706-707):
    // the texture just packed used fraction 0, the first vv_idle uses 1 / InterpSize
    vv_time_cursor_tick(&r->cursor, nullptr);
    return VV_OK;
}

// One idle() tick of the reference's animation (VV/3DLIC.cpp:129-172): createTextureIterp + checkInterpolateStage
int vv_idle(VVRenderer *r)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    if (!r->have_dat || !r->have_field) return fail(VV_ERR_STATE, "vv_idle: the vector field was not loaded with vv_load_dat");
    CU(cudaSetDevice(r->device));
    r->cursor.interp_size = r->interp_size;
    int advanced = 0;
    r->interp_index = vv_time_cursor_tick(&r->cursor, &advanced);
    r->field_dirty = true;
    r->frame_valid = false;
    r->licvol_valid = false;
    int rc = pack_field(r);                    // this tick's texture: data + interpIndex / InterpSize (newData - data)
    if (rc) return rc;
    if (advanced && r->have_next) {
        // data <- time step `current`, newData <- the one after it; the texture changes with the next tick
        const size_t bytes = dat_bytes(&r->dat);
        std::vector<uint8_t> a(bytes), b(bytes);
        rc = read_raw(&r->dat, r->cursor.current, a.data(), bytes);
        if (rc) return rc;
        rc = read_raw(&r->dat, vv_time_cursor_next(&r->cursor), b.data(), bytes);
        if (rc) return rc;
        CU(cudaMemcpyAsync(r->raw0.p, a.data(), bytes, cudaMemcpyHostToDevice, r->stream));
        CU(cudaMemcpyAsync(r->raw1.p, b.data(), bytes, cudaMemcpyHostToDevice, r->stream));
        CU(cudaStreamSynchronize(r->stream));
    }
    return VV_OK;
}

int vv_get_time_cursor(VVRenderer *r, VVTimeCursor *out)
{
    if (!r || !out) return fail(VV_ERR_INVALID, "vv_get_time_cursor: null argument");
    if (!r->have_dat) return fail(VV_ERR_STATE, "the vector field was not loaded with vv_load_dat");
    *out = r->cursor;
    return VV_OK;
}

int vv_load_scalar_dat(VVRenderer *r, const char *dat_path)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    VVDatInfo info;
    int rc = parse_dat(dat_path, &info);
    if (rc) return rc;
    if (info.data_dim != 1) return fail(VV_ERR_IO, std::string("VolumeData:  DAT file \"") + dat_path + "\" refers not to a scalar data set.");
    if (info.data_type != VV_UCHAR && info.data_type != VV_FLOAT)
        return fail(VV_ERR_IO, "VolumeData:  Only 8bit integer and 32bit float scalar is supported.");
    const size_t bytes = dat_bytes(&info);
    std::vector<uint8_t> a(bytes);
    rc = read_raw(&info, info.time_begin, a.data(), bytes);
    if (rc) return rc;
    return vv_set_scalar(r, a.data(), info.data_type, info.resolution);
}

int vv_load_noise(VVRenderer *r, const char *path, int with_gradients)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    std::vector<uint8_t> data;
    int dims[3];
    int rc = read_noise_file(path, data, dims);
    if (rc) return rc;
    if (!with_gradients) return vv_set_noise(r, data.data(), dims, 0);
    // NoiseDataSet::createTexture, VV/dataset.cpp:1238-1267: stored gradients if there are any, else compute and store
    std::vector<uint8_t> grd((size_t)3 * dims[0] * dims[1] * dims[2]);
    if (vv_grd_read(path, dims, grd.data()) == VV_OK) return vv_set_noise_with_gradients(r, data.data(), grd.data(), dims);
    set_error("");
    rc = vv_set_noise(r, data.data(), dims, 1);
    if (rc) return rc;
    rc = vv_read_noise_gradients(r, grd.data(), grd.size());
    if (rc) return rc;
    if (vv_grd_write(path, dims, grd.data()) != VV_OK) set_error("");   // "Saving gradients was not sucessful" is only a warning there
    return VV_OK;
}

int vv_load_filter_png(VVRenderer *r, const char *path)
{
    if (!r || !path) return fail(VV_ERR_INVALID, "FilterKernel:  No filename set.");
    std::vector<uint8_t> img;
    int w, h, ch;
    std::string err;
    if (!png_read_file(path, img, w, h, ch, err)) return fail(VV_ERR_IO, std::string("FilterKernel:  Could not load filter kernel (\"") + path + "\"): " + err);
    return vv_set_filter(r, img.data(), w, ch);   // first row, first channel (VV/dataset.cpp:1439-1461)
}

int vv_load_tf_png(VVRenderer *r, const char *name)
{
    if (!r) return fail(VV_ERR_INVALID, "null renderer");
    int rc = load_tf_png(name, r->tf);
    if (rc) return rc;
    r->tables_dirty = true;
    r->frame_valid = false;
    return VV_OK;
}

} // extern "C"
