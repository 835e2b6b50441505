/*
 * vv_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU (fp32, OpenMP) restatement of the reference's 3D-LIC hot path.  It is
 * the checker that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg use; the product library (vectorvisualization_b200/csrc) never includes,
 * links or calls anything in this directory.
 *
 * PARITY PIN: the reference (liaoyg/VectorVisualization) ships no tests, no
 * golden vectors and no input data.  This restatement
 * is pinned two ways instead (see oracle/README.md):
 *   1. oracle/_ref/libvv_ref.so -- the reference's OWN sources compiled where
 *      they lie: the C++ loaders/pre-processing (reader, parseArg, gradient,
 *      dataset, mmath, transferEdit) against a capturing GL shim, and the GLSL
 *      shaders (inc_lic / inc_illum / lic3d_* / raycast_lic3d_*) compiled as
 *      C++ through a GLSL-syntax shim.  tests/test_oracle_vs_ref.py compares.
 *   2. closed-form known-answer tests (tests/test_oracle_closed_form.py).
 * GL texture filtering itself (GL 2.1 spec 3.8) is third-party arithmetic the
 * reference does not pin; both sides use the rules in SURVEY.md B.6.
 *
 * Every function cites the reference file:line it follows (VV/ =
 * /root/reference/VectorVisualization/).
 */
#ifndef VV_ORACLE_H_
#define VV_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VVO_ILLUM_NONE = 0, VVO_ILLUM_GRADIENT = 1, VVO_ILLUM_MALLO = 2, VVO_ILLUM_ZOECKLER = 3 };
/* TF index source: lic3d_fragment.glsl:53-56 and siblings */
enum { VVO_TF_B = 0, VVO_TF_A = 1, VVO_TF_R = 2, VVO_TF_LENGTH = 3, VVO_TF_SCALAR = 4 };
/* LIC gate: lic3d_fragment.glsl:59-61, lic3d_slicing_fragment.glsl:46 */
enum { VVO_GATE_ALWAYS = 0, VVO_GATE_TF_ALPHA = 1 };
enum { VVO_TECH_RAYCAST = 1, VVO_TECH_SLICING = 2, VVO_TECH_LICVOLUME = 3 };

typedef struct VVOScene {
    /* textures (non-owning) */
    const float   *vec;     int vdim[3];   /* RGBA float [z][y][x][4], already fp16-rounded (Q11) */
    const uint8_t *scalar;  int sdim[3];   /* LUMINANCE8 [z][y][x] */
    const uint8_t *noise;   int ndim[3];   int noise_channels; /* 1: LUMINANCE, 4: RGBA (grad.xyz, noise) */
    const uint8_t *kernel;  int kwidth;    float inv_filter_area;
    const uint8_t *tf;                     /* 256 x 5: R,G,B,alpha,LIC-opacity (transferEdit.cpp:76-82) */
    const float   *licvol;  int ldim[3];   /* scalar LIC volume [z][y][x], sampled REPEAT (Q14) */
    const float   *illum_tex[3];           /* zoeckler (LA -> 2ch), mallo diffuse (1ch), mallo specular (1ch), 2D [h][w][c] */
    int            illum_dim[2];
    /* volume geometry (dataset.cpp:144-176) */
    float extent[3], scale[3], scale_inv[3], center[3];
    /* LICParams (types.h:91-109) */
    float step_size_vol, gradient_scale, illum_scale, freq_scale;
    int   num_iterations, steps_fwd, steps_bwd;
    float step_size_lic;
    /* camera (camera.cpp:42-68) and light (renderer.cpp:431-466) */
    float cam_quat[4];  /* x,y,z,w */
    float cam_pos[3];
    float cam_dist, fovy;
    float light_quat[4];
    float light_dist;
    float spec_exp;     /* gl_LightSource[0].spotExponent, 3DLIC.cpp:736 */
    int   width, height;
    /* shader specialisation / quirks */
    int illum_mode, tf_mode, gate_mode;
    int noise_gate;            /* 1: scalar band (0.1,0.3) gates the noise (inc_lic.glsl:76-89) */
    int lowres;                /* renderer.cpp:947-966 */
    int quirk_scalevolinv;     /* Q1 */
    int quirk_luminance_alpha; /* Q7 */
    int speed_of_flow;         /* inc_lic.glsl:108-110,120-122 */
    int licvol_fp16;           /* Q14: LIC volume stored as RGBA16F */
    int weight_bits;           /* 0: exact fp32 lerp weights, filter formula as GL 2.1 3.8.8 writes it; 8: quantise f to 8 fractional bits
                                * (B.6); -1: exact weights, 3-D filtering with fused multiply-adds (u = fma(s,n,-.5), a + f (b - a) as sub + fma) */
    /* SURVEY 8(f) N4 */
    const float *mc_offsets;   /* USE_MC_OFFSET (lic3d_fragment.glsl:31-33, renderer.cpp:636-679): [height][width] ray-start
                                  offsets in [0,1], already fp16-rounded (GL_LUMINANCE16F rectangle texture); NULL = off */
    int    num_clip_planes;    /* active user clip planes (transform.cpp:424-449, renderer.cpp:156-163,1294-1309) */
    double clip_planes[3][4];  /* ClipPlane::setNormal arguments (n.xyz, d) in volume-centred object space.  The reference normalises
                                  n IN PLACE the first time the plane's cap is drawn (drawClippedPolygon hands getNormal() to
                                  ViewSlicing::setupSingleSlice, renderer.cpp:1301, slicing.cpp:337-348), d untouched: from then
                                  on the GL plane is n^.q + d >= 0.  The oracle evaluates that steady state. */
    float  near_clip, far_clip; /* gluPerspective near / far (camera.cpp:42-47: 0.1, 50): GL clips the proxy geometry to the view
                                  volume, so a fragment exists only where its eye-space depth lies in [near, far] */
    float  window_aspect;      /* Camera::setWindow (transform.h:79-80): aspect of gluPerspective when the frame is not the window
                                  (low-res preset: half-size viewport, renderer.cpp:111-119); 0 = width / height */
    int    fbo_fp16;           /* slicing: the ping-pong targets of the FBO path are GL_RGBA16F_ARB (renderer.cpp:566-606), so what a
                                  slice reads back is the previous slices' result rounded to fp16.  1: round after every slice (what
                                  the reference's targets do); 0: keep fp32 (the idealised model, used to bound the effect) */
    int    fbo_pingpong;       /* slicing: Renderer::sliceVolume swaps _imgBufferTex0 / _imgBufferTex1 before EVERY slice and a slice
                                  writes only the pixels its polygon covers (renderer.cpp:1201-1225): slice i reads the target of
                                  slice i-1 and writes the other texture, whose uncovered pixels keep what slice i-2 left there.
                                  The display pass and saveTexture read the target of the LAST slice (N-1).  1: two buffers per
                                  pixel, exactly that; 0: one accumulator carried from fragment to fragment (idealised) */
} VVOScene;

/* ---- hot path ------------------------------------------------------- */
/* lic3d_fragment.glsl:5-99 ; out_rgba [h][w][4] premultiplied, GL row order (row 0 = bottom);
 * out_samples [h][w] ray-sample counts (may be NULL).  Returns total ray samples. */
uint64_t vvo_raycast_lic(const VVOScene *s, float *out_rgba, uint32_t *out_samples);
/* same, restricted to pixels x0<=x<x1, y0<=y<y1 (others untouched) -- bounded CPU-baseline samples */
uint64_t vvo_raycast_lic_rect(const VVOScene *s, int x0, int y0, int x1, int y1,
                              float *out_rgba, uint32_t *out_samples);
/* lic3d_volume_fragment.glsl:2-21 ; out [d][h][w] for z in [z0,z1) (full-size buffer) */
void vvo_lic_volume(const VVOScene *s, int w, int h, int d, int z0, int z1, float *out);
/* raycast_lic3d_fragment.glsl:5-73 */
uint64_t vvo_raycast_licvolume(const VVOScene *s, float *out_rgba, uint32_t *out_samples);
/* lic3d_slicing_fragment.glsl:5-74 over the view-aligned slices of slicing.cpp:42-114 / renderer.cpp:1123-1267 */
uint64_t vvo_slicing_lic(const VVOScene *s, float *out_rgba, uint32_t *out_samples);
/* Slicing WITHOUT the FBO (Renderer::sliceVolume with _useFBO == false, the start-up state; renderer.cpp:1150-1160, 1238-1262):
 * lic3d_slicingblend_fragment.glsl returns the premultiplied sample of every fragment (no dest.a < 0.95 skip), and the GL blends
 * it into the RGBA8 back buffer with (ONE_MINUS_DST_ALPHA, ONE); at the end a white plane is blended in the same way.
 * vvo_slice_fragment_colors: the fragment colours of pixel (x, y), slice order, clamped to [0, 1] as they enter the blend.
 * vvo_slicing_blend8: the back buffer after all slices and the white plane, RGBA8 [h][w][4], each blend computed on the stored
 * 8-bit values and rounded to nearest (GL 2.1 4.1.8, 2.14.9); out_samples counts fragments whose gate passed. */
int vvo_slice_fragment_colors(const VVOScene *s, int x, int y, float *out_rgba, int cap);
uint64_t vvo_slicing_blend8(const VVOScene *s, uint8_t *out_rgba8, uint32_t *out_samples);
void vvo_slicing_setup(const VVOScene *s, float *out5);
int vvo_slice_fragments(const VVOScene *s, int x, int y, float *out_xyzv, int cap);
/* one computeLIC (inc_lic.glsl:152-202) at pos; out[4] */
void vvo_compute_lic(const VVOScene *s, const float pos[3], float out[4]);
/* one direction of the walk, step by step: out[16 i ..] = (newPos.xyz, step.rgb, weighted tap / weight, kernel weight, Pos2.xyz, step2.rgb, 0, 0) */
void vvo_debug_walk(const VVOScene *s, const float pos[3], int dir_sign, int nsteps, float *out);
/* background_fragment.glsl:7-20 and the RGBA8 store (renderer.cpp:216-226) */
void vvo_background(const float *rgba, int n_pixels, float *out);
/* the same pass over a WINDOW of ww x wh pixels showing a stored frame of rw x rh (low-res preset: rw = ww/2, rh = wh/2):
 * texture2DRect(imageFBOSampler, gl_FragCoord.xy * viewport.xy), viewport = (rw / ww, rh / wh) as floats (renderer.cpp:1438-1441),
 * rectangle textures are NEAREST, coordinates clamp to the edge texel */
void vvo_display_window(const float *rgba, int rw, int rh, int ww, int wh, float *out);
void vvo_quantize_rgba8(const float *rgba, int n_values, uint8_t *out);

/* ---- samplers, exposed for unit tests -------------------- */
void vvo_sample_vec(const VVOScene *s, const float p[3], float out[4]);
void vvo_sample_noise(const VVOScene *s, const float p[3], float out[4]);
void vvo_sample_scalar(const VVOScene *s, const float p[3], float out[4]);
float vvo_sample_kernel(const VVOScene *s, float x);
void vvo_sample_tf(const VVOScene *s, float x, float out_rgba[4], float out_la[2]);

/* ---- uniforms / view ------------------------------------------------- */
/* renderer.cpp:925-996 ; out[16] = stepSize, gradient.xyz, licParams.xyz, licKernel.xyz,
 * alphaCorrection, numIterations, h_eff (Q3), 3 spare */
void vvo_derive_uniforms(const VVOScene *s, float *out16);
/* camera.cpp:56-68 + renderer.cpp:146 ; cam_obj[3], rot[9] row-major */
void vvo_view(const VVOScene *s, float cam_obj[3], float rot[9]);
/* renderer.cpp:431-466 ; gl_LightSource[0].position */
void vvo_light_position(const VVOScene *s, float out[4]);
/* per-pixel ray: returns 1 on hit; entry[3] = gl_TexCoord[0], dir[3] = normalize(entry - camera) */
int vvo_pixel_ray(const VVOScene *s, int x, int y, float entry[3], float dir[3]);
void vvo_pixel_rays(const VVOScene *s, int x0, int y0, int x1, int y1, float *out_xyzh);
void vvo_scale_uniforms(const VVOScene *s, float *out9);
/* dataset.cpp:144-176 */
void vvo_volume_geometry(const int size[3], const float slice_dist[3],
                         float extent[3], float scale[3], float scale_inv[3], float center[3]);

/* ---- pre-processing (a13-a17) ---------------------------------------- */
/* dataset.cpp:533-635 (FLOAT3 branch): out RGBA float; fp16 != 0 applies the RGBA16F upload rounding */
void vvo_pack_vector_field(const float *v0, const float *v1, const int dim[3],
                           int interp_index, int interp_size, int fp16, float *out_rgba);
/* dataset.cpp:637-823 non-mutating UCHAR3 formula (Q20) */
void vvo_pack_vector_field_u8(const uint8_t *v0, const int dim[3], int fp16, float *out_rgba);
/* gradient.cpp:190-374, 377-459, 462-532 ; in u8 [z][y][x]; out u8 [z][y][x][3] */
void vvo_noise_gradients(const uint8_t *noise, const int dim[3], const float slice_dist[3], uint8_t *out_grad);
void vvo_compute_gradients_f(const uint8_t *noise, const int dim[3], const float slice_dist[3], float *out);
void vvo_filter_gradients_f(const int dim[3], float *grad);
/* dataset.cpp:1230-1290 */
void vvo_pack_noise_rgba(const uint8_t *noise, const uint8_t *grad, int n, uint8_t *out);
/* dataset.cpp:1159-1162 with an explicit PRNG (mt19937) */
void vvo_white_noise(int n, uint32_t seed, float p, uint8_t *out);
/* dataset.cpp:1405-1512 ; returns padded width; out has room for next_pow2(width) */
int vvo_filter_from_row(const uint8_t *row, int width, int channels, uint8_t *out, float *inv_area);
int vvo_box_filter(int width, uint8_t *out, float *inv_area);
/* transferEdit.cpp:61-98 */
void vvo_default_tf(uint8_t *tf5);
/* illumination.cpp:96-390: zoeckler [h][w][2], mallo diffuse / specular [h][w]; decoded 8-bit values */
void vvo_illum_tables(float spec_exp, int w, int h, float *zoeckler, float *mallo_diff, float *mallo_spec);
/* fp16 round trip */
float vvo_half_round(float x);

int vvo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
