/*
 * vv_c_api.h -- C ABI of libvv_b200.so, the B200-native 3D-LIC renderer.
 *
 * The reference (liaoyg/VectorVisualization, "VV/" below) has no plugin/FFI layer; its de-facto
 * boundary is the public surface of `class Renderer` (VV/renderer.h:28-125) plus the loaders it is
 * handed (VV/dataset.h:54-240, VV/reader.h, VV/parseArg.h, VV/transferEdit.h), all driven by the
 * GLUT callbacks in VV/3DLIC.cpp.  Every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only; the caller keeps ownership of every host buffer it
 * passes in (the library copies to the device), output buffers are caller-allocated.
 *
 * Conventions
 *   - all functions return VV_OK (0) or a negative VVStatus; vv_last_error() gives the message
 *     (the reference prints to stderr and exit(1)s: VV/3DLIC.cpp:690,720 -- this library never exits).
 *   - images are GL-ordered: row 0 is the BOTTOM row, premultiplied RGBA (what
 *     Renderer::saveTexture(_imgBufferTex0) writes, VV/renderer.cpp:340-428,1500).
 *   - volumes are [z][y][x] little-endian, x fastest (VV/reader.cpp:266-305).
 *   - a handle is not re-entrant; use one handle per GPU / host thread.
 *   - there is NO CPU fallback: every compute entry point fails with VV_ERR_CUDA without a device.
 */
#ifndef VV_C_API_H_
#define VV_C_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VV_API __attribute__((visibility("default")))
#else
#define VV_API
#endif

typedef struct VVRenderer VVRenderer;

typedef enum VVStatus {
    VV_OK = 0,
    VV_ERR_INVALID = -1,   /* bad argument / missing input */
    VV_ERR_IO = -2,        /* file not found / parse error */
    VV_ERR_CUDA = -3,      /* CUDA runtime error or no device */
    VV_ERR_STATE = -4      /* call sequence error (e.g. render before data was set) */
} VVStatus;

/* RenderTechnique, VV/types.h:63-71 (same numeric values) */
typedef enum VVTechnique {
    VV_VOLIC_VOLUME = 0,     /* raw vector-field DVR (F1) -- out of scope, rejected */
    VV_VOLIC_RAYCAST = 1,    /* F2: LIC ray-cast, lic3d_fragment.glsl */
    VV_VOLIC_SLICING = 2,    /* F3: view-aligned slicing, lic3d_slicing_fragment.glsl */
    VV_VOLIC_LICVOLUME = 3,  /* F4: ray-cast of the precomputed LIC volume, raycast_lic3d_fragment.glsl */
    VV_VOLIC_VOLUMEANI = 4   /* F5: F4 with per-tick time interpolation + LIC-volume update */
} VVTechnique;

/* struct LICParams, VV/types.h:91-109 (same fields, same defaults via vv_default_lic_params) */
typedef struct VVLicParams {
    float stepSizeVol;     /* 1/128 */
    float gradientScale;   /* 30 */
    float illumScale;      /* 1 */
    float freqScale;       /* 1 */
    int   numIterations;   /* 255 */
    int   stepsForward;    /* 32 */
    int   stepsBackward;   /* 32 */
    float stepSizeLIC;     /* 0.01 */
} VVLicParams;

/* data formats of DatFile, VV/reader.h (DATRAW_*) */
typedef enum VVDataType { VV_UCHAR = 1, VV_USHORT = 2, VV_FLOAT = 3 } VVDataType;

/* Shader-source choices the reference makes by editing/commenting GLSL. */
typedef enum VVTfMode { VV_TF_B = 0, VV_TF_A = 1, VV_TF_R = 2, VV_TF_LENGTH = 3, VV_TF_SCALAR = 4 } VVTfMode;
typedef enum VVGateMode { VV_GATE_ALWAYS = 0, VV_GATE_TF_ALPHA = 1 } VVGateMode;

/* vv_set_option keys */
typedef enum VVOption {
    VV_OPT_TF_MODE = 1,            /* VVTfMode; lic3d_fragment.glsl:53-56 */
    VV_OPT_GATE_MODE = 2,          /* VVGateMode; lic3d_fragment.glsl:59-61 */
    VV_OPT_NOISE_GATE = 3,         /* 1 (default): scalar band (0.1,0.3) gates the noise, inc_lic.glsl:76-89 */
    VV_OPT_QUIRK_SCALEVOLINV = 4,  /* 1 (default): reproduce VV/renderer.cpp:941-944 */
    VV_OPT_QUIRK_LUMINANCE_ALPHA = 5, /* 0 (default): noise .a is the noise value; 1: GL_LUMINANCE .a == 1 (Q7) */
    VV_OPT_LICVOL_FP16 = 6,        /* 1 (default): LIC volume rounded to fp16 like the RGBA16F target (Q14) */
    VV_OPT_FIELD_LAYOUT = 7,       /* 0: float4 [z][y][x]; 1: x-pair-packed fp16 (16 B/voxel); 2: xy-quad-packed fp16 (32 B/voxel, one 256-bit load per
                                      cell face); 3 (default): 2 up to 48 GiB of packed field, else 1.  Same frames and LIC volumes bit for bit. */
    VV_OPT_COUNT_SAMPLES = 8,      /* 1 (default): count ray samples per frame */
    VV_OPT_LICVOL_SIZE = 9,        /* LIC-volume edge length; 0 (default) = field resolution (reference: 512) */
    VV_OPT_SPEC_EXP = 10,          /* gl_LightSource[0].spotExponent as int (default 40, VV/illumination.h:52) */
    VV_OPT_SAMPLE_MAP = 11,        /* 1: keep per-pixel ray-sample counts (vv_read_sample_map) */
    VV_OPT_RAYCAST_MODE = 12,      /* 1 (default): sample-parallel pipeline; 0: one thread per ray (cross-check) */
    VV_OPT_LIC_CTAS_PER_SM = 13,   /* persistent CTAs per SM of the lic_sample kernel (0 = default: all resident) */
    VV_OPT_WALK_FAST_PATHS = 14,   /* 1 (default): unclamped walk inside the field's guard band + shared field / noise cell coordinates where they apply; 0: the clamping samplers (check; same frames) */
    VV_OPT_DEPTH_MAJOR = 15,       /* 1 (default): work items ordered band-major / depth-major for L2 locality; 0: tile-major */
    VV_OPT_BAND_ROWS = 16,         /* 16-pixel block rows per band of the depth-major order (default 4) */
    VV_OPT_FIRST_WINDOW = 18,      /* early-termination frames: ray samples in the first depth window (1..4096, default 2) */
    VV_OPT_WINDOW_GROWTH = 19,     /* ... and the length of every further window in percent of the previous one (100..400, default 300) */
    VV_OPT_PARTITION_UNIT = 20,    /* multi-GPU: the sort-first partition deals units of n x n 16-pixel blocks to the ranks (1, 2, 4 or 8; every rank the same
                                      value, before vv_p2p_export); larger units keep a rank's rays together (L2 locality), smaller ones balance better */
    VV_OPT_NOISE_LAYOUT = 17       /* RGBA (-g) noise: 2 (default) bf16 {t0, t1 - t0}, 1 fp16 x-pair, 0 u8 xy-quad; same values, same frames */
} VVOption;

/* ---- lifecycle: Renderer() / init / resize / ~Renderer, VV/renderer.h:31-37 ------------------- */
VV_API int  vv_create(VVRenderer **out, int cuda_device);
VV_API void vv_destroy(VVRenderer *r);
/* Renderer::init(char* defines) / loadGLSLShader(char* defines), VV/renderer.h:34,115: accepts the strings
 * the keyboard handler passes (VV/3DLIC.cpp:416-436): "#define ILLUM_GRADIENT", "#define ILLUM_MALLO",
 * "#define ILLUM_ZOECKLER", "#define SPEED_OF_FLOW", NULL/"" = plain illumLIC; selects kernel variants. */
VV_API int  vv_init(VVRenderer *r, const char *defines);
VV_API int  vv_load_glsl_shader(VVRenderer *r, const char *defines);
VV_API int  vv_resize(VVRenderer *r, int width, int height);
VV_API const char *vv_last_error(void);
VV_API const char *vv_version(void);

/* ---- technique: setTechnique / updateLICVolume / updateSlices, VV/renderer.h:44,119-123 ------- */
VV_API int vv_set_technique(VVRenderer *r, int technique);
VV_API int vv_update_lic_volume(VVRenderer *r);
VV_API int vv_update_slices(VVRenderer *r);

/* ---- inputs: setVolumeData/setDataTex (VectorDataSet), VV/renderer.h:48-60, VV/dataset.cpp:97-366 --- */
/* FLOAT3 / UCHAR3 vector field; `next` may be NULL (single time step).  Packs to the RGBA16F texture
 * contents of VectorDataSet::createTextureIterp (VV/dataset.cpp:290-366, 533-635) on the GPU: bit-identical for FLOAT3.
 * UCHAR3 is a stated DEVIATION: the path the reference runs subtracts 128 in `unsigned char` IN PLACE on every call
 * (VV/dataset.cpp:563-571: u = 100 -> 228, never negative, and the source is mutated again by the next frame); here a
 * UCHAR3 component is the signed (float)u - 128 of the reference's non-mutating fillTexDataFloat (VV/dataset.cpp:474-490). */
VV_API int vv_set_vector_field(VVRenderer *r, const void *data, const void *next, int dtype,
                               const int dims[3], const float slice_dist[3]);
/* interpIndex / InterpSize of VectorDataSet (VV/dataset.cpp:202-210, VV/3DLIC.cpp:705): re-packs on the GPU */
VV_API int vv_set_time_interp(VVRenderer *r, int interp_index, int interp_size);
/* Animation tick.  idle() (VV/3DLIC.cpp:129-172) calls, per tick, VectorDataSet::createTextureIterp -- pack with the fraction
 * interpIndex / InterpSize, then interpIndex++ (VV/dataset.cpp:590-633) -- and checkInterpolateStage -- once interpIndex reaches
 * InterpSize: data <- time step getNextTimeStep() (wrapping from the last to the first), newData <- the step after it,
 * interpIndex <- 0 (VV/dataset.cpp:202-210, VV/reader.cpp:327-336).  VVTimeCursor is that bookkeeping as plain host logic;
 * vv_time_cursor_tick returns the fraction index this tick's texture is packed with and sets *advanced when the pair of time
 * steps moved on (data = step cursor->current, newData = step vv_time_cursor_next).
 * vv_idle does the whole tick on a handle whose field came from vv_load_dat: re-pack on the GPU, advance, and re-read the two
 * RAW files when the pair moves on.  Returns VV_OK, or VV_ERR_STATE when the field was not loaded from a DAT file.
 * (FLOAT3 fields; a UCHAR3 field keeps its first time step: the reference's UCHAR interpolation path corrupts its own source
 * array on every call, SURVEY Q20, and is not reproduced.) */
typedef struct VVTimeCursor { int time_begin, time_end, current, interp_index, interp_size; } VVTimeCursor;
VV_API void vv_time_cursor_init(VVTimeCursor *c, int time_begin, int time_end, int interp_size);
VV_API int vv_time_cursor_next(const VVTimeCursor *c);
VV_API int vv_time_cursor_tick(VVTimeCursor *c, int *advanced);
VV_API int vv_idle(VVRenderer *r);
VV_API int vv_get_time_cursor(VVRenderer *r, VVTimeCursor *out);
/* setScalarTex (VolumeDataSet, VV/dataset.cpp:840-1050); UCHAR or FLOAT scalar */
VV_API int vv_set_scalar(VVRenderer *r, const void *data, int dtype, const int dims[3]);
/* setNoiseTex (NoiseDataSet, VV/dataset.cpp:1124-1344): u8 noise; with_gradients = the `-g` flag:
 * Sobel + 5^3 smoothing + quantise on the GPU (VV/gradient.cpp:190-532) and RGBA8 packing. */
VV_API int vv_set_noise(VVRenderer *r, const uint8_t *data, const int dims[3], int with_gradients);
/* NoiseDataSet::loadData fallback (VV/dataset.cpp:1142-1163): n^3 white noise, P(255) = p, mt19937(seed) */
VV_API int vv_generate_white_noise(VVRenderer *r, int n, uint32_t seed, float p, int with_gradients);
/* setLICFilter (LICFilter, VV/dataset.cpp:1405-1512): first row of a kernel image / box filter */
VV_API int vv_set_filter(VVRenderer *r, const uint8_t *row, int width, int channels);
VV_API int vv_set_box_filter(VVRenderer *r, int width);
/* setTFrgbTex + setTFalphaOpacTex (TransferEdit, VV/transferEdit.cpp:61-98,480-543): 256 x (R,G,B,A,opacity) */
VV_API int vv_set_tf(VVRenderer *r, const uint8_t *tf256x5);
VV_API int vv_set_default_tf(VVRenderer *r);
/* setLICParams, VV/renderer.h:103 */
VV_API int vv_set_lic_params(VVRenderer *r, const VVLicParams *p);
VV_API void vv_default_lic_params(VVLicParams *p);
/* setCamera (Camera, VV/camera.cpp:42-68): quaternion (x,y,z,w), translation, distance, fovy (deg), near, far (the
 * reference's defaults: 0.1, 50).  near / far bound the view volume the GL clips the proxy geometry to: a pixel whose ray enters
 * the box (or a slice / cap polygon) at an eye-space depth outside [near, far] has no fragment, as in the reference. */
VV_API int vv_set_camera(VVRenderer *r, const float quat[4], const float pos[3], float dist, float fovy,
                         float near_clip, float far_clip);
/* setLight + updateLightPos (Transform, VV/renderer.cpp:431-466) */
VV_API int vv_set_light(VVRenderer *r, const float quat[4], float dist);
VV_API int vv_update_light_pos(VVRenderer *r);
/* enableLowRes / enableFBO, VV/renderer.h:76-83.  Low-res = the interaction preset of VV/renderer.cpp:947-966 (step x2,
 * 15+15 LIC steps of 1/64, frequency x0.7, alpha correction for the doubled step).  The reference additionally renders
 * into half the window in that mode (VV/renderer.cpp:111-119, 159-160) while gluPerspective keeps the WINDOW's aspect ratio
 * (Camera::setWindow, VV/transform.h:79-80; VV/3DLIC.cpp:192): call vv_set_window(w, h) + vv_resize(max(w/2,1), max(h/2,1))
 * for that.  vv_set_window(0, 0) returns to the default, aspect = frame width / frame height. */
VV_API int vv_enable_lowres(VVRenderer *r, int enable);
VV_API int vv_set_window(VVRenderer *r, int window_width, int window_height);
/* enableFBO (key 'F'): the stored frame lives in a GL_RGBA16F_ARB texture instead of the RGBA8 back buffer (VV/renderer.cpp:
 * 562-606, Q18).  With it on, vv_read_rgba32f returns the frame rounded to fp16 and the stored-frame PNG (vv_save_png(.., 0),
 * screenshots, recordings) is written the way saveTexture converts a float texture: (int)(255 * texel), truncated
 * (VV/renderer.cpp:386-403).  vv_read_rgba8 is always the RGBA8 back-buffer frame. */
VV_API int vv_enable_float_target(VVRenderer *r, int enable);
VV_API int vv_set_option(VVRenderer *r, int option, int value);
/* Screenshot / recording: Renderer::screenshot / switchRecording (VV/renderer.h:99-101, keys '0' / 'R', VV/3DLIC.cpp:259-270)
 * and the tail of Renderer::renderFBO (VV/renderer.cpp:1478-1513): after a screenshot request, and for every vv_render
 * while recording, the stored RGBA8 frame is written as PNG to "<dir>/<frames>_<file_name>" (recording or animation on;
 * frames counts the recorded frames) or "<dir>/<dd-mm-YYYY HH-MM-SS> <file_name>".  Defaults: dir "snapshotOut",
 * file_name "snapshot.png".  vv_switch_recording returns the new state (1 = recording). */
VV_API int vv_set_snapshot(VVRenderer *r, const char *dir, const char *file_name, int animation_on);
VV_API int vv_screenshot(VVRenderer *r);
VV_API int vv_switch_recording(VVRenderer *r);
VV_API const char *vv_last_snapshot_path(VVRenderer *r);
/* Monte-Carlo ray-start offsets.  Renderer::updateMCOffsetTex (VV/renderer.cpp:636-679) fills a width x height
 * GL_LUMINANCE16F rectangle texture with rand()/RAND_MAX; programs built with "#define USE_MC_OFFSET" (passed to vv_init /
 * vv_load_glsl_shader) start each ray / slice fragment at pos + dir * stepSize * offset[pixel]
 * (lic3d_fragment.glsl:31-33, lic3d_slicing_fragment.glsl:31-33).  vv_set_mc_offsets takes the values (row 0 = bottom
 * row, rounded to fp16 like the upload; NULL removes the texture); vv_update_mc_offset_tex draws them from
 * mt19937(seed).  The size must equal the frame size at vv_render. */
VV_API int vv_set_mc_offsets(VVRenderer *r, const float *offsets, int width, int height);
VV_API int vv_update_mc_offset_tex(VVRenderer *r, int width, int height, uint32_t seed);
/* ---- key map: keyboard / keyboardSpecial of VV/3DLIC.cpp:243-488 -------------------------------------------------------------
 * VVAppState is the counterpart of the application's globals (licParams, renderTechnique, animationMode, updateSceneCont,
 * clipPlanes[i]._active, currentClipPlane, VV/3DLIC.h:29-55) and is owned by the caller, as they are there.
 * vv_key_apply is plain host logic (no device): it applies one key to the state with the reference's increments and clamps
 * ('[' ']' sample distance /2 x2 in [0,1]; 's' 'x' 'S' 'X' LIC steps +-1 >= 1; 'a' 'z' LIC step x2 /2 >= 0.0005; 'h' 'n' noise
 * frequency +-0.2 >= 0.5; 'j' 'm' illumination scale +-0.05 >= 0.05; 'g' 'b' gradient scale +-0.2 >= 0.2; 'L' low-res; 'F' float
 * target; ' ' continuous; '1'-'4' clip planes; '6'-'9', '.', 'r' shader defines; 'u' LIC volume; '0' screenshot; 'R' recording;
 * special keys 1..5 = F1..F5 technique / animation) and returns a mask of VVKeyAction saying what the reference does next.
 * vv_keyboard applies the key and carries those actions out on the handle (returns the mask, or -1 with vv_last_error set).
 * 'q' / Esc return VV_KEY_QUIT instead of calling exit(1). */
typedef struct VVAppState {
    VVLicParams lic;
    int technique;                  /* VVTechnique; VV_VOLIC_VOLUME at start-up (VV/3DLIC.h:51) */
    int lowres, float_target, continuous, recording, animation, screenshot;
    int clip_active[3], selected_clip;   /* selected_clip: 0..2 or -1 */
    char defines[64];               /* the string of the last loadGLSLShader(defines) request, "" = none */
} VVAppState;
typedef enum VVKeyAction {
    VV_KEY_UPDATE_SCENE = 1, VV_KEY_RELOAD_SHADER = 2, VV_KEY_UPDATE_LICVOLUME = 4, VV_KEY_UPDATE_SLICES = 8, VV_KEY_QUIT = 16,
    VV_KEY_SCREENSHOT = 32, VV_KEY_SWITCH_RECORDING = 64, VV_KEY_SET_TECHNIQUE = 128
} VVKeyAction;
VV_API void vv_app_state_init(VVAppState *s);
VV_API int vv_key_apply(VVAppState *s, int key, int special);
VV_API int vv_keyboard(VVRenderer *r, VVAppState *s, int key, int special);

/* (Mouse interaction -- trackball, translate, dolly of VV/3DLIC.cpp:490-600 -- is caller-side state handling and not part of this
 * library: a caller hands its camera, light and clip planes to vv_set_camera / vv_set_light / vv_set_clip_plane.) */

/* User clip planes: ClipPlane::setNormal(x, y, z, d) + activation (VV/transform.cpp:296-315, 446-483), index 0..2 =
 * GL_CLIP_PLANE0 + index (VV/3DLIC.cpp:763-781).  equation = (n.xyz, d) in volume-centred object coordinates, the
 * half-space n.q + d >= 0 is kept; NULL keeps the stored equation.  Semantics as drawn by Renderer::render
 * (VV/renderer.cpp:156-163, 1294-1309): the front faces of the proxy cube and the slice polygons are clipped, and for
 * the two ray-cast techniques each active plane adds the box cross-section n^.q = -(d - 0.0001) as a ray-entry polygon
 * (culled when it faces away); rays still leave through the box, as in the reference.  The plane is evaluated as
 * (n / |n|, d): the reference normalises the normal in place the first time it draws the plane's cap (VV/renderer.cpp:1301
 * hands ClipPlane::getNormal() to ViewSlicing::setupSingleSlice, VV/slicing.cpp:337-348), d untouched, so that is the plane
 * every frame but the first is clipped with (the reference's own callers only ever pass unit normals). */
VV_API int vv_set_clip_plane(VVRenderer *r, int index, const double equation[4], int active);
/* setIllum*Tex (Illumination, VV/illumination.cpp:96-390): the Zoeckler / Mallo look-up tables are generated inside the
 * library when an ILLUM_MALLO / ILLUM_ZOECKLER build is selected; this host-only entry point returns the same tables
 * (decoded UNORM8 values): zoeckler [h][w][2] (luminance, alpha), mallo [h][w] each */
VV_API int vv_make_illum_tables(float spec_exp, int width, int height, float *zoeckler_la, float *mallo_diffuse, float *mallo_specular);

/* ---- frame: Renderer::render(update), VV/renderer.cpp:126-312 -------------------------------- */
VV_API int vv_render(VVRenderer *r, int update);
/* stored frame (_imgBufferTex0): RGBA8 (default back-buffer path, Q18) or RGBA32F */
VV_API int vv_read_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes);
VV_API int vv_read_rgba32f(VVRenderer *r, float *out, size_t out_bytes);
/* displayed frame: background_fragment.glsl:9-16 composited over white */
VV_API int vv_read_display_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes);
/* the same pass over a window that is larger than the stored frame (low-res preset: the frame is half the window): the shader
 * reads texture2DRect(imageFBOSampler, gl_FragCoord.xy * viewport.xy), viewport = (frame / window) per axis (VV/renderer.cpp:
 * 1436-1441), i.e. a NEAREST, edge-clamped up-scaling; out = window_width * window_height RGBA8, rows bottom-up like the frame */
VV_API int vv_read_display_window_rgba8(VVRenderer *r, uint8_t *out, size_t out_bytes, int window_width, int window_height);
/* LIC volume contents (fp32 scalar [d][h][w]) */
VV_API int vv_read_lic_volume(VVRenderer *r, float *out, size_t out_bytes, int dims_out[3]);
/* texture read-back (glGetTexImage equivalents, used by the parity tests):
 * vector texture as float RGBA [z][y][x][4] (the RGBA16F contents), noise texture as L8 or RGBA8 */
VV_API int vv_read_field_texture(VVRenderer *r, float *out_rgba, size_t out_bytes);
VV_API int vv_read_noise_texture(VVRenderer *r, uint8_t *out, size_t out_bytes, int *channels);
/* per-pixel ray-sample counts of the last frame (needs VV_OPT_SAMPLE_MAP = 1 before vv_render) */
VV_API int vv_read_sample_map(VVRenderer *r, uint32_t *out, size_t out_bytes);
/* saveFrameBuffer / saveTexture, VV/renderer.h:39-42 */
VV_API int vv_save_png(VVRenderer *r, const char *path, int displayed);
VV_API int vv_save_raw(VVRenderer *r, const char *path);

/* ---- measurement ------------------------------------------------------------------------------ */
VV_API uint64_t vv_last_ray_samples(VVRenderer *r);   /* ray samples of the last vv_render */
VV_API float    vv_last_kernel_ms(VVRenderer *r);     /* CUDA-event time of the dominant kernel, last frame */
VV_API int      vv_last_launch_count(VVRenderer *r);  /* kernels launched by the last vv_render */
VV_API int      vv_field_layout(VVRenderer *r);       /* layout the vector field is packed in (VV_OPT_FIELD_LAYOUT resolved: 0, 1 or 2) */
VV_API int      vv_synchronize(VVRenderer *r);

/* ---- device-side / multi-GPU hooks (pointers are CUDA device pointers on the handle's device) --- */
/* sort-first partition: this handle renders only blocks b with b % world == rank (16x16-pixel blocks,
 * row-major block index).  world = 1 restores the whole image. */
VV_API int vv_set_partition(VVRenderer *r, int rank, int world);
/* LIC-volume output slab [z0,z1) computed by vv_update_lic_volume on this handle (input field replicated) */
VV_API int vv_set_licvol_slab(VVRenderer *r, int z0, int z1);
/* compact block-major tile buffer of the last frame: n_blocks x 256 x float4 (device pointer) */
VV_API int vv_get_tile_buffer(VVRenderer *r, void **dev_ptr, int *n_local_blocks, int *n_total_blocks);
/* assemble a row-major frame on this handle from `world` gathered tile buffers laid out [rank][block][256][4] */
VV_API int vv_assemble_tiles(VVRenderer *r, const void *gathered_dev, int world);
VV_API int vv_get_lic_volume_ptr(VVRenderer *r, void **dev_ptr, int dims_out[3]);
/* Diagnostic: one direction (dir_sign < 0 backward, else forward) of computeLIC's streamline walk (inc_lic.glsl:104-145) from the
 * texture-space position pos, with the device functions of the hot path; out[16 i ..] = (newPos.xyz, step.rgb, noise tap, kernel
 * weight, Pos2.xyz, step2.rgb, 0, 0) of step i.  walk_variant: 0 clamping samplers, 1 guard band, 3 guard band + shared field / noise cell. */
VV_API int vv_debug_walk(VVRenderer *r, const float pos[3], int dir_sign, int n_steps, int walk_variant, float *out, size_t out_bytes);
/* Peer-to-peer frame exchange for one process per GPU on one NVLink / NVSwitch node (no reference counterpart: the
 * reference is single-GPU).  Instead of gathering tile buffers with a collective, vv_p2p_render stores this rank's
 * finished tiles straight into every rank's gather buffer over NVLink, signals arrival with system-scope atomics, waits
 * for the other ranks' arrivals and un-blocks the frame, all on the handle's stream.  Protocol, after vv_resize and
 * vv_set_partition on every rank: vv_p2p_export (allocates the gather buffer, returns its 64-byte cudaIpcMemHandle_t
 * and/or base pointer) -> exchange the handles between the ranks (any host channel) -> vv_p2p_connect with the `world`
 * handles concatenated in rank order (or base pointers for handles living in the same process) -> vv_p2p_render once
 * per frame on EVERY rank (it is collective).  A rank that never arrives makes the others give up after about 17 s;
 * vv_p2p_status then reports VV_ERR_STATE. */
VV_API int vv_p2p_export(VVRenderer *r, void *ipc_handle_out64, void **base_out);
VV_API int vv_p2p_connect(VVRenderer *r, const void *ipc_handles, void *const *local_bases, int world);
VV_API int vv_p2p_render(VVRenderer *r);
VV_API int vv_p2p_status(VVRenderer *r);
VV_API int vv_p2p_disconnect(VVRenderer *r);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the handle's own stream */
VV_API int vv_set_stream(VVRenderer *r, void *cuda_stream);

/* ---- loaders: DatFile / NoiseDataSet / LICFilter / TransferEdit / ParseArguments ---------------- */
typedef struct VVDatInfo {          /* DatFile, VV/reader.cpp:81-263 */
    char raw_file[512];             /* resolved ObjectFileName (printf pattern for time-dependent sets) */
    int  resolution[3];
    float slice_thickness[3];
    int  data_type;                 /* VVDataType */
    int  data_dim;                  /* 1 scalar, 3 vector */
    int  time_begin, time_end;      /* TimeDependent: b e */
} VVDatInfo;
VV_API int vv_parse_dat(const char *dat_path, VVDatInfo *out);
/* DatFile::readRawData(timeStep), VV/reader.cpp:266-305: reads into caller buffer */
VV_API int vv_read_raw(const VVDatInfo *info, int time_step, void *out, size_t out_bytes);
VV_API int vv_load_dat(VVRenderer *r, const char *dat_path);          /* vector field (.dat, FLOAT3/UCHAR3) */
VV_API int vv_load_scalar_dat(VVRenderer *r, const char *dat_path);   /* scalar volume (.dat, UCHAR/FLOAT) */
/* 3 x int32 + u8, VV/dataset.cpp:1347-1389.  With gradients the quantised gradients are cached next to the noise file as
 * NoiseDataSet::createTexture does (VV/dataset.cpp:1238-1267): "<path>.grd" is used when it exists and has the right
 * size, otherwise the gradients are computed (on the GPU) and written there; a failed write is not an error. */
VV_API int vv_load_noise(VVRenderer *r, const char *path, int with_gradients);
/* the gradient cache itself: loadGradients / saveGradients(DATRAW_UCHAR), VV/gradient.cpp:93-187 -- "<file_name>.grd" holds
 * 3 * nx*ny*nz bytes, [z][y][x][3].  vv_grd_read returns VV_ERR_IO when the file is missing or short.  Host-only. */
VV_API int vv_grd_read(const char *file_name, const int dims[3], uint8_t *gradients3);
VV_API int vv_grd_write(const char *file_name, const int dims[3], const uint8_t *gradients3);
/* noise volume + already quantised gradients -> the RGBA8 noise texture (VV/dataset.cpp:1269-1282) */
VV_API int vv_set_noise_with_gradients(VVRenderer *r, const uint8_t *noise, const uint8_t *gradients3, const int dims[3]);
/* read back the quantised gradients of the current noise ([z][y][x][3]) */
VV_API int vv_read_noise_gradients(VVRenderer *r, uint8_t *out, size_t out_bytes);
VV_API int vv_load_filter_png(VVRenderer *r, const char *path);       /* LICFilter::loadData, VV/dataset.cpp:1415-1467 */
VV_API int vv_load_tf_png(VVRenderer *r, const char *name);           /* TransferEdit::loadTF, VV/transferEdit.cpp:224-337 */

typedef struct VVArgs {             /* ParseArguments, VV/parseArg.h:43-51 */
    char vol_file[512], noise_file[512], tf_file[512], filter_file[512], redirect_file[512], halton_file[512];
    int  use_gradients, use_lambda2, show_help;
} VVArgs;
/* ParseArguments::parse, VV/parseArg.cpp:97-365: returns VV_OK or VV_ERR_INVALID (the reference prints usage
 * and exit(1)s, VV/3DLIC.cpp:848-852); -h/--help sets show_help instead of exit(0). */
VV_API int vv_parse_args(int argc, const char *const *argv, VVArgs *out);
VV_API const char *vv_usage(void);

/* standalone PNG helpers (imageUtils pngRead/pngWrite semantics: 8-bit gray/GA/RGB/RGBA, row 0 = top) */
VV_API int vv_png_read(const char *path, uint8_t **data, int *w, int *h, int *channels);
VV_API void vv_free(void *p);
VV_API int vv_png_write(const char *path, const uint8_t *data, int w, int h, int channels);

#ifdef __cplusplus
}
#endif
#endif /* VV_C_API_H_ */
