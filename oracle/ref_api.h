/* ref_api.h -- TEST INFRASTRUCTURE ONLY: C interface of oracle/_ref/libvv_ref.so (see build_ref.py) */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefTex {
    const void *data;
    int dim[3];
    int fmt;    /* glsl::Fmt  */
    int wrap;   /* glsl::Wrap */
} RefTex;

typedef struct RefUniforms {
    RefTex volume, scalar, noise, kernel, tf_rgba, tf_alphaopac, licvol, zoeckler, mallo_diff, mallo_spec;
    float texMax[4], scaleVol[4], scaleVolInv[4];
    float stepSize;
    float gradient[3];
    int numIterations;
    float alphaCorrection;
    float licParams[3];
    float licKernel[3];
    float camera[4];              /* gl_ModelViewMatrixInverse[3] */
    float light_position[4];      /* gl_LightSource[0].position */
    float light_ambient[4], light_diffuse[4], light_specular[4];
    float spot_exponent;
    RefTex mc_offset;             /* mcOffsetSampler: F_L32F [frame height][frame width] (USE_MC_OFFSET programs) */
    int frag_x0, frag_y0, frag_w; /* fragment i of a run is pixel (frag_x0 + i % frag_w, frag_y0 + i / frag_w); frag_w = 0: unset */
    /* the frame-buffer side of the FBO slicing pass (Renderer::sliceVolume, VV/renderer.cpp:1176-1225), which the reference leaves
     * to the GL: slice_count > 0 = two ping-pong targets per pixel, a fragment's .w is its slice index (slice i samples the target
     * of slice i - 1 and is written to the other one; the frame is the target of slice slice_count - 1); 0 = one accumulator.
     * fbo_fp16: the targets are GL_RGBA16F_ARB (VV/renderer.cpp:566-606), every write rounds to fp16. */
    int slice_count, fbo_fp16;
} RefUniforms;

/* sets the gl_* state shared by all programs */
void vvref_set_gl_state(const RefUniforms *u);

#ifdef __cplusplus
}
#endif
