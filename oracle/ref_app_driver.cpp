/* ref_app_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * VV/3DLIC.cpp -- the reference's application file with its GLUT callbacks -- compiled UNMODIFIED against oracle/ref_shim/
 * (Windows.h, GL\glew.h, GL\freeglut.h stand-ins).  Nothing opens a window or enters a main loop: this driver wires the
 * application's own global objects the way its init() does (VV/3DLIC.cpp:677-799), then calls keyboard() / keyboardSpecial()
 * directly and reads the globals back.  What is checked against it: the key map of the product's vv_keyboard /
 * vv_keyboard_special (tests/test_host_vs_ref.py).
 * The HUD (VV/hud.cpp, out of scope: VBO text rendering) is replaced by a recorder of the status line it is handed. */
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "GL/glew.h"
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#define private public
#define protected public
#include "renderer.h"
#include "transform.h"
#undef private
#undef protected
#include "camera.h"
#include "hud.h"
#include "transferEdit.h"
#include "imageUtils.h"

/* ---- the application's globals (defined in VV/3DLIC.h, which only VV/3DLIC.cpp includes) ---- */
extern Camera cam;
extern Renderer renderer;
extern ClipPlane clipPlanes[3];
extern ClipPlane *currentClipPlane;
extern Transform light;
extern VectorDataSet vd;
extern LICFilter licFilter;
extern LICParams licParams;
extern RenderTechnique renderTechnique;
extern bool animationMode, updateSceneCont, updateScene, screenshot, lightVisible, wire, useIdle, requestHighRes;
void keyboard(unsigned char key, int x, int y);
void keyboardSpecial(int key, int x, int y);
void resize(int width, int height);

/* ---- HUD stand-in: keeps the status line of updateHUD (VV/3DLIC.cpp:203-240) ---- */
static std::string g_hud_text;
OpenGLHUD::OpenGLHUD(bool, bool) {}
OpenGLHUD::~OpenGLHUD() {}
bool OpenGLHUD::Init() { return true; }
void OpenGLHUD::SetNumLines(int, bool) {}
void OpenGLHUD::SetViewport(int *, bool) {}
void OpenGLHUD::SetText(const char *text, bool) { g_hud_text = text ? text : ""; }
void OpenGLHUD::DrawHUD(const char *, unsigned int, unsigned int) {}

/* the shader sources a key hands to the GL: the defines string arrives as its own source string (VV/GLSLShader.cpp:120-180) */
static std::string g_last_defines;
static int g_shader_loads = 0;
extern "C" void glShaderSourceARB(GLhandleARB, GLsizei n, const GLcharARB **src, const GLint *)
{
    g_last_defines.clear();                                   /* the defines of the LAST source hand-over ("" = none) */
    for (int i = 0; i < n; ++i)
        if (src && src[i] && !std::strncmp(src[i], "#define", 7)) g_last_defines = src[i];
    ++g_shader_loads;
}

static bool g_wired = false;
static int g_modifiers = 0;
extern "C" int glutGetModifiers(void) { return g_modifiers; }
void mouseInteract(int button, int state, int x, int y);
void mouseMotionInteract(int x, int y);

extern "C" {

static int wire_app(const char *dat);
static int keyboard_impl(const char *ref_dir, const unsigned char *keys, const int *special, int n, float *out,
                         char *defines_out, int defines_cap, char *hud_out, int hud_cap);

/* keys[i] is fed to keyboard() (special[i] == 0) or keyboardSpecial() (special[i] != 0; 1..5 = GLUT_KEY_F1..F5) after the
 * application state has been reset to its start-up values.  out[0..7] = LICParams (stepSizeVol, gradientScale, illumScale,
 * freqScale, numIterations, stepsForward, stepsBackward, stepSizeLIC); [8] technique; [9] low-res; [10] FBO; [11] recording;
 * [12] animation flag; [13..15] clip plane active; [16] selected clip plane or -1; [17] screenshot requested;
 * [18] shader (re)loads triggered; [19] continuous mode; [20] frame store.  defines_out / hud_out receive the last "#define ..."
 * string handed to glShaderSourceARB (empty: none since reset) and the HUD status line. */
int vvref_keyboard(const char *dat, const char *ref_dir, const unsigned char *keys, const int *special, int n, float *out,
                   char *defines_out, int defines_cap, char *hud_out, int hud_cap)
{
    if (wire_app(dat) != 0) return -10;
    return keyboard_impl(ref_dir, keys, special, n, out, defines_out, defines_cap, hud_out, hud_cap);
}

static int wire_app(const char *dat)
{
    static Texture dummy[10];
    if (!g_wired) {
        if (!vd.loadData(dat)) return -10;
        licFilter.createBoxFilter();
        light.setDistance(1.0f);
        renderer.setLight(&light);
        renderer.setCamera(&cam);
        VolumeData *v = vd.getVolumeData();
        for (int i = 0; i < 3; ++i) {
            clipPlanes[i].setPlaneId(GL_CLIP_PLANE0 + i);
            clipPlanes[i].setBoundingBox(-v->extent[0] / 2.0f, -v->extent[1] / 2.0f, -v->extent[2] / 2.0f, v->extent[0] / 2.0f, v->extent[1] / 2.0f, v->extent[2] / 2.0f);
        }
        renderer.setClipPlanes(clipPlanes, 3);
        renderer.setVolumeData(v);
        renderer.setLICFilter(&licFilter);
        renderer.setDataTex(&dummy[0]); renderer.setScalarTex(&dummy[1]); renderer.setNoiseTex(&dummy[2]);
        renderer.setTFrgbTex(&dummy[3]); renderer.setTFalphaOpacTex(&dummy[4]);
        renderer.setIllumZoecklerTex(&dummy[5]); renderer.setIllumMalloDiffTex(&dummy[6]); renderer.setIllumMalloSpecTex(&dummy[7]);
        renderer._licKernelTex = &dummy[8];
        renderer.setLICParams(&licParams);
        renderer._licvolumebuffer = new VolumeBuffer(GL_RGBA16F_ARB, 4, 4, 4, 2);   /* Renderer::init allocates 512^3 (VV/renderer.cpp:97) */
        resize(64, 48);
        g_wired = true;
    }
    return 0;
}

static int keyboard_impl(const char *ref_dir, const unsigned char *keys, const int *special, int n, float *out,
                         char *defines_out, int defines_cap, char *hud_out, int hud_cap)
{
    /* start-up state: LICParams ctor (VV/types.h:91-109), VV/3DLIC.h:29-55 */
    licParams = LICParams();
    renderTechnique = VOLIC_VOLUME;
    renderer.setTechnique(renderTechnique);
    renderer._lowRes = false; renderer._useFBO = false; renderer._recording = false; renderer._screenShot = false;
    renderer._isAnimationOn = false; renderer._storeFrame = true; renderer._wireframe = false;
    for (int i = 0; i < 3; ++i) { clipPlanes[i]._active = false; clipPlanes[i]._visible = false; }
    currentClipPlane = NULL;
    animationMode = false; updateSceneCont = false; updateScene = true; screenshot = false; lightVisible = false; wire = false;
    requestHighRes = false;
    g_last_defines.clear(); g_shader_loads = 0; g_hud_text.clear();
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -11;
    if (ref_dir && chdir(ref_dir) != 0) return -12;        /* "shader/..." is opened relative to the working directory */
    for (int i = 0; i < n; ++i) {
        if (special[i]) keyboardSpecial(keys[i], 0, 0);
        else keyboard(keys[i], 0, 0);
    }
    if (chdir(cwd) != 0) return -13;
    out[0] = licParams.stepSizeVol; out[1] = licParams.gradientScale; out[2] = licParams.illumScale; out[3] = licParams.freqScale;
    out[4] = (float)licParams.numIterations; out[5] = (float)licParams.stepsForward; out[6] = (float)licParams.stepsBackward;
    out[7] = licParams.stepSizeLIC;
    out[8] = (float)(int)renderer._renderMode;
    out[9] = renderer._lowRes; out[10] = renderer._useFBO; out[11] = renderer._recording; out[12] = renderer._isAnimationOn;
    for (int i = 0; i < 3; ++i) out[13 + i] = clipPlanes[i]._active;
    out[16] = currentClipPlane ? (float)(currentClipPlane - clipPlanes) : -1.0f;
    out[17] = renderer._screenShot;
    out[18] = (float)g_shader_loads;
    out[19] = updateSceneCont;
    out[20] = renderer._storeFrame;
    if (defines_out && defines_cap > 0) { std::strncpy(defines_out, g_last_defines.c_str(), defines_cap - 1); defines_out[defines_cap - 1] = 0; }
    if (hud_out && hud_cap > 0) { std::strncpy(hud_out, g_hud_text.c_str(), hud_cap - 1); hud_out[hud_cap - 1] = 0; }
    return 0;
}

/* The animation as init() + idle() drive it (VV/3DLIC.cpp:686-707, 129-142): after loadData the two first time steps are
 * loaded, createTextureIterp + checkInterpolateStage run once, and then once per idle tick.  For tick i (0-based) out[3i] =
 * the current time step and out[3i+1] = interpIndex when the tick starts (the fraction index its texture is packed with),
 * out[3i+2] = 1 if checkInterpolateStage moved to the next pair of time steps during the tick.  The texture uploaded by tick
 * `tex_tick` (RGBA float, as handed to glTexImage3D) is copied to tex_out. */
int vvref_animation_ticks(const char *dat, int n, int *out, int tex_tick, float *tex_out, size_t tex_bytes)
{
    VectorDataSet v;
    if (!v.loadData(dat)) return -10;
    v.getVolumeData()->data = v.loadTimeStep(v.getCurTimeStep());
    v.getVolumeData()->newData = v.loadTimeStep(v.NextTimeStep());
    v.setInterpolateSize(10);
    v.createTextureIterp("VectorData_Tex", GL_TEXTURE2_ARB, true);
    v.checkInterpolateStage();
    for (int i = 0; i < n; ++i) {
        const int step = v.getCurTimeStep(), idx = v.interpIndex;
        v.createTextureIterp("VectorData_Tex", GL_TEXTURE2_ARB, true);
        if (i == tex_tick && tex_out) {
            VVStubTex *t = vv_stub_texture(vv_stub_last_texture());
            if (!t || t->bytes > tex_bytes) return -11;
            std::memcpy(tex_out, t->data, t->bytes);
        }
        v.checkInterpolateStage();
        out[3 * i] = step; out[3 * i + 1] = idx; out[3 * i + 2] = (v.getCurTimeStep() != step) ? 1 : 0;
    }
    return 0;
}

/* Mouse interaction: mouseInteract / mouseMotionInteract (VV/3DLIC.cpp:490-600) fed with events[i] = (type, button, x, y,
 * modifiers): type 0 = button press (GLUT_DOWN), type 1 = motion, type 2 = select clip plane `button` (-1: none; what keys
 * '1'-'4' do to currentClipPlane), on freshly constructed cam / light / clipPlanes set up as init() does (VV/3DLIC.cpp:681,
 * 763-781) in a width x height window.  out: per object (camera, light, clip plane 0..2) 13 floats: _q_internal[4], _q[4],
 * _dist, _pos[3] (camera only, else 0), locked; then the three plane equations as 12 doubles in out_normals. */
int vvref_mouse(const char *dat, int width, int height, const int *events, int n, float *out, double *out_normals)
{
    if (wire_app(dat) != 0) return -10;
    cam.~Camera(); new (&cam) Camera();
    light.~Transform(); new (&light) Transform();
    for (int i = 0; i < 3; ++i) { clipPlanes[i].~ClipPlane(); new (&clipPlanes[i]) ClipPlane(); }
    cam.setPosition(Vector3_new(0.0f, 0.0f, 0.0f));         /* a global there: zero-initialised */
    light.setDistance(1.0f);
    clipPlanes[0].rotate(Quaternion_fromAngleAxis(static_cast<float>(M_PI / 2.0), Vector3_new(0.0f, 1.0f, 0.0f)));
    clipPlanes[1].rotate(Quaternion_fromAngleAxis(static_cast<float>(M_PI / 2.0), Vector3_new(-1.0f, 0.0f, 0.0f)));
    currentClipPlane = NULL;
    for (int i = 0; i < 3; ++i) clipPlanes[i].setPlaneId(GL_CLIP_PLANE0 + i);
    resize(width, height);
    for (int i = 0; i < n; ++i) {
        const int *e = events + 5 * i;
        g_modifiers = e[4];
        if (e[0] == 0) mouseInteract(e[1], GLUT_DOWN, e[2], e[3]);
        else if (e[0] == 1) mouseMotionInteract(e[2], e[3]);
        else currentClipPlane = (e[1] >= 0 && e[1] < 3) ? &clipPlanes[e[1]] : NULL;
    }
    g_modifiers = 0;
    Transform *obj[5] = {&cam, &light, &clipPlanes[0], &clipPlanes[1], &clipPlanes[2]};
    for (int k = 0; k < 5; ++k) {
        float *o = out + 13 * k;
        Transform *t = obj[k];
        o[0] = t->_q_internal.x; o[1] = t->_q_internal.y; o[2] = t->_q_internal.z; o[3] = t->_q_internal.w;
        o[4] = t->_q.x; o[5] = t->_q.y; o[6] = t->_q.z; o[7] = t->_q.w;
        o[8] = t->_dist;
        o[9] = o[10] = o[11] = 0.0f;
        if (k == 0) { Vector3 p = cam.getPosition(); o[9] = p.x; o[10] = p.y; o[11] = p.z; }
        o[12] = t->_locked ? 1.0f : 0.0f;
    }
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 4; ++k) out_normals[4 * i + k] = clipPlanes[i].getNormal()[k];
    currentClipPlane = NULL;
    resize(64, 48);
    return 0;
}

int vvref_has_app(void) { return 1; }

} /* extern "C" */
