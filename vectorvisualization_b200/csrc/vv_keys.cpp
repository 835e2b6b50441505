// vv_keys.cpp -- the reference's key map (keyboard / keyboardSpecial of VV/3DLIC.cpp:243-488) as plain host logic on a
// caller-owned VVAppState (the counterpart of the application globals licParams, renderTechnique, animationMode,
// updateSceneCont, currentClipPlane of VV/3DLIC.h:29-55), and its application to a renderer handle.  No GLUT: a caller's
// event loop hands in the key; nothing here touches the device except through the public entry points.
#include <cstring>

#include "vv_host.h"

using namespace vvb200;

namespace {

void set_defines(VVAppState *s, const char *d)
{
    std::strncpy(s->defines, d, sizeof(s->defines) - 1);
    s->defines[sizeof(s->defines) - 1] = 0;
}

// switchClipPlane, VV/transform.cpp:461-483: selecting the selected plane deactivates and deselects it; anything else
// selects (and activates) the new plane; NULL only deselects
void switch_clip_plane(VVAppState *s, int plane)
{
    if (s->selected_clip >= 0 && s->selected_clip == plane) {
        s->clip_active[plane] = 0;
        s->selected_clip = -1;
        return;
    }
    s->selected_clip = plane;
    if (plane >= 0) s->clip_active[plane] = 1;
}

}  // namespace

extern "C" {

void vv_app_state_init(VVAppState *s)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    vv_default_lic_params(&s->lic);          // LICParams ctor, VV/types.h:91-109
    s->technique = VV_VOLIC_VOLUME;          // VV/3DLIC.h:51
    s->selected_clip = -1;
}

int vv_key_apply(VVAppState *s, int key, int special)
{
    if (!s) return 0;
    int act = 0;
    VVLicParams &p = s->lic;
    if (special) {
        // keyboardSpecial, VV/3DLIC.cpp:457-488: every special key ends in setTechnique + updateScene
        switch (key) {
        case 1: s->technique = VV_VOLIC_VOLUME; break;
        case 2: s->technique = VV_VOLIC_RAYCAST; break;
        case 3: s->technique = VV_VOLIC_SLICING; act |= VV_KEY_UPDATE_SLICES; break;
        case 4: s->technique = VV_VOLIC_LICVOLUME; act |= VV_KEY_UPDATE_LICVOLUME; break;
        case 5: s->animation = !s->animation; break;
        default: break;
        }
        return act | VV_KEY_SET_TECHNIQUE | VV_KEY_UPDATE_SCENE;
    }
    bool update = false;
    switch (key) {
    case 'q': case 27: return VV_KEY_QUIT;                                  // exit(1) there; the caller decides here
    case '0': s->screenshot = 1; act |= VV_KEY_SCREENSHOT; update = true; break;
    case 'R': s->animation = 1; s->recording = !s->recording; act |= VV_KEY_SWITCH_RECORDING; update = true; break;
    case 'r': set_defines(s, ""); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case 'F': s->float_target = !s->float_target; update = true; break;
    case 'p': update = true; break;                                         // tfEdit.updateTextures(): the TF is pushed by the caller
    case 'L': s->lowres = !s->lowres; update = true; break;
    case 'I': update = true; break;                                         // idle redrawing on / off: the caller's loop
    case '[': p.stepSizeVol /= 2.0f; if (p.stepSizeVol < 0.0) p.stepSizeVol = 0.0f; update = true; break;       // MIN_STEPSIZE, VV/types.h:58
    case ']': p.stepSizeVol *= 2.0f; if (p.stepSizeVol > 1.0) p.stepSizeVol = 1.0f; update = true; break;       // MAX_STEPSIZE, VV/types.h:57
    case 's': ++p.stepsForward; update = true; break;
    case 'x': --p.stepsForward; if (p.stepsForward < 1) p.stepsForward = 1; update = true; break;
    case 'S': ++p.stepsBackward; update = true; break;
    case 'X': --p.stepsBackward; if (p.stepsBackward < 1) p.stepsBackward = 1; update = true; break;
    case 'a': p.stepSizeLIC *= 2.0f; update = true; break;
    case 'z': p.stepSizeLIC *= 0.5f; if (p.stepSizeLIC < 0.0005f) p.stepSizeLIC = 0.0005f; update = true; break;
    case 'h': p.freqScale += 0.2f; update = true; break;
    case 'n': p.freqScale -= 0.2f; if (p.freqScale < 0.5f) p.freqScale = 0.5f; update = true; break;
    case 'j': p.illumScale += 0.05f; update = true; break;
    case 'm': p.illumScale -= 0.05f; if (p.illumScale < 0.05f) p.illumScale = 0.05f; update = true; break;
    case 'g': p.gradientScale += 0.2f; update = true; break;
    case 'b': p.gradientScale -= 0.2f; if (p.gradientScale < 0.2f) p.gradientScale = 0.2f; update = true; break;
    case 'u': act |= VV_KEY_UPDATE_LICVOLUME; update = true; break;
    case '1': case '2': case '3': switch_clip_plane(s, key - '1'); update = true; break;
    case '4': switch_clip_plane(s, -1); update = true; break;
    case '7': set_defines(s, "#define ILLUM_ZOECKLER"); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case '8': set_defines(s, "#define ILLUM_MALLO"); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case '9': set_defines(s, "#define ILLUM_GRADIENT"); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case '6': set_defines(s, "#define SPEED_OF_FLOW"); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case '.': set_defines(s, "#define VOLUME_ANIMATION"); act |= VV_KEY_RELOAD_SHADER; update = true; break;
    case ' ': s->continuous = !s->continuous; update = true; break;         // enableFrameStore(!updateSceneCont)
    default: break;                                                         // H, w, t, 5: HUD / wireframe / TF editor / light gizmo
    }
    // VV/3DLIC.cpp:450-455.  updateScene is a global there that stays set until the next frame; a caller that batches keys
    // between frames gets the same effect by OR-ing the returned actions.
    if (update) {
        act |= VV_KEY_UPDATE_SCENE;
        if (s->technique == VV_VOLIC_LICVOLUME) act |= VV_KEY_UPDATE_LICVOLUME;
        if (s->technique == VV_VOLIC_SLICING) act |= VV_KEY_UPDATE_SLICES;
    }
    return act;
}

// ---- animation bookkeeping: VectorDataSet::interpIndex / checkInterpolateStage + DatFile::getNextTimeStep ----
void vv_time_cursor_init(VVTimeCursor *c, int time_begin, int time_end, int interp_size)
{
    if (!c) return;
    c->time_begin = time_begin; c->time_end = time_end; c->current = time_begin;   // DatFile: _timestep = _timeStepBeg
    c->interp_index = 0; c->interp_size = interp_size;                              // VV/dataset.cpp:86, VV/3DLIC.cpp:705
}

int vv_time_cursor_next(const VVTimeCursor *c)
{
    return (c->current == c->time_end) ? c->time_begin : c->current + 1;            // DatFile::NextTimeStep, VV/reader.cpp:333-336
}

int vv_time_cursor_tick(VVTimeCursor *c, int *advanced)
{
    const int used = c->interp_index;          // createTextureIterp packs with interpIndex / InterpSize ...
    ++c->interp_index;                         // ... and increments (VV/dataset.cpp:633)
    int adv = 0;
    if (c->interp_index >= c->interp_size) {   // checkInterpolateStage, VV/dataset.cpp:202-210
        c->current = vv_time_cursor_next(c);   // getNextTimeStep() moves on; newData = NextTimeStep() of the new position
        c->interp_index = 0;
        adv = 1;
    }
    if (advanced) *advanced = adv;
    return used;
}

int vv_keyboard(VVRenderer *r, VVAppState *s, int key, int special)
{
    if (!r || !s) { fail(VV_ERR_INVALID, "vv_keyboard: null argument"); return -1; }
    const int act = vv_key_apply(s, key, special);
    if (act == 0 || act == VV_KEY_QUIT) return act;
    int rc = VV_OK;
    if (act & VV_KEY_RELOAD_SHADER) rc = vv_load_glsl_shader(r, s->defines[0] ? s->defines : nullptr);
    if (rc == VV_OK) rc = vv_set_lic_params(r, &s->lic);
    if (rc == VV_OK) rc = vv_enable_lowres(r, s->lowres);
    if (rc == VV_OK) rc = vv_enable_float_target(r, s->float_target);
    for (int i = 0; i < 3 && rc == VV_OK; ++i) rc = vv_set_clip_plane(r, i, nullptr, s->clip_active[i]);
    // F1 (raw vector-field DVR) is outside this library: the state records it, the handle keeps its technique
    if (rc == VV_OK && (act & VV_KEY_SET_TECHNIQUE) && s->technique != VV_VOLIC_VOLUME) rc = vv_set_technique(r, s->technique);
    if (rc == VV_OK && (act & VV_KEY_SCREENSHOT)) rc = vv_screenshot(r);
    if (rc == VV_OK && (act & VV_KEY_SWITCH_RECORDING)) rc = vv_switch_recording(r) < 0 ? VV_ERR_INVALID : VV_OK;
    if (rc == VV_OK && (act & VV_KEY_UPDATE_LICVOLUME)) rc = vv_update_lic_volume(r);
    if (rc == VV_OK && (act & VV_KEY_UPDATE_SLICES)) rc = vv_update_slices(r);
    return rc == VV_OK ? act : -1;
}

}  // extern "C"
