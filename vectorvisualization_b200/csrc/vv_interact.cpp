// vv_interact.cpp -- the reference's mouse interaction (mouseInteract / mouseMotionInteract of VV/3DLIC.cpp:490-600) as plain
// host logic on a caller-owned VVInteractState: trackball rotation of the camera, the light or the selected clip plane
// (VV/trackball.cpp:26-74, VV/transform.cpp:70-234, 330-372), camera translate / dolly (VV/camera.cpp:71-81), light and clip-plane
// distance.  Every expression keeps the reference's types and order of operations (float vs double, libm overloads), so the
// resulting quaternions, positions and plane equations are the same bits (tests/test_host_vs_ref.py).  vv_apply_interaction
// hands the state to a renderer through the public entry points.
#include <math.h>

#include <cstring>

#include "vv_host.h"

using namespace vvb200;

namespace {

const float kEps = 1e-5f;                 // EPS, VV/mmath.h:43
const float kTrackballSize = 0.8f;        // TRACKBALLSIZE, VV/trackball.h:29
const float kMouseScale = 1.7f;           // MOUSE_SCALE, VV/camera.h:9
const float kCosVar = (float)cos(15.0 * M_PI / 180.0);   // Transform::cos_var, LOCK_VARIANCE 15 (VV/transform.cpp:33, transform.h:38)
const float kSqrt13 = 1.0f / sqrt(3.0f);                 // Transform::sqrt1_3 (VV/transform.cpp:34)

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };

#define SQR(x) ((x) * (x))

V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
V3 v3_cross(V3 u, V3 v) { return v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
float v3_dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
V3 v3_smult(float s, V3 v) { return v3(s * v.x, s * v.y, s * v.z); }
V3 v3_add(V3 u, V3 v) { return v3(u.x + v.x, u.y + v.y, u.z + v.z); }
V3 v3_sub(V3 u, V3 v) { return v3(u.x - v.x, u.y - v.y, u.z - v.z); }

// Vector3_normalize, VV/mmath.cpp
float v3_normalize(V3 *v)
{
    float d = (float)sqrt(SQR(v->x) + SQR(v->y) + SQR(v->z));
    if (d > kEps) { v->x /= d; v->y /= d; v->z /= d; return d; }
    v->x = v->y = v->z = 0.f;
    return 0.f;
}

// Quaternion_fromAngleAxis, VV/mmath.cpp
Q4 q_from_angle_axis(float angle, V3 axis)
{
    Q4 q;
    float l = v3_normalize(&axis);
    if (l > kEps) {
        l = (float)sin(0.5f * angle);
        q.x = axis.x * l; q.y = axis.y * l; q.z = axis.z * l;
        q.w = (float)cos(0.5f * angle);
    } else {
        q.x = 0.f; q.y = 0.f; q.z = 0.f; q.w = 1.f;
    }
    return q;
}

// Quaternion_mult, VV/mmath.cpp
Q4 q_mult(Q4 p, Q4 q)
{
    Q4 r;
    r.w = p.w * q.w - (p.x * q.x + p.y * q.y + p.z * q.z);
    r.x = p.w * q.x + q.w * p.x + p.y * q.z - p.z * q.y;
    r.y = p.w * q.y + q.w * p.y + p.z * q.x - p.x * q.z;
    r.z = p.w * q.z + q.w * p.z + p.x * q.y - p.y * q.x;
    return r;
}

// Quaternion_normalize, VV/mmath.cpp
void q_normalize(Q4 *q)
{
    float d = (float)sqrt(SQR(q->w) + SQR(q->x) + SQR(q->y) + SQR(q->z));
    if (d > kEps) {
        d = 1.f / d;
        q->w *= d; q->x *= d; q->y *= d; q->z *= d;
    } else {
        q->w = 1.f; q->x = q->y = q->z = 0.f;
    }
}

// Quaternion_multVector3, VV/mmath.cpp
V3 q_mult_v3(Q4 q, V3 v)
{
    V3 u = v3(q.x, q.y, q.z);
    float uu = v3_dot(u, u), uv = v3_dot(u, v);
    V3 r = v3_smult(2.f, v3_add(v3_smult(uv, u), v3_smult(q.w, v3_cross(u, v))));
    return v3_add(r, v3_smult(SQR(q.w) - uu, v));
}

// projectToSphere / trackBall, VV/trackball.cpp:26-74
float project_to_sphere(float radius, float x, float y)
{
    float dist, t, z;
    dist = sqrt(x * x + y * y);
    if (dist < radius * M_SQRT1_2) {
        z = sqrt(radius * radius - dist * dist);
    } else {
        t = radius / (float)M_SQRT2;
        z = t * t / dist;
    }
    return z;
}

Q4 track_ball(float pX, float pY, float qX, float qY)
{
    if ((fabs(pX - qX) < kEps) && (fabs(pY - qY) < kEps)) { Q4 id; id.x = id.y = id.z = 0.0f; id.w = 1.0f; return id; }
    V3 v1 = v3(pX, pY, project_to_sphere(kTrackballSize, pX, pY));
    V3 v2 = v3(qX, qY, project_to_sphere(kTrackballSize, qX, qY));
    V3 cross = v3_cross(v1, v2);
    V3 t = v3_sub(v1, v2);
    float rot = sqrt(t.x * t.x + t.y * t.y + t.z * t.z) / (2.0f * kTrackballSize);
    rot = (rot > 1.0f) ? 1.0f : ((rot < -1.0f) ? -1.0f : rot);
    float phi = 2.0f * asin(rot);
    return q_from_angle_axis(phi, cross);
}

Q4 load_q(const float *p) { Q4 q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3]; return q; }
void store_q(float *p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

// Transform::update, VV/transform.cpp:155-234: _q from _q_internal and the lock state
void transform_update(VVTransformState *t)
{
    Q4 qi = load_q(t->q_internal);
    if (t->locked) {
        Q4 q_conj; q_conj.x = -qi.x; q_conj.y = -qi.y; q_conj.z = -qi.z; q_conj.w = qi.w;
        Q4 p; p.x = 0.0f; p.y = 0.0f; p.z = -1.0f; p.w = 0.0f;
        p = q_mult(q_mult(qi, p), q_conj);
        V3 v_vec = v3(p.x, p.y, p.z), v_lock = v_vec;
        V3 sign = v3((v_lock.x < 0.0) ? -1.0f : 1.0f, (v_lock.y < 0.0) ? -1.0f : 1.0f, (v_lock.z < 0.0) ? -1.0f : 1.0f);
        v_lock.x = fabs(v_lock.x); v_lock.y = fabs(v_lock.y); v_lock.z = fabs(v_lock.z);
        if (v_lock.x > kCosVar) { v_lock.x = 1.0f; v_lock.y = v_lock.z = 0.0f; }
        else if (v_lock.y > kCosVar) { v_lock.y = 1.0f; v_lock.x = v_lock.z = 0.0f; }
        else if (v_lock.z > kCosVar) { v_lock.z = 1.0f; v_lock.x = v_lock.y = 0.0f; }
        else if ((v_lock.x + v_lock.y) * M_SQRT1_2 > kCosVar) { v_lock.x = v_lock.y = (float)M_SQRT1_2; v_lock.z = 0.0f; }
        else if ((v_lock.y + v_lock.z) * M_SQRT1_2 > kCosVar) { v_lock.y = v_lock.z = (float)M_SQRT1_2; v_lock.x = 0.0f; }
        else if ((v_lock.x + v_lock.z) * M_SQRT1_2 > kCosVar) { v_lock.x = v_lock.z = (float)M_SQRT1_2; v_lock.y = 0.0f; }
        else if ((v_lock.x + v_lock.y + v_lock.z) * kSqrt13 > kCosVar) { v_lock.x = v_lock.y = v_lock.z = kSqrt13; }
        v_lock = v3(v_lock.x * sign.x, v_lock.y * sign.y, v_lock.z * sign.z);
        V3 axis = v3_cross(v_vec, v_lock);
        float phi = acos(v3_dot(v_vec, v_lock));
        store_q(t->q, q_mult(q_from_angle_axis(phi, axis), qi));
    } else {
        store_q(t->q, qi);
    }
}

// Transform::rotate(Quaternion), VV/transform.cpp:70-80
void transform_rotate(VVTransformState *t, Q4 q_rot)
{
    Q4 qi = q_mult(q_rot, load_q(t->q_internal));
    q_normalize(&qi);
    store_q(t->q_internal, qi);
    transform_update(t);
}

// Transform::rotate(Quaternion q_rot, Quaternion q_cam), VV/transform.cpp:83-104: rotation in camera space
void transform_rotate_cam(VVTransformState *t, Q4 q_rot, Q4 q_cam)
{
    Q4 cam_conj; cam_conj.x = -q_cam.x; cam_conj.y = -q_cam.y; cam_conj.z = -q_cam.z; cam_conj.w = q_cam.w;
    Q4 r = q_mult(q_cam, load_q(t->q_internal));
    Q4 qi = q_mult(cam_conj, q_mult(q_rot, r));
    q_normalize(&qi);
    store_q(t->q_internal, qi);
    transform_update(t);
}

Q4 mouse_trackball(const VVInteractState *s, int xNew, int yNew, int xOld, int yOld)
{
    const int _w = s->w, _h = s->h;
    return track_ball((2.0f * xOld - _w) / _w, (_h - 2.0f * yOld) / _h, (2.0f * xNew - _w) / _w, (_h - 2.0f * yNew) / _h);
}

// Transform::setLock, VV/transform.cpp:127-140
void transform_set_lock(VVTransformState *t, int enable)
{
    if (!enable && t->locked) std::memcpy(t->q_internal, t->q, sizeof(t->q));
    t->locked = enable != 0;
}

// the normal of a clip plane follows its orientation: _normal.xyz = _q * _initialNormal (0, 0, -1), VV/transform.cpp:330-372
void clip_update_normal(VVInteractState *s, int i)
{
    V3 v = q_mult_v3(load_q(s->clip[i].q), v3(0.0f, 0.0f, -1.0f));
    s->clip_normal[i][0] = v.x; s->clip_normal[i][1] = v.y; s->clip_normal[i][2] = v.z;
}

void transform_init(VVTransformState *t, float dist)
{
    std::memset(t, 0, sizeof(*t));
    t->q[3] = 1.0f; t->q_internal[3] = 1.0f;
    t->dist = dist;
}

}  // namespace

extern "C" {

void vv_interact_init(VVInteractState *s, int width, int height)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    transform_init(&s->cam, 4.0f);                     // Camera ctor, VV/camera.cpp:42-47
    transform_init(&s->light, 1.0f);                   // light.setDistance(1.0f), VV/3DLIC.cpp:681
    for (int i = 0; i < 3; ++i) {
        transform_init(&s->clip[i], 0.0f);
        s->clip_normal[i][0] = 0.0; s->clip_normal[i][1] = 0.0; s->clip_normal[i][2] = -1.0; s->clip_normal[i][3] = 0.0;   // ClipPlane ctor
    }
    s->w = width > 0 ? width : 1; s->h = height > 0 ? height : 1;
    // VV/3DLIC.cpp:763-764
    transform_rotate(&s->clip[0], q_from_angle_axis(static_cast<float>(M_PI / 2.0), v3(0.0f, 1.0f, 0.0f)));
    clip_update_normal(s, 0);
    transform_rotate(&s->clip[1], q_from_angle_axis(static_cast<float>(M_PI / 2.0), v3(-1.0f, 0.0f, 0.0f)));
    clip_update_normal(s, 1);
    s->mouse_mode = VV_MOUSE_ROTATE;
}

void vv_interact_resize(VVInteractState *s, int width, int height)
{
    if (!s) return;
    s->w = width < 1 ? 1 : width; s->h = height < 1 ? 1 : height;       // resize(), VV/3DLIC.cpp:174-200
}

// mouseInteract, VV/3DLIC.cpp:490-547.  modifiers: VV_MOD_SHIFT locks the rotation to the nearest axis / diagonal, VV_MOD_CTRL
// addresses the selected clip plane (selected_clip >= 0) or else the light
void vv_mouse(VVInteractState *s, int selected_clip, int button, int x, int y, int modifiers)
{
    if (!s) return;
    const int locked = (modifiers & VV_MOD_SHIFT) ? 1 : 0;
    s->old_x = x; s->old_y = y;
    switch (button) {
    case VV_BUTTON_LEFT:
        if (modifiers & VV_MOD_CTRL) {
            if (selected_clip >= 0 && selected_clip < 3) { transform_set_lock(&s->clip[selected_clip], locked); s->mouse_mode = VV_MOUSE_ROTATE_CLIP; }
            else { transform_set_lock(&s->light, locked); s->mouse_mode = VV_MOUSE_ROTATE_LIGHT; }
        } else {
            transform_set_lock(&s->cam, locked);
            s->mouse_mode = VV_MOUSE_ROTATE;
        }
        break;
    case VV_BUTTON_MIDDLE: s->mouse_mode = VV_MOUSE_TRANSLATE; break;
    case VV_BUTTON_RIGHT:
        if (modifiers & VV_MOD_CTRL) s->mouse_mode = (selected_clip >= 0 && selected_clip < 3) ? VV_MOUSE_TRANSLATE_CLIP : VV_MOUSE_TRANSLATE_LIGHT;
        else s->mouse_mode = VV_MOUSE_DOLLY;
        break;
    default: break;
    }
}

// mouseMotionInteract, VV/3DLIC.cpp:550-600
void vv_motion(VVInteractState *s, int selected_clip, int x, int y)
{
    if (!s) return;
    const Q4 q_cam = load_q(s->cam.q);
    const int ox = s->old_x, oy = s->old_y;
    const bool have_clip = selected_clip >= 0 && selected_clip < 3;
    switch (s->mouse_mode) {
    case VV_MOUSE_ROTATE: transform_rotate(&s->cam, mouse_trackball(s, x, y, ox, oy)); break;
    case VV_MOUSE_TRANSLATE:                                        // Camera::translate, VV/camera.cpp:71-75
        s->cam.pos[0] += kMouseScale * (x - ox) / (float)s->w;
        s->cam.pos[1] += kMouseScale * (oy - y) / (float)s->h;
        break;
    case VV_MOUSE_DOLLY:                                            // Camera::dolly, VV/camera.cpp:78-81
        s->cam.pos[2] -= 2.0f * kMouseScale * (oy - y) / (float)s->h;
        break;
    case VV_MOUSE_ROTATE_LIGHT: transform_rotate(&s->light, mouse_trackball(s, x, y, ox, oy)); break;
    case VV_MOUSE_TRANSLATE_LIGHT:
        s->light.dist += 0.5f * kMouseScale * (y - oy) / static_cast<float>(s->w);
        if (s->light.dist < 1.0e-6f) s->light.dist = 1.0e-6f;
        break;
    case VV_MOUSE_ROTATE_CLIP:
        if (have_clip) {
            transform_rotate_cam(&s->clip[selected_clip], mouse_trackball(s, x, y, ox, oy), q_cam);
            clip_update_normal(s, selected_clip);
        }
        break;
    case VV_MOUSE_TRANSLATE_CLIP:
        if (have_clip) s->clip_normal[selected_clip][3] += kMouseScale * (oy - y) / static_cast<float>(s->w);    // ClipPlane::operator+=(double)
        break;
    default: break;
    }
    s->old_x = x; s->old_y = y;
}

int vv_apply_interaction(VVRenderer *r, const VVInteractState *s, const VVAppState *app)
{
    if (!r || !s) return fail(VV_ERR_INVALID, "vv_apply_interaction: null argument");
    int rc = vv_set_window(r, s->w, s->h);
    if (rc == VV_OK) rc = vv_set_camera(r, s->cam.q, s->cam.pos, s->cam.dist, 35.0f, 0.1f, 50.0f);   // VV/camera.cpp:42-47
    if (rc == VV_OK) rc = vv_set_light(r, s->light.q, s->light.dist);
    if (rc == VV_OK) rc = vv_update_light_pos(r);                                                  // mouseMotionInteract ends with it
    for (int i = 0; i < 3 && rc == VV_OK; ++i) rc = vv_set_clip_plane(r, i, s->clip_normal[i], app ? app->clip_active[i] : 0);
    return rc;
}

}  // extern "C"
