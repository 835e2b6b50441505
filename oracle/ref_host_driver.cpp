/* ref_host_driver.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's OWN host code (VV/reader.cpp, parseArg.cpp, gradient.cpp, dataset.cpp, transferEdit.cpp,
 * mmath.cpp, texture.cpp -- compiled unmodified from /root/reference by oracle/build_ref.py against the capturing
 * GL stub in oracle/ref_shim/) and hands back what it would have uploaded to OpenGL:
 *   the packed RGBA16F vector texture (VectorDataSet::createTextureIterp), the noise texture with or without
 *   gradients (NoiseDataSet::createTexture), the LIC filter kernel + inverse area (LICFilter), the two transfer
 *   function textures (TransferEdit::updateTextures), DatFile / ParseArguments results and the mmath quaternion helpers.
 * Only the PNG codec (the reference links libpng, which is not in this image) and the GL entry points are ours.
 */
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "GL/glew.h"
#include "dataset.h"
#include "gradient.h"
#include "illumination.h"
#include "imageUtils.h"
#include "mmath.h"
#include "parseArg.h"
#include "reader.h"
#include "transferEdit.h"
/* ViewSlicing / Renderer keep their state private; the driver only READS it (and calls the protected
 * Renderer::setRenderVolParams).  The standard headers those files pull in are included first, untouched. */
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#define private public
#define protected public
#include "slicing.h"
#include "renderer.h"
#undef private
#undef protected
#include <cmath>

/* ------------------------------------------------------------------------------------------------ GL capture */
/* never destroyed: the reference's global objects (VV/3DLIC.h) call glDeleteTextures from their static destructors */
static std::map<GLuint, VVStubTex> &g_tex = *new std::map<GLuint, VVStubTex>();
static GLuint g_next_id = 1, g_bound = 0, g_last = 0;
static std::map<GLuint, VVStubTex> &g_dead = *new std::map<GLuint, VVStubTex>();   /* records of deleted textures (no data) */
static std::vector<GLuint> &g_upload_log = *new std::vector<GLuint>();   /* texture ids in glTexImage* order */

static VVStubTex &cur()
{
    VVStubTex &t = g_tex[g_bound];
    return t;
}

extern "C" {
VVStubTex *vv_stub_texture(GLuint id) { auto it = g_tex.find(id); return it == g_tex.end() ? nullptr : &it->second; }
GLuint vv_stub_last_texture(void) { return g_last; }
void vv_stub_reset(void)
{
    for (auto &kv : g_tex) std::free(kv.second.data);
    g_tex.clear();
    g_bound = g_last = 0;
}
void glGenTextures(GLsizei n, GLuint *ids) { for (int i = 0; i < n; ++i) { ids[i] = g_next_id++; g_tex[ids[i]] = VVStubTex(); } }
void glDeleteTextures(GLsizei n, const GLuint *ids)
{
    for (int i = 0; i < n; ++i) {
        auto it = g_tex.find(ids[i]);
        if (it == g_tex.end()) continue;
        VVStubTex dead = it->second;                /* keep format / sampler state of deleted textures for vvref_texture_state */
        dead.data = nullptr; dead.bytes = 0;
        g_dead[ids[i]] = dead;
        std::free(it->second.data);
        g_tex.erase(it);
    }
}
void glBindTexture(GLenum target, GLuint id) { g_bound = id; g_tex[id].target = target; }
void glTexParameteri(GLenum, GLenum pname, GLint v)
{
    VVStubTex &t = cur();
    switch (pname) {
    case GL_TEXTURE_MIN_FILTER: t.min_filter = v; break;
    case GL_TEXTURE_MAG_FILTER: t.mag_filter = v; break;
    case GL_TEXTURE_WRAP_S: t.wrap_s = v; break;
    case GL_TEXTURE_WRAP_T: t.wrap_t = v; break;
    case GL_TEXTURE_WRAP_R: t.wrap_r = v; break;
    default: break;
    }
}
static void upload(GLint ifmt, int w, int h, int d, GLenum fmt, GLenum type, const void *data)
{
    VVStubTex &t = cur();
    int ch = (fmt == GL_RGBA) ? 4 : (fmt == GL_RGB ? 3 : (fmt == GL_LUMINANCE_ALPHA ? 2 : 1));
    int bs = (type == GL_FLOAT) ? 4 : (type == GL_UNSIGNED_SHORT ? 2 : 1);
    t.internal_format = ifmt; t.format = fmt; t.type = type;
    t.dim[0] = w; t.dim[1] = h; t.dim[2] = d;
    t.bytes = (size_t)w * h * d * ch * bs;
    std::free(t.data);
    t.data = std::malloc(t.bytes ? t.bytes : 1);
    if (data) std::memcpy(t.data, data, t.bytes);
    g_last = g_bound;
    g_upload_log.push_back(g_bound);
}
void glTexImage1D(GLenum, GLint, GLint ifmt, GLsizei w, GLint, GLenum fmt, GLenum type, const void *data) { upload(ifmt, w, 1, 1, fmt, type, data); }
void glTexImage2D(GLenum, GLint, GLint ifmt, GLsizei w, GLsizei h, GLint, GLenum fmt, GLenum type, const void *data) { upload(ifmt, w, h, 1, fmt, type, data); }
void glTexImage3D(GLenum, GLint, GLint ifmt, GLsizei w, GLsizei h, GLsizei d, GLint, GLenum fmt, GLenum type, const void *data) { upload(ifmt, w, h, d, fmt, type, data); }
}

/* ------------------------------------------------------------------------------------------------ PNG codec (ours) */
static unsigned be32(const unsigned char *p) { return (p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

bool pngRead(const char *fileName, Image *img)
{
    FILE *fp = std::fopen(fileName, "rb");
    if (!fp) { std::fprintf(stderr, "Could not open PNG file %s.\n", fileName); return false; }
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(fp);
    if (buf.size() < 8) return false;
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = -1;
    std::vector<unsigned char> idat;
    while (pos + 12 <= buf.size()) {
        unsigned len = be32(&buf[pos]);
        const char *type = (const char *)&buf[pos + 4];
        const unsigned char *d = &buf[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) { w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; }
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (depth != 8) return false;
    int ch = ctype == 0 ? 1 : ctype == 4 ? 2 : ctype == 2 ? 3 : ctype == 6 ? 4 : 0;
    if (!ch) return false;
    size_t stride = (size_t)w * ch;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf rl = raw.size();
    if (uncompress(raw.data(), &rl, idat.data(), idat.size()) != Z_OK) return false;
    img->imgData = new unsigned char[stride * h];
    for (int y = 0; y < h; ++y) {
        const unsigned char *src = &raw[(stride + 1) * y];
        unsigned char *dst = img->imgData + stride * y, *up = y ? img->imgData + stride * (y - 1) : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= (size_t)ch ? dst[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0, x = src[1 + i];
            switch (src[0]) {
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: break;
            }
            dst[i] = (unsigned char)x;
        }
    }
    img->width = w; img->height = h; img->channel = ch; img->gamma = 1.0;
    return true;
}
bool pngWrite(const char *, const Image *, bool) { return false; }
bool ppmRead(const char *, Image *) { return false; }
bool ppmWrite(const char *, const Image *) { return false; }

/* ------------------------------------------------------------------------------------------------ exported drivers */
static int copy_tex(GLuint id, void *out, size_t cap, int dims[3], int *ifmt, int *wrap)
{
    VVStubTex *t = vv_stub_texture(id);
    if (!t || !t->data) return -1;
    if (out) { if (cap < t->bytes) return -2; std::memcpy(out, t->data, t->bytes); }
    if (dims) { dims[0] = t->dim[0]; dims[1] = t->dim[1]; dims[2] = t->dim[2]; }
    if (ifmt) *ifmt = t->internal_format;
    if (wrap) *wrap = t->wrap_s;
    return (int)(t->bytes > 0x7fffffff ? 0x7fffffff : t->bytes);
}

extern "C" {

/* VectorDataSet: loadData(.dat) + loadTimeStep x2 + setInterpolateSize + createTextureIterp (VV/3DLIC.cpp:686-707),
 * repeated `interp_index + 1` times so the captured upload is the one made with that interpIndex */
int vvref_vector_texture(const char *dat, int interp_index, int interp_size, float *out, size_t cap, int dims[3], float geom[17],
                         int *ifmt, int *wrap)
{
    VectorDataSet vd;
    if (!vd.loadData(dat)) return -10;
    vd.getVolumeData()->data = vd.loadTimeStep(vd.getCurTimeStep());
    vd.getVolumeData()->newData = vd.loadTimeStep(vd.NextTimeStep());
    vd.setInterpolateSize(interp_size);
    for (int i = 0; i <= interp_index; ++i) vd.createTextureIterp("VectorData_Tex", GL_TEXTURE2_ARB, true);
    VolumeData *v = vd.getVolumeData();
    if (geom) {
        for (int i = 0; i < 3; ++i) { geom[i] = v->extent[i]; geom[3 + i] = v->center[i]; geom[14 + i] = v->sliceDist[i]; }
        for (int i = 0; i < 4; ++i) { geom[6 + i] = v->scale[i]; geom[10 + i] = v->scaleInv[i]; }
    }
    return copy_tex(vd.getTextureRef()->id, out, cap, dims, ifmt, wrap);
}

/* NoiseDataSet: loadData(file) + enableGradient + createTexture (VV/3DLIC.cpp:711-713) */
int vvref_noise_texture(const char *file, int use_gradient, unsigned char *out, size_t cap, int dims[3], int *ifmt, int *wrap)
{
    NoiseDataSet nd;
    if (!nd.loadData(file)) return -10;
    if (!nd.getFileName()) return -11;          /* fell back to rand() white noise: not reproducible */
    nd.enableGradient(use_gradient != 0);
    nd.createTexture("Noise_Tex", GL_TEXTURE3_ARB);
    /* the gradient cache file <noise>.grd is written next to the input: remove it so runs stay independent */
    std::string grd = std::string(file) + ".grd";
    std::remove(grd.c_str());
    return copy_tex(nd.getTextureRef()->id, out, cap, dims, ifmt, wrap);
}

/* the same with the gradient cache left alone: an existing <noise>.grd is used (loadGradients, VV/gradient.cpp:112-149),
 * otherwise the reference writes one (saveGradients, :152-187) and it stays on disk */
int vvref_noise_texture_cached(const char *file, unsigned char *out, size_t cap, int dims[3], int *ifmt, int *wrap)
{
    NoiseDataSet nd;
    if (!nd.loadData(file)) return -10;
    if (!nd.getFileName()) return -11;
    nd.enableGradient(true);
    nd.createTexture("Noise_Tex", GL_TEXTURE3_ARB);
    return copy_tex(nd.getTextureRef()->id, out, cap, dims, ifmt, wrap);
}

/* Sampler state of what the reference uploads: the ids of all glTexImage* calls since the last reset, in call order, and per id
 * (target, internal format, format, type, min filter, mag filter, wrap s, wrap t, wrap r, width, height, depth) as it stands now
 * (the reference sets the parameters after the upload). */
void vvref_upload_log_reset(void) { g_upload_log.clear(); g_dead.clear(); }
int vvref_upload_log(unsigned int *ids, int cap)
{
    int n = (int)g_upload_log.size();
    for (int i = 0; i < n && i < cap; ++i) ids[i] = g_upload_log[i];
    return n;
}
int vvref_texture_state(unsigned int id, int out[12])
{
    VVStubTex *t = vv_stub_texture(id);
    if (!t) { auto it = g_dead.find(id); if (it == g_dead.end()) return -1; t = &it->second; }
    out[0] = (int)t->target; out[1] = t->internal_format; out[2] = (int)t->format; out[3] = (int)t->type;
    out[4] = t->min_filter; out[5] = t->mag_filter; out[6] = t->wrap_s; out[7] = t->wrap_t; out[8] = t->wrap_r;
    out[9] = t->dim[0]; out[10] = t->dim[1]; out[11] = t->dim[2];
    return 0;
}
/* the textures Renderer itself creates: FBO colour targets (createFBO / updateFBO, VV/renderer.cpp:536-620), the MC-offset
 * texture (updateMCOffsetTex, :636-679) and the two layers of the LIC volume buffer (VolumeBuffer ctor, VV/VolumeBuffer.cpp:5-57) */
int vvref_renderer_textures(int width, int height, int lw, int lh, int ld)
{
    Renderer r;
    r.createFBO();
    r.resize(width, height);
    r.updateMCOffsetTex(width, height);
    VolumeBuffer vb(GL_RGBA16F_ARB, lw, lh, ld, 2);
    return 0;
}

/* VolumeDataSet: loadData(.dat) + createTexture (VV/3DLIC.cpp:716-722; the scalar volume that gates the noise and feeds the
 * scalar TF index).  out receives the bytes handed to glTexImage3D (UCHAR or FLOAT source), type_out its GL type. */
int vvref_scalar_texture(const char *dat, unsigned char *out, size_t cap, int dims[3], int *ifmt, int *wrap, int *type_out, int *filter_out)
{
    VolumeDataSet sd;
    if (!sd.loadData(dat)) return -10;
    sd.createTexture("Scalar_Tex", GL_TEXTURE4_ARB);
    VVStubTex *t = vv_stub_texture(sd.getTextureRef()->id);
    if (t && type_out) *type_out = (int)t->type;
    if (t && filter_out) *filter_out = (t->min_filter == GL_LINEAR && t->mag_filter == GL_LINEAR) ? 1 : 0;
    return copy_tex(sd.getTextureRef()->id, out, cap, dims, ifmt, wrap);
}

/* LICFilter: loadData(png) or createBoxFilter + createTexture (VV/3DLIC.cpp:725-731) */
int vvref_filter_texture(const char *png, unsigned char *out, size_t cap, int *width, float *inv_area, int *wrap)
{
    LICFilter f;
    if (!png || !f.loadData(png)) f.createBoxFilter();
    f.createTexture("LIC_kernel_Tex", GL_TEXTURE5_ARB);
    int dims[3];
    int rc = copy_tex(f.getTextureRef()->id, out, cap, dims, nullptr, wrap);
    if (width) *width = f.getFilterWidth();
    if (inv_area) *inv_area = f.getInverseFilterArea();
    return rc;
}

/* TransferEdit: ctor (+ loadTF(name)) + updateTextures (VV/3DLIC.cpp:739-747): RGBA8 256 + LUMINANCE_ALPHA8 256 */
int vvref_tf_textures(const char *name, unsigned char *rgba, unsigned char *la, int *loaded)
{
    TransferEdit te;
    int ok = 0;
    if (name) ok = te.loadTF(name) ? 1 : 0;
    if (loaded) *loaded = ok;
    te.updateTextures();
    if (copy_tex(te.getTextureRGB()->id, rgba, 1024, nullptr, nullptr, nullptr) < 0) return -1;
    if (copy_tex(te.getTextureAlphaOpac()->id, la, 512, nullptr, nullptr, nullptr) < 0) return -2;
    return 0;
}

/* gradient.cpp entry points on a caller-provided u8 volume */
int vvref_noise_gradients(const unsigned char *noise, const int dims[3], const float sd[3], float *grad_f, float *filtered_f, unsigned char *quant)
{
    VolumeData vd;
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    vd.data = new unsigned char[n];
    std::memcpy(vd.data, noise, n);
    vd.dataType = DATRAW_UCHAR;
    for (int i = 0; i < 3; ++i) { vd.size[i] = dims[i]; vd.sliceDist[i] = sd[i]; }
    float *g = computeGradients(&vd);
    if (grad_f) std::memcpy(grad_f, g, 3 * n * sizeof(float));
    filterGradients(&vd, g);
    if (filtered_f) std::memcpy(filtered_f, g, 3 * n * sizeof(float));
    unsigned char *q = (unsigned char *)quantizeGradients(&vd, g, DATRAW_UCHAR);
    if (quant) std::memcpy(quant, q, 3 * n);
    delete[] g;
    delete[] q;
    return 0;
}

/* DatFile::parseDatFile */
int vvref_parse_dat(const char *dat, int res[3], float dist[3], int *dtype, int *ddim, int *tb, int *te)
{
    DatFile d;
    if (!d.parseDatFile((char *)dat)) return -1;
    for (int i = 0; i < 3; ++i) { res[i] = d.getDataSizes()[i]; dist[i] = d.getDataDists()[i]; }
    *dtype = (int)d.getDataType(); *ddim = d.getDataDimension(); *tb = d.getTimeStepBegin(); *te = d.getTimeStepEnd();
    return 0;
}

/* ParseArguments::parse; strings copied into out[6][512]: vol, noise, tf, filter, redirect, halton; flags[2]: gradients, lambda2 */
int vvref_parse_args(int argc, char **argv, char *out, int *flags)
{
    ParseArguments pa;
    pa.setArguments(argc, argv);
    bool ok = pa.parse();
    const char *s[6] = {pa.getVolFileName(), pa.getNoiseFileName(), pa.getTfFileName(), pa.getLicFilterFileName(),
                        pa.getRedirectFileName(), pa.getHaltonFileName()};
    for (int i = 0; i < 6; ++i) { std::memset(out + 512 * i, 0, 512); if (s[i]) std::strncpy(out + 512 * i, s[i], 511); }
    flags[0] = pa.getGradientsFlag(); flags[1] = pa.getLambda2Flag();
    return ok ? 0 : -1;
}

/* mmath.cpp */
void vvref_quat_angle_axis(const float q[4], float *angle, float axis[3])
{
    Quaternion qq; qq.x = q[0]; qq.y = q[1]; qq.z = q[2]; qq.w = q[3];
    Vector3 a;
    Quaternion_getAngleAxis(qq, angle, &a);
    axis[0] = a.x; axis[1] = a.y; axis[2] = a.z;
}
void vvref_quat_mult_vec(const float q[4], const float v[3], float out[3])
{
    Quaternion qq; qq.x = q[0]; qq.y = q[1]; qq.z = q[2]; qq.w = q[3];
    Vector3 a = Vector3_new(v[0], v[1], v[2]);
    Vector3 r = Quaternion_multVector3(qq, a);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* Illumination::createIllumTextures(true, true, true) as VV/3DLIC.cpp:735 calls it; returns the three uploads:
 * zoeckler float LA [h][w][2], mallo diffuse / specular float RGBA [h][w][4], and their internal formats */
int vvref_illum_tables(float *zoeckler, float *mdiff, float *mspec, int dims[2], int ifmt[3], float *spec_exp)
{
    Illumination il;
    il.createIllumTextures(true, true, true);
    int d[3];
    if (copy_tex(il.getTexZoeckler()->id, zoeckler, (size_t)256 * 256 * 2 * 4, d, &ifmt[0], nullptr) < 0) return -1;
    if (copy_tex(il.getTexMalloDiffuse()->id, mdiff, (size_t)256 * 256 * 4 * 4, d, &ifmt[1], nullptr) < 0) return -2;
    if (copy_tex(il.getTexMalloSpecular()->id, mspec, (size_t)256 * 256 * 4 * 4, d, &ifmt[2], nullptr) < 0) return -3;
    dims[0] = d[0]; dims[1] = d[1];
    *spec_exp = il.getSpecularExp();
    return 0;
}
/* ---- VV/slicing.cpp: view-aligned slices (ViewSlicing::setupSlicing / drawSlice) and the clip-plane cap polygon
 * (ClipPlane::drawSlice, VV/transform.cpp:432-444 = setupSingleSlice(normal, extent) + drawSingleSlice(-(d - 0.0001))).
 * The polygons are captured from the glMultiTexCoord3fvARB / glVertex3fv calls the reference makes. */
static std::vector<float> g_poly_vert, g_poly_tex;
/* draw list: every glBegin/glEnd primitive with the state it was issued under; vertices = position + the texcoord0
 * current at glVertex time (GL 2.1 spec 2.6: current values are latched per vertex) */
struct Mat4 { double m[16]; };
struct DrawRec {
    unsigned program, mode;
    int cull, clip_mask, viewport[4];
    int blend, blend_src, blend_dst;
    unsigned fbo_tex, bound_tex;    /* texture attached as colour target / texture bound last (the slicing pass binds its source, VV/renderer.cpp:1213) */
    Mat4 mv, proj;
    double clip_eye[6][4];
    std::vector<double> v;          /* x y z s t r per vertex */
};
static std::vector<DrawRec> g_draws;
static bool g_record_draws = false, g_in_prim = false;
static float g_cur_tc[3] = {0.0f, 0.0f, 0.0f};
static void draw_vertex(const float *v)
{
    if (!g_record_draws || !g_in_prim || g_draws.empty()) return;
    std::vector<double> &d = g_draws.back().v;
    for (int k = 0; k < 3; ++k) d.push_back((double)v[k]);
    for (int k = 0; k < 3; ++k) d.push_back((double)g_cur_tc[k]);
}
void glVertex3fv(const GLfloat *v) { g_poly_vert.insert(g_poly_vert.end(), v, v + 3); draw_vertex(v); }
void glMultiTexCoord3fvARB(GLenum unit, const GLfloat *v)
{
    g_poly_tex.insert(g_poly_tex.end(), v, v + 3);
    if (unit == GL_TEXTURE0_ARB) for (int k = 0; k < 3; ++k) g_cur_tc[k] = v[k];
}
void glTexCoord3f(GLfloat x, GLfloat y, GLfloat z) { g_cur_tc[0] = x; g_cur_tc[1] = y; g_cur_tc[2] = z; }   /* VolumeBuffer::drawSlice */
void glVertex2f(GLfloat x, GLfloat y) { const float v[3] = {x, y, 0.0f}; draw_vertex(v); }
void glVertex3f(GLfloat x, GLfloat y, GLfloat z) { const float v[3] = {x, y, z}; g_poly_vert.insert(g_poly_vert.end(), v, v + 3); draw_vertex(v); }
void glMultiTexCoord3fARB(GLenum unit, GLfloat x, GLfloat y, GLfloat z)
{
    if (unit != GL_TEXTURE0_ARB) return;
    const float v[3] = {x, y, z};
    g_poly_tex.insert(g_poly_tex.end(), v, v + 3);
    for (int k = 0; k < 3; ++k) g_cur_tc[k] = v[k];
}

static int copy_poly(float *verts, float *tex, int cap)
{
    int n = (int)(g_poly_vert.size() / 3);
    if (n > cap) return -1;
    if (verts) std::memcpy(verts, g_poly_vert.data(), g_poly_vert.size() * sizeof(float));
    if (tex) std::memcpy(tex, g_poly_tex.data(), g_poly_tex.size() * sizeof(float));
    return n;
}

/* out5 = view vector (3), covered depth d, number of slices */
int vvref_slicing_setup(const float mv[16], float samp_dist, const float ext[3], float out5[5])
{
    ViewSlicing vs;
    float m[16], e[3] = {ext[0], ext[1], ext[2]};
    std::memcpy(m, mv, sizeof(m));
    int n = vs.setupSlicing(m, samp_dist, e);
    out5[0] = vs._v[0]; out5[1] = vs._v[1]; out5[2] = vs._v[2]; out5[3] = vs._d; out5[4] = (float)n;
    return n;
}

/* polygon of slice `slice` (front to back as Renderer::sliceVolume draws them): returns the vertex count */
int vvref_slice_polygon(const float mv[16], float samp_dist, const float ext[3], int slice, float *verts, float *tex, int cap)
{
    ViewSlicing vs;
    float m[16], e[3] = {ext[0], ext[1], ext[2]};
    std::memcpy(m, mv, sizeof(m));
    vs.setupSlicing(m, samp_dist, e);
    g_poly_vert.clear(); g_poly_tex.clear();
    vs.drawSlice(slice);
    return copy_poly(verts, tex, cap);
}

/* cap polygon of a user clip plane (n.xyz, d) */
int vvref_clip_cap_polygon(const double plane[4], const float ext[3], float *verts, float *tex, int cap)
{
    ViewSlicing vs;
    double n[4] = {plane[0], plane[1], plane[2], plane[3]};
    float e[3] = {ext[0], ext[1], ext[2]};
    vs.setupSingleSlice(n, e);
    g_poly_vert.clear(); g_poly_tex.clear();
    vs.drawSingleSlice((float)-(n[3] - 0.0001));
    return copy_poly(verts, tex, cap);
}

/* ---- GL matrix stack / uniform / light capture for VV/renderer.cpp, camera.cpp, transform.cpp -------------------
 * The matrix arithmetic is the OpenGL 2.1 specification's (section 2.11.2: Translate, Rotate; gluPerspective per the
 * GLU reference), in double; what the reference contributes is the sequence of calls and their arguments. */
static Mat4 mat_identity() { Mat4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return r; }
static Mat4 mat_mul(const Mat4 &a, const Mat4 &b)      /* column-major a * b */
{
    Mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a.m[4 * k + rr] * b.m[4 * c + k];
            r.m[4 * c + rr] = s;
        }
    return r;
}
static std::vector<Mat4> g_mv(1, mat_identity()), g_proj(1, mat_identity());
static GLenum g_mode = GL_MODELVIEW;
static std::vector<Mat4> &stack() { return g_mode == GL_PROJECTION ? g_proj : g_mv; }
static std::map<int, std::vector<float>> g_uniform;
static float g_light_pos[4];
static double g_clip[6][4];
static int g_viewport[4];

void glMatrixMode(GLenum mode) { g_mode = mode; }
void glLoadIdentity(void) { stack().back() = mat_identity(); }
void glPushMatrix(void) { stack().push_back(stack().back()); }
void glPopMatrix(void) { if (stack().size() > 1) stack().pop_back(); }
void glTranslated(GLdouble x, GLdouble y, GLdouble z)
{
    Mat4 t = mat_identity();
    t.m[12] = x; t.m[13] = y; t.m[14] = z;
    stack().back() = mat_mul(stack().back(), t);
}
void glTranslatef(GLfloat x, GLfloat y, GLfloat z) { glTranslated(x, y, z); }
void glRotatef(GLfloat angle, GLfloat x, GLfloat y, GLfloat z)
{
    double len = std::sqrt((double)x * x + (double)y * y + (double)z * z);
    if (len == 0.0) return;                               /* degenerate axis: implementations leave the matrix alone */
    double ux = x / len, uy = y / len, uz = z / len, a = (double)angle * M_PI / 180.0, c = std::cos(a), s = std::sin(a), t = 1.0 - c;
    Mat4 r = mat_identity();
    r.m[0] = ux * ux * t + c;      r.m[4] = ux * uy * t - uz * s; r.m[8] = ux * uz * t + uy * s;
    r.m[1] = uy * ux * t + uz * s; r.m[5] = uy * uy * t + c;      r.m[9] = uy * uz * t - ux * s;
    r.m[2] = uz * ux * t - uy * s; r.m[6] = uz * uy * t + ux * s; r.m[10] = uz * uz * t + c;
    stack().back() = mat_mul(stack().back(), r);
}
void gluPerspective(GLdouble fovy, GLdouble aspect, GLdouble zn, GLdouble zf)
{
    double f = 1.0 / std::tan(fovy * M_PI / 360.0);
    Mat4 p;
    for (int i = 0; i < 16; ++i) p.m[i] = 0.0;
    p.m[0] = f / aspect; p.m[5] = f; p.m[10] = (zf + zn) / (zn - zf); p.m[11] = -1.0; p.m[14] = 2.0 * zf * zn / (zn - zf);
    stack().back() = mat_mul(stack().back(), p);
}
void glGetDoublev(GLenum pname, GLdouble *out)
{
    const Mat4 &m = (pname == GL_PROJECTION_MATRIX) ? g_proj.back() : g_mv.back();
    for (int i = 0; i < 16; ++i) out[i] = m.m[i];
}
void glGetFloatv(GLenum pname, GLfloat *out)
{
    if (pname != GL_MODELVIEW_MATRIX && pname != GL_PROJECTION_MATRIX) { for (int i = 0; i < 4; ++i) out[i] = 0.0f; return; }
    double d[16];
    glGetDoublev(pname, d);
    for (int i = 0; i < 16; ++i) out[i] = (float)d[i];
}
void glGetIntegerv(GLenum pname, GLint *out) { for (int i = 0; i < 4; ++i) out[i] = (pname == GL_VIEWPORT) ? g_viewport[i] : 0; }
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h) { g_viewport[0] = x; g_viewport[1] = y; g_viewport[2] = w; g_viewport[3] = h; }
void glLightfv(GLenum, GLenum pname, const GLfloat *v)
{
    if (pname != GL_POSITION) return;
    const Mat4 &m = g_mv.back();                          /* positions are transformed by the current model-view */
    for (int r = 0; r < 4; ++r)
        g_light_pos[r] = (float)(m.m[r] * v[0] + m.m[4 + r] * v[1] + m.m[8 + r] * v[2] + m.m[12 + r] * v[3]);
}
/* inverse of a general 4x4 (cofactors), column-major */
static bool mat_inverse(const Mat4 &a, Mat4 &out)
{
    const double *m = a.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0) return false;
    for (int i = 0; i < 16; ++i) out.m[i] = inv[i] / det;
    return true;
}
static double g_clip_eye[6][4];
static int g_clip_mask = 0, g_cull = 0, g_blend = 0, g_blend_src = GL_ONE, g_blend_dst = GL_ZERO;
static GLhandleARB g_program = 0;
/* GL 2.1 spec 2.12: the plane is stored in eye coordinates, (p1' p2' p3' p4') = (p1 p2 p3 p4) M^-1 with the model-view
 * of the time of the call */
void glClipPlane(GLenum plane, const GLdouble *eq)
{
    int i = (int)plane - GL_CLIP_PLANE0;
    if (i < 0 || i >= 6) return;
    for (int k = 0; k < 4; ++k) g_clip[i][k] = eq[k];
    Mat4 inv;
    if (!mat_inverse(g_mv.back(), inv)) inv = mat_identity();
    for (int c = 0; c < 4; ++c)
        g_clip_eye[i][c] = eq[0] * inv.m[4 * c + 0] + eq[1] * inv.m[4 * c + 1] + eq[2] * inv.m[4 * c + 2] + eq[3] * inv.m[4 * c + 3];
}
void glEnable(GLenum cap)
{
    if (cap == GL_CULL_FACE) g_cull = 1;
    else if (cap == GL_BLEND) g_blend = 1;
    else if (cap >= GL_CLIP_PLANE0 && cap < GL_CLIP_PLANE0 + 6) g_clip_mask |= 1 << (cap - GL_CLIP_PLANE0);
}
void glDisable(GLenum cap)
{
    if (cap == GL_CULL_FACE) g_cull = 0;
    else if (cap == GL_BLEND) g_blend = 0;
    else if (cap >= GL_CLIP_PLANE0 && cap < GL_CLIP_PLANE0 + 6) g_clip_mask &= ~(1 << (cap - GL_CLIP_PLANE0));
}
void glUseProgramObjectARB(GLhandleARB program) { g_program = program; }
void glBlendFunc(GLenum src, GLenum dst) { g_blend_src = (int)src; g_blend_dst = (int)dst; }
static unsigned g_fbo_tex = 0;
void glFramebufferTexture2DEXT(GLenum, GLenum attachment, GLenum, GLuint texture, GLint) { if (attachment == GL_COLOR_ATTACHMENT0_EXT) g_fbo_tex = texture; }
void glBegin(GLenum mode)
{
    g_in_prim = true;
    if (!g_record_draws) return;
    DrawRec d;
    d.program = g_program; d.mode = mode; d.cull = g_cull; d.clip_mask = g_clip_mask;
    d.blend = g_blend; d.blend_src = g_blend_src; d.blend_dst = g_blend_dst;
    d.fbo_tex = g_fbo_tex; d.bound_tex = g_bound;
    for (int k = 0; k < 4; ++k) d.viewport[k] = g_viewport[k];
    d.mv = g_mv.back(); d.proj = g_proj.back();
    std::memcpy(d.clip_eye, g_clip_eye, sizeof(d.clip_eye));
    g_draws.push_back(d);
}
void glEnd(void) { g_in_prim = false; }
void glUniform1iARB(GLint loc, GLint v) { g_uniform[loc] = {(float)v}; }
void glUniform1fARB(GLint loc, GLfloat v) { g_uniform[loc] = {v}; }
void glUniform3fARB(GLint loc, GLfloat a, GLfloat b, GLfloat c) { g_uniform[loc] = {a, b, c}; }
void glUniform4fARB(GLint loc, GLfloat a, GLfloat b, GLfloat c, GLfloat d) { g_uniform[loc] = {a, b, c, d}; }
void glUniform4fvARB(GLint loc, GLsizei, const GLfloat *v) { g_uniform[loc] = {v[0], v[1], v[2], v[3]}; }
void glUniform4iARB(GLint loc, GLint a, GLint b, GLint c, GLint d) { g_uniform[loc] = {(float)a, (float)b, (float)c, (float)d}; }

/* Renderer + Camera + Transform driven as VV/3DLIC.cpp does (init :733-760, display :93-126, keyboard F3 / 'l'):
 *   out[0..15]   GL_MODELVIEW of the frame: Camera::setCamera() + glTranslatef(-center)      (VV/renderer.cpp:134-146)
 *   out[16..19]  gl_LightSource[0].position after Renderer::updateLightPos()                  (VV/renderer.cpp:431-466)
 *   out[20..24]  ViewSlicing state after Renderer::updateSlices(): v[3], d, numSlices          (VV/renderer.cpp:1270-1292)
 *   out[25..]    uniforms of Renderer::setRenderVolParams, 4 floats per slot in the order
 *                texMax, scaleVol, scaleVolInv(-> what ended up in slot scaleVol when has_scalevolinv), stepSize, gradient,
 *                licParams, licKernel, numIterations, alphaCorrection, viewport                (VV/renderer.cpp:925-996)
 * has_scalevolinv: whether the program has an active scaleVolInv uniform (ILLUM_* builds of the ray-cast program, Q1). */
int vvref_renderer_state(const char *dat, const char *filter_png, const float cam_quat[4], const float cam_pos[3], float cam_dist,
                         const float light_quat[4], float light_dist, const float lic[8], int lowres, int has_scalevolinv,
                         int width, int height, float *out)
{
    VectorDataSet vd;
    if (!vd.loadData(dat)) return -10;
    LICFilter filt;
    if (!filter_png || !filt.loadData(filter_png)) filt.createBoxFilter();
    LICParams lp;
    lp.stepSizeVol = lic[0]; lp.gradientScale = lic[1]; lp.illumScale = lic[2]; lp.freqScale = lic[3];
    lp.numIterations = (int)lic[4]; lp.stepsForward = (int)lic[5]; lp.stepsBackward = (int)lic[6]; lp.stepSizeLIC = lic[7];
    Camera cam;
    Quaternion q; q.x = cam_quat[0]; q.y = cam_quat[1]; q.z = cam_quat[2]; q.w = cam_quat[3];
    cam.rotate(q);                                          /* (Transform::setQuaternion only ever writes .x) */
    cam.setPosition(Vector3_new(cam_pos[0], cam_pos[1], cam_pos[2]));
    cam.setDistance(cam_dist);
    cam.setWindow(width, height);
    Transform light;
    Quaternion ql; ql.x = light_quat[0]; ql.y = light_quat[1]; ql.z = light_quat[2]; ql.w = light_quat[3];
    light.rotate(ql);
    light.setDistance(light_dist);
    Renderer r;
    r.setVolumeData(vd.getVolumeData());
    r._licFilter = &filt;
    r.setLICParams(&lp);
    r.setCamera(&cam);
    r.setLight(&light);
    r._winWidth = width; r._winHeight = height;
    r.enableLowRes(lowres != 0);
    g_mv.assign(1, mat_identity()); g_proj.assign(1, mat_identity()); g_mode = GL_MODELVIEW;
    g_uniform.clear();
    /* frame model-view, Renderer::render :134-146 */
    cam.setCamera();
    VolumeData *v = vd.getVolumeData();
    glTranslatef(-v->center[0], -v->center[1], -v->center[2]);
    glGetFloatv(GL_MODELVIEW_MATRIX, out);
    r.updateLightPos();
    for (int i = 0; i < 4; ++i) out[16 + i] = g_light_pos[i];
    r.updateSlices();
    out[20] = r._slices._v[0]; out[21] = r._slices._v[1]; out[22] = r._slices._v[2]; out[23] = r._slices._d; out[24] = (float)r._slices._numSlices;
    GLSLParamsLIC p;
    p.texMax = 1; p.scaleVol = 2; p.scaleVolInv = has_scalevolinv ? 3 : -1; p.stepSize = 4; p.gradient = 5; p.licParams = 6;
    p.licKernel = 7; p.numIterations = 8; p.alphaCorrection = 9; p.viewport = 10;
    r.setRenderVolParams(&p);
    for (int slot = 1; slot <= 10; ++slot) {
        float *o = out + 25 + 4 * (slot - 1);
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        auto it = g_uniform.find(slot);
        if (it != g_uniform.end()) for (size_t k = 0; k < it->second.size() && k < 4; ++k) o[k] = it->second[k];
    }
    r._licFilter = NULL;                                    /* stack objects: nothing for ~Renderer to touch */
    return 0;
}

/* the proxy cube of Renderer::drawCubeFaces (VV/renderer.cpp:682-736): 6 quads, texcoord0 + position per vertex */
int vvref_cube_faces(const char *dat, float *verts, float *tex, int cap)
{
    VectorDataSet vd;
    if (!vd.loadData(dat)) return -10;
    Renderer r;
    r.setVolumeData(vd.getVolumeData());
    g_poly_vert.clear(); g_poly_tex.clear();
    r.drawCubeFaces();
    return copy_poly(verts, tex, cap);
}

static int serialize_draws(double *out, int cap)
{
    size_t need = 1;
    for (const DrawRec &d : g_draws) need += 4 + 3 + 2 + 4 + 16 + 16 + 24 + 1 + d.v.size();
    if ((size_t)cap < need) { g_draws.clear(); return -2; }
    size_t k = 0;
    out[k++] = (double)g_draws.size();
    for (const DrawRec &d : g_draws) {
        out[k++] = d.program; out[k++] = d.mode; out[k++] = d.cull; out[k++] = d.clip_mask;
        out[k++] = d.blend; out[k++] = d.blend_src; out[k++] = d.blend_dst;
        out[k++] = d.fbo_tex; out[k++] = d.bound_tex;
        for (int i = 0; i < 4; ++i) out[k++] = d.viewport[i];
        for (int i = 0; i < 16; ++i) out[k++] = d.mv.m[i];
        for (int i = 0; i < 16; ++i) out[k++] = d.proj.m[i];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) out[k++] = d.clip_eye[i][j];
        out[k++] = (double)(d.v.size() / 6);
        for (double x : d.v) out[k++] = x;
    }
    g_draws.clear();
    return (int)k;
}

/* One frame of Renderer::render(true) in ray-cast mode (VV/renderer.cpp:126-312), run unmodified; every glBegin/glEnd primitive is
 * recorded with the program, matrices, viewport, cull / clip state it was issued under.  Serialised as doubles:
 *   out[0] = number of primitives, then per primitive
 *   program, mode, cull, clip_mask, blend enabled, blend src, blend dst, viewport[4], modelview[16], projection[16], clip planes in eye space [6][4], nverts,
 *   nverts x (x y z s t r)
 * The ray-cast program has the handle 77, the slicing program 79 (set below; no GLSL is compiled in the shim); slicing = 1
 * renders with VOLIC_SLICING through the FBO ping-pong (after Renderer::updateSlices), 3 with VOLIC_SLICING without the FBO (program 82 +
 * fixed-function blending), 2 with VOLIC_LICVOLUME (program 81) instead of VOLIC_RAYCAST; per primitive also the texture attached as
 * colour target and the texture bound last (after blend dst); step_size_vol > 0 overrides LICParams.  planes: up to 3 user clip planes
 * (n.xyz, d) as ClipPlane::setNormal takes them, active[i] != 0 to enable; frames >= 1: how many frames to render, the last one
 * is the one returned.  Returns the number of doubles written, < 0 on error. */
int vvref_raycast_draws(const char *dat, const float cam_quat[4], const float cam_pos[3], float cam_dist, int width, int height,
                        int lowres, const double planes[3][4], const int active[3], int frames, int slicing, float step_size_vol,
                        double *out, int cap)
{
    VectorDataSet vd;
    if (!vd.loadData(dat)) return -10;
    LICFilter filt;
    filt.createBoxFilter();
    LICParams lp;
    if (step_size_vol > 0.0f) lp.stepSizeVol = step_size_vol;
    Camera cam;
    Quaternion q; q.x = cam_quat[0]; q.y = cam_quat[1]; q.z = cam_quat[2]; q.w = cam_quat[3];
    cam.rotate(q);
    cam.setPosition(Vector3_new(cam_pos[0], cam_pos[1], cam_pos[2]));
    cam.setDistance(cam_dist);
    cam.setWindow(width, height);
    Transform light;
    ClipPlane cp[3];
    VolumeData *v = vd.getVolumeData();
    for (int i = 0; i < 3; ++i) {
        cp[i].setPlaneId(GL_CLIP_PLANE0 + i);                          /* VV/3DLIC.cpp:768-781 */
        cp[i].setBoundingBox(-v->extent[0] / 2.0f, -v->extent[1] / 2.0f, -v->extent[2] / 2.0f, v->extent[0] / 2.0f, v->extent[1] / 2.0f, v->extent[2] / 2.0f);
        cp[i].setNormal(planes[i][0], planes[i][1], planes[i][2], planes[i][3]);
        cp[i].setActive(active[i] != 0);
    }
    Texture dummy[8];
    Renderer r;
    r.setVolumeData(v);
    r._licFilter = &filt;
    r.setLICParams(&lp);
    r.setCamera(&cam);
    r.setLight(&light);
    r.setClipPlanes(cp, 3);
    r._dataTex = &dummy[0]; r._tfRGBTex = &dummy[1]; r._tfAlphaOpacTex = &dummy[2]; r._noiseTex = &dummy[3];
    r._scalarTex = &dummy[4]; r._licKernelTex = &dummy[5];
    r._raycastShader._programObj = 77;
    r._bgShader._programObj = 78;
    std::memset(&r._paramRaycast, 0xff, sizeof(r._paramRaycast));      /* every uniform location -1: no GLSL program in the shim */
    std::memset(&r._paramBackground, 0xff, sizeof(r._paramBackground));
    r._sliceShader._programObj = 79;
    std::memset(&r._paramSlice, 0xff, sizeof(r._paramSlice));
    r._paramSlice.imageFBOSampler = 20;                                 /* sliceVolume returns early without it (:1127) */
    /* the two image textures get their names in Renderer::initFBO (VV/renderer.cpp:547-555), which needs a GL; any two distinct
     * names do -- what is recorded is which of them is the colour target / the bound source of every primitive */
    r._imgBufferTex0->setTex(GL_TEXTURE_RECTANGLE_ARB, 501, "FBO-Tex0");
    r._imgBufferTex0->texUnit = GL_TEXTURE1_ARB;
    r._imgBufferTex1->setTex(GL_TEXTURE_RECTANGLE_ARB, 502, "FBO-Tex1");
    r._imgBufferTex1->texUnit = GL_TEXTURE1_ARB;
    r.enableLowRes(lowres != 0);
    r.resize(width, height);                                            /* VV/3DLIC.cpp:174-200 */
    r._licRaycastShader._programObj = 81;
    std::memset(&r._paramLicRaycast, 0xff, sizeof(r._paramLicRaycast));
    if (slicing == 2) {
        r.setTechnique(VOLIC_LICVOLUME);                                /* ray-cast of the LIC volume (F4) */
    } else if (slicing == 3) {
        r._useFBO = false;                                              /* the start-up state of F3: lic3d_slicingblend + fixed-function blending */
        r._sliceBlendShader._programObj = 82;
        std::memset(&r._paramSliceBlend, 0xff, sizeof(r._paramSliceBlend));
        r.setTechnique(VOLIC_SLICING);
    } else if (slicing) {
        r._useFBO = true;                                               /* the FBO ping-pong branch of sliceVolume (keys F3, VV/3DLIC.cpp:457-475) */
        r.setTechnique(VOLIC_SLICING);
    } else {
        r.setTechnique(VOLIC_RAYCAST);
    }
    g_mv.assign(1, mat_identity()); g_proj.assign(1, mat_identity()); g_mode = GL_MODELVIEW;
    g_clip_mask = 0; g_cull = 0; g_program = 0; g_blend = 0; g_blend_src = GL_ONE; g_blend_dst = GL_ZERO;
    glViewport(0, 0, width, height);                                    /* resize callback, VV/3DLIC.cpp:176 */
    /* `frames` calls of render(true); the LAST one is recorded.  (The first frame differs for non-unit plane normals:
     * drawClippedPolygon normalises ClipPlane::_normal in place, VV/renderer.cpp:1301 -> VV/slicing.cpp:337-348.) */
    if (slicing == 1 || slicing == 3) r.updateSlices();                      /* VV/3DLIC.cpp:168, 469 */
    for (int f = 0; f < frames; ++f) {
        g_draws.clear();
        g_record_draws = true;
        r.render(true);
        g_record_draws = false;
    }
    r._licFilter = NULL;
    return serialize_draws(out, cap);
}

/* Renderer::updateLICVolume -> renderLICVolume (VV/renderer.cpp:1311-1374) into a w x h x d VolumeBuffer (the reference
 * allocates 512^3 in Renderer::init, :97): one screen-filling quad per layer, texcoord0 = (s, t, (z + 0.5) / d).  Same record
 * format as vvref_raycast_draws; the LIC-volume program has the handle 80. */
int vvref_licvolume_draws(int w, int h, int d, double *out, int cap)
{
    LICFilter filt;
    filt.createBoxFilter();
    LICParams lp;
    Texture dummy[3];
    VolumeData vol;
    std::memset(&vol, 0, sizeof(vol));
    for (int i = 0; i < 3; ++i) { vol.size[i] = 8; vol.extent[i] = 1.0f; vol.scale[i] = 1.0f; vol.scaleInv[i] = 1.0f; vol.center[i] = 0.5f; }
    Renderer r;
    r.setVolumeData(&vol);
    r._licFilter = &filt;
    r.setLICParams(&lp);
    r._dataTex = &dummy[0]; r._tfRGBTex = &dummy[1]; r._tfAlphaOpacTex = &dummy[2];
    r._licvolumebuffer = new VolumeBuffer(GL_RGBA16F_ARB, w, h, d, 2);
    r._volumeRenderShader._programObj = 80;
    std::memset(&r._paramLICVolume, 0xff, sizeof(r._paramLICVolume));
    g_mv.assign(1, mat_identity()); g_proj.assign(1, mat_identity()); g_mode = GL_MODELVIEW;
    g_clip_mask = 0; g_cull = 0; g_program = 0; g_blend = 0; g_blend_src = GL_ONE; g_blend_dst = GL_ZERO;
    glViewport(0, 0, 640, 480);
    g_draws.clear();
    g_record_draws = true;
    r.updateLICVolume();
    g_record_draws = false;
    delete r._licvolumebuffer;
    r._licvolumebuffer = NULL;
    r._licFilter = NULL;
    return serialize_draws(out, cap);
}

int vvref_next_pow2(int v) { return nextPowerTwo(v); }
int vvref_has_host(void) { return 1; }

} /* extern "C" */
