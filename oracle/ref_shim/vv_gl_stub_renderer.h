/* vv_gl_stub_renderer.h -- TEST INFRASTRUCTURE ONLY.
 * The rest of the OpenGL surface that VV/renderer.cpp, camera.cpp, transform.cpp, GLSLShader.cpp and VolumeBuffer.cpp
 * need to compile UNMODIFIED.  What matters is captured by oracle/ref_host_driver.cpp: the matrix stack
 * (glMatrixMode / glLoadIdentity / glPush/PopMatrix / glTranslate / glRotatef / glGetFloatv), glUniform*ARB values per
 * location, glLightfv(GL_POSITION) and glClipPlane.  Everything else is a no-op. */
#ifndef VV_GL_STUB_RENDERER_H_
#define VV_GL_STUB_RENDERER_H_

typedef struct GLUquadric GLUquadricObj;
typedef void (*PFNGLGENFRAMEBUFFERSEXTPROC)(GLsizei, GLuint *);

enum {
    GL_FALSE = 0, GL_TRUE = 1, GL_TRIANGLE_FAN = 6, GL_POLYGON_BIT = 0x0008, GL_CLIP_PLANE0 = 0x3000, GL_CLIP_PLANE1, GL_CLIP_PLANE2,
    GL_CLIP_PLANE3, GL_CLIP_PLANE4, GL_CLIP_PLANE5, GL_COLOR_ATTACHMENT0_EXT = 0x8CE0, GL_COLOR_ATTACHMENT1_EXT,
    GL_DEPTH_ATTACHMENT_EXT = 0x8D00, GL_RENDERBUFFER_EXT = 0x8D41, GL_DEPTH_COMPONENT24 = 0x81A6, GL_DEPTH_COMPONENT = 0x1902,
    GL_OBJECT_INFO_LOG_LENGTH_ARB = 0x8B84, GL_OBJECT_COMPILE_STATUS_ARB = 0x8B81, GL_OBJECT_ACTIVE_UNIFORMS_ARB = 0x8B86,
    GL_OBJECT_LINK_STATUS_ARB = 0x8B82, GL_VERTEX_SHADER_ARB = 0x8B31, GL_FRAGMENT_SHADER_ARB = 0x8B30,
    GL_OBJECT_ACTIVE_UNIFORM_MAX_LENGTH_ARB = 0x8B87,
    GL_FRONT_AND_BACK = 0x0408, GL_FRONT = 0x0404, GL_BACK = 0x0405, GL_FILL = 0x1B02, GL_LINE = 0x1B01, GL_MODELVIEW_MATRIX = 0x0BA6,
    GL_PROJECTION_MATRIX = 0x0BA7, GL_VIEWPORT = 0x0BA2, GL_POSITION = 0x1203, GL_LIGHT0 = 0x4000, GL_LIGHT1 = 0x4001, GL_AMBIENT = 0x1200,
    GL_DIFFUSE = 0x1201, GL_SPECULAR = 0x1202, GL_SHININESS = 0x1601, GL_SPOT_EXPONENT = 0x1205, GL_DEPTH_BUFFER_BIT = 0x0100,
    GL_RGBA8 = 0x8058, GL_RGBA32F_ARB = 0x8814, GL_RGB16F_ARB = 0x881B, GL_ONE = 1, GL_ZERO = 0, GL_LESS = 0x0201, GL_LEQUAL = 0x0203,
    GL_NORMALIZE = 0x0BA1, GL_COLOR_MATERIAL = 0x0B57, GL_SMOOTH = 0x1D01, GL_FLAT = 0x1D00, GL_TEXTURE_BIT = 0x00040000,
    GL_LIGHTING_BIT = 0x0040, GL_CURRENT_BIT = 0x0001, GL_DEPTH_BUFFER_BIT_ = 0, GL_VIEWPORT_BIT = 0x0800, GL_ALL_ATTRIB_BITS = 0x000fffff,
    GLU_FILL = 100012, GLU_SMOOTH = 100000, GL_POLYGON = 9, GL_TRIANGLES = 4, GL_TRIANGLE_STRIP = 5, GL_QUAD_STRIP = 8,
    GL_PACK_ALIGNMENT = 0x0D05, GL_UNPACK_ALIGNMENT = 0x0CF5, GL_AMBIENT_AND_DIFFUSE = 0x1602, GL_EMISSION = 0x1600,
    GL_RED = 0x1903, GL_ONE_MINUS_DST_ALPHA = 0x0305, GL_DST_ALPHA = 0x0304, GL_FRAMEBUFFER_BINDING_EXT = 0x8CA6, GL_COLOR_CLEAR_VALUE = 0x0C22,
    GL_INT = 0x1404, GL_FLOAT_VEC2_ARB = 0x8B50, GL_FLOAT_VEC3_ARB, GL_FLOAT_VEC4_ARB, GL_INT_VEC2_ARB, GL_INT_VEC3_ARB, GL_INT_VEC4_ARB,
    GL_BOOL_ARB, GL_BOOL_VEC2_ARB, GL_BOOL_VEC3_ARB, GL_BOOL_VEC4_ARB, GL_FLOAT_MAT2_ARB, GL_FLOAT_MAT3_ARB, GL_FLOAT_MAT4_ARB,
    GL_SAMPLER_1D_ARB, GL_SAMPLER_2D_ARB, GL_SAMPLER_3D_ARB, GL_SAMPLER_CUBE_ARB, GL_SAMPLER_1D_SHADOW_ARB, GL_SAMPLER_2D_SHADOW_ARB,
    GL_SAMPLER_2D_RECT_ARB, GL_SAMPLER_2D_RECT_SHADOW_ARB
};

#ifdef __cplusplus
extern "C" {
#endif
/* ---- captured by the driver ---- */
void glRotatef(GLfloat angle, GLfloat x, GLfloat y, GLfloat z);
void glTranslated(GLdouble x, GLdouble y, GLdouble z);
void glGetFloatv(GLenum pname, GLfloat *out);
void glGetDoublev(GLenum pname, GLdouble *out);
void glGetIntegerv(GLenum pname, GLint *out);
void gluPerspective(GLdouble fovy, GLdouble aspect, GLdouble znear, GLdouble zfar);
void glLightfv(GLenum light, GLenum pname, const GLfloat *v);
void glClipPlane(GLenum plane, const GLdouble *eq);
void glUniform1iARB(GLint loc, GLint v);
void glUniform1fARB(GLint loc, GLfloat v);
void glUniform3fARB(GLint loc, GLfloat a, GLfloat b, GLfloat c);
void glUniform4fARB(GLint loc, GLfloat a, GLfloat b, GLfloat c, GLfloat d);
void glUniform4fvARB(GLint loc, GLsizei n, const GLfloat *v);
void glUniform4iARB(GLint loc, GLint a, GLint b, GLint c, GLint d);
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h);
void glMultiTexCoord3fARB(GLenum unit, GLfloat x, GLfloat y, GLfloat z);
void glTexCoord3f(GLfloat s, GLfloat t, GLfloat r);      /* captured: current texcoord0 */
void glUseProgramObjectARB(GLhandleARB program);       /* the program bound when a primitive is drawn */
void glFramebufferTexture2DEXT(GLenum target, GLenum attachment, GLenum textarget, GLuint texture, GLint level);   /* captured: the
                                                        colour target a primitive is drawn into (the slicing pass's ping-pong) */
/* ---- no-ops ---- */
static inline void glNormal3f(GLfloat, GLfloat, GLfloat) {}
static inline void glDepthMask(GLboolean) {}
static inline void glBindFramebufferEXT(GLenum, GLuint) {}
static inline void glBindRenderbufferEXT(GLenum, GLuint) {}
static inline void glFramebufferTexture1DEXT(GLenum, GLenum, GLenum, GLuint, GLint) {}
static inline void glFramebufferTexture3DEXT(GLenum, GLenum, GLenum, GLuint, GLint, GLint) {}
static inline void glFramebufferRenderbufferEXT(GLenum, GLenum, GLenum, GLuint) {}
static inline void glRenderbufferStorageEXT(GLenum, GLenum, GLsizei, GLsizei) {}
/* GLEW exposes extension entry points as assignable function pointers (VV/renderer.cpp:545 re-loads this one) */
static inline void vv_stub_gen_framebuffers(GLsizei n, GLuint *ids) { for (int i = 0; i < n; ++i) ids[i] = 1000 + i; }
static PFNGLGENFRAMEBUFFERSEXTPROC glGenFramebuffersEXT = vv_stub_gen_framebuffers;
/* the reference only ever asks for "glGenFramebuffersEXT" (VV/renderer.cpp:545, VV/VolumeBuffer.cpp:9) */
static inline void *wglGetProcAddress(const char *) { return (void *)vv_stub_gen_framebuffers; }
static inline void glGenRenderbuffersEXT(GLsizei n, GLuint *ids) { for (int i = 0; i < n; ++i) ids[i] = 2000 + i; }
static inline void glDeleteFramebuffersEXT(GLsizei, const GLuint *) {}
static inline void glDeleteRenderbuffersEXT(GLsizei, const GLuint *) {}
static inline void glCopyTexSubImage3D(GLenum, GLint, GLint, GLint, GLint, GLint, GLint, GLsizei, GLsizei) {}
static inline void glCopyTexImage2D(GLenum, GLint, GLenum, GLint, GLint, GLsizei, GLsizei, GLint) {}
/* compile / link status: success; info-log lengths and active-uniform counts: 0 */
static inline void glGetObjectParameterivARB(GLhandleARB, GLenum pname, GLint *v)
{ if (v) *v = (pname == GL_OBJECT_COMPILE_STATUS_ARB || pname == GL_OBJECT_LINK_STATUS_ARB) ? 1 : 0; }
void glShaderSourceARB(GLhandleARB, GLsizei n, const GLcharARB **src, const GLint *len);   /* captured: the defines string (ref_app_driver.cpp) */
static inline void glGetInfoLogARB(GLhandleARB, GLsizei, GLsizei *len, GLcharARB *log) { if (len) *len = 0; if (log) log[0] = 0; }
static inline void glGetActiveUniformARB(GLhandleARB, GLuint, GLsizei, GLsizei *len, GLint *size, GLenum *type, GLcharARB *name)
{ if (len) *len = 0; if (size) *size = 0; if (type) *type = 0; if (name) name[0] = 0; }
static inline GLint glGetUniformLocationARB(GLhandleARB, const GLcharARB *) { return -1; }
static inline void glDetachObjectARB(GLhandleARB, GLhandleARB) {}
static inline void glCompileShaderARB(GLhandleARB) {}
static inline void glAttachObjectARB(GLhandleARB, GLhandleARB) {}
static inline void glLinkProgramARB(GLhandleARB) {}
static inline void glDeleteObjectARB(GLhandleARB) {}
static inline GLhandleARB glCreateShaderObjectARB(GLenum) { return 1; }
static inline GLhandleARB glCreateProgramObjectARB(void) { return 2; }
static inline void glPolygonMode(GLenum, GLenum) {}
static inline void glGetTexImage(GLenum, GLint, GLenum, GLenum, GLvoid *) {}
static inline void glClearColor(GLclampf, GLclampf, GLclampf, GLclampf) {}
static inline void glClear(GLbitfield) {}
static inline void glColorMask(GLboolean, GLboolean, GLboolean, GLboolean) {}
static inline void glReadBuffer(GLenum) {}
static inline void glReadPixels(GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, GLvoid *) {}
static inline void glMaterialf(GLenum, GLenum, GLfloat) {}
static inline void glMaterialfv(GLenum, GLenum, const GLfloat *) {}
static inline void glVertex3d(GLdouble, GLdouble, GLdouble) {}
static inline void glVertex3dv(const GLdouble *) {}
static inline void glMultiTexCoord4fARB(GLenum, GLfloat, GLfloat, GLfloat, GLfloat) {}
static inline void glMultiTexCoord3dARB(GLenum, GLdouble, GLdouble, GLdouble) {}
static inline void glMultiTexCoord3dvARB(GLenum, const GLdouble *) {}
static inline void gluOrtho2D(GLdouble, GLdouble, GLdouble, GLdouble) {}
static inline void gluLookAt(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble) {}
static inline int gluUnProject(GLdouble, GLdouble, GLdouble, const GLdouble *, const GLdouble *, const GLint *, GLdouble *x, GLdouble *y, GLdouble *z)
{ if (x) *x = 0; if (y) *y = 0; if (z) *z = 0; return 1; }
static inline GLUquadricObj *gluNewQuadric(void) { return (GLUquadricObj *)0; }
static inline void gluDeleteQuadric(GLUquadricObj *) {}
static inline void gluQuadricDrawStyle(GLUquadricObj *, GLenum) {}
static inline void gluQuadricNormals(GLUquadricObj *, GLenum) {}
static inline void gluCylinder(GLUquadricObj *, GLdouble, GLdouble, GLdouble, GLint, GLint) {}
static inline void gluDisk(GLUquadricObj *, GLdouble, GLdouble, GLint, GLint) {}
#ifdef __cplusplus
}
#endif
#include "vv_glut_stub.h"
#endif
