/* vv_gl_stub.h -- TEST INFRASTRUCTURE ONLY.
 * A capturing stand-in for <GL/glew.h> + <GL/freeglut.h>: just enough OpenGL 1.x/2.x surface for the reference's
 * host translation units (VV/dataset.cpp, transferEdit.cpp, texture.cpp, gradient.cpp, reader.cpp, parseArg.cpp,
 * mmath.cpp) to compile UNMODIFIED.  Texture uploads (glTexImage1D/3D) and sampler state (glTexParameteri) are
 * recorded per texture id so the test driver can read back exactly what the reference would have handed to the GL. */
#ifndef VV_GL_STUB_H_
#define VV_GL_STUB_H_
#include <stddef.h>
#include <stdio.h>

typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef double GLdouble;
typedef unsigned char GLubyte;
typedef unsigned char GLboolean;
typedef void GLvoid;
typedef unsigned int GLbitfield;
typedef float GLclampf;
typedef char GLcharARB;
typedef unsigned int GLhandleARB;

enum {
    GL_NO_ERROR = 0, GL_TEXTURE_1D = 0x0DE0, GL_TEXTURE_2D = 0x0DE1, GL_TEXTURE_3D = 0x806F, GL_TEXTURE_RECTANGLE_ARB = 0x84F5,
    GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_MIN_FILTER = 0x2801, GL_TEXTURE_WRAP_S = 0x2802, GL_TEXTURE_WRAP_T = 0x2803,
    GL_TEXTURE_WRAP_R = 0x8072, GL_NEAREST = 0x2600, GL_LINEAR = 0x2601, GL_CLAMP = 0x2900, GL_REPEAT = 0x2901,
    GL_CLAMP_TO_EDGE = 0x812F, GL_RGBA = 0x1908, GL_RGB = 0x1907, GL_LUMINANCE = 0x1909, GL_LUMINANCE_ALPHA = 0x190A,
    GL_RGBA16F_ARB = 0x881A, GL_LUMINANCE16F_ARB = 0x881E, GL_LUMINANCE_ALPHA16F_ARB = 0x881F, GL_UNSIGNED_BYTE = 0x1401,
    GL_FLOAT = 0x1406, GL_UNSIGNED_SHORT = 0x1403, GL_TEXTURE_ENV = 0x2300, GL_TEXTURE_ENV_MODE = 0x2200, GL_REPLACE = 0x1E01,
    GL_TEXTURE0_ARB = 0x84C0, GL_TEXTURE1_ARB, GL_TEXTURE2_ARB, GL_TEXTURE3_ARB, GL_TEXTURE4_ARB, GL_TEXTURE5_ARB, GL_TEXTURE6_ARB,
    GL_TEXTURE7_ARB, GL_TEXTURE8_ARB, GL_TEXTURE9_ARB, GL_TEXTURE10_ARB, GL_TEXTURE11_ARB,
    GL_QUADS = 7, GL_LINE_STRIP = 3, GL_LINES = 1, GL_LINE_LOOP = 2, GL_POINTS = 0, GL_PROJECTION = 0x1701, GL_MODELVIEW = 0x1700,
    GL_BLEND = 0x0BE2, GL_DEPTH_TEST = 0x0B71, GL_SRC_ALPHA = 0x0302, GL_ONE_MINUS_SRC_ALPHA = 0x0303, GL_TRANSFORM_BIT = 0x1000,
    GL_ENABLE_BIT = 0x2000, GL_COLOR_BUFFER_BIT = 0x4000, GL_FRAMEBUFFER_EXT = 0x8D40, GL_FRAMEBUFFER_COMPLETE_EXT = 0x8CD5,
    GL_FRAMEBUFFER_INCOMPLETE_ATTACHMENT_EXT, GL_FRAMEBUFFER_INCOMPLETE_MISSING_ATTACHMENT_EXT, GL_FRAMEBUFFER_INCOMPLETE_DIMENSIONS_EXT = 0x8CD9,
    GL_FRAMEBUFFER_INCOMPLETE_FORMATS_EXT, GL_FRAMEBUFFER_INCOMPLETE_DRAW_BUFFER_EXT, GL_FRAMEBUFFER_INCOMPLETE_READ_BUFFER_EXT,
    GL_FRAMEBUFFER_UNSUPPORTED_EXT, GL_FRAMEBUFFER_STATUS_ERROR_EXT = 0x8CDE, GL_INVALID_FRAMEBUFFER_OPERATION_EXT = 0x0506,
    GL_LINE_SMOOTH = 0x0B20, GL_LIGHTING = 0x0B50, GL_CULL_FACE = 0x0B44, GL_POINT_SMOOTH = 0x0B10
};
#define GLUT_BITMAP_HELVETICA_12 ((void *)7)
#define GLUT_BITMAP_HELVETICA_10 ((void *)6)
#define GLUT_BITMAP_8_BY_13 ((void *)3)

#ifdef __cplusplus
extern "C" {
#endif
/* recorded textures */
typedef struct VVStubTex {
    GLenum target;
    GLint internal_format;
    GLenum format, type;
    int dim[3];
    void *data;
    size_t bytes;
    GLint min_filter, mag_filter, wrap_s, wrap_t, wrap_r;
} VVStubTex;
VVStubTex *vv_stub_texture(GLuint id);
GLuint vv_stub_last_texture(void);
void vv_stub_reset(void);

void glGenTextures(GLsizei n, GLuint *ids);
void glDeleteTextures(GLsizei n, const GLuint *ids);
void glBindTexture(GLenum target, GLuint id);
void glTexParameteri(GLenum target, GLenum pname, GLint v);
void glTexImage1D(GLenum target, GLint level, GLint ifmt, GLsizei w, GLint border, GLenum fmt, GLenum type, const void *data);
void glTexImage2D(GLenum target, GLint level, GLint ifmt, GLsizei w, GLsizei h, GLint border, GLenum fmt, GLenum type, const void *data);
void glTexImage3D(GLenum target, GLint level, GLint ifmt, GLsizei w, GLsizei h, GLsizei d, GLint border, GLenum fmt, GLenum type, const void *data);
#define glTexImage3DEXT glTexImage3D
static inline void glTexEnvi(GLenum, GLenum, GLint) {}
static inline void glActiveTextureARB(GLenum) {}
static inline GLenum glGetError(void) { return GL_NO_ERROR; }
static inline const GLubyte *gluErrorString(GLenum) { return (const GLubyte *)"stub"; }
static inline GLenum glCheckFramebufferStatusEXT(GLenum) { return GL_FRAMEBUFFER_COMPLETE_EXT; }
void glMatrixMode(GLenum mode);
void glPushMatrix(void);
void glPopMatrix(void);
void glLoadIdentity(void);
static inline void glOrtho(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble) {}
void glTranslatef(GLfloat x, GLfloat y, GLfloat z);
static inline void glScalef(GLfloat, GLfloat, GLfloat) {}
void glBegin(GLenum mode);                              /* captured: draw list (oracle/softgl.py rasterises it) */
void glEnd(void);
static inline void glVertex2i(GLint, GLint) {}
void glVertex2f(GLfloat x, GLfloat y);                /* captured (z = 0): VolumeBuffer::drawSlice, screen-filling quads */
void glVertex3f(GLfloat x, GLfloat y, GLfloat z);      /* captured: cube faces of Renderer::drawCubeFaces */
/* captured (ref_host_driver.cpp): the slice / cap polygons of VV/slicing.cpp */
void glVertex3fv(const GLfloat *v);
void glMultiTexCoord3fvARB(GLenum unit, const GLfloat *v);
static inline void glTexCoord1f(GLfloat) {}
static inline void glTexCoord2f(GLfloat, GLfloat) {}
static inline void glColor3f(GLfloat, GLfloat, GLfloat) {}
static inline void glColor4f(GLfloat, GLfloat, GLfloat, GLfloat) {}
static inline void glColor4fv(const GLfloat *) {}
static inline void glColor3fv(const GLfloat *) {}
void glEnable(GLenum cap);                              /* captured: GL_CULL_FACE, GL_CLIP_PLANEi */
void glDisable(GLenum cap);
void glBlendFunc(GLenum src, GLenum dst);            /* captured */
static inline void glPushAttrib(GLbitfield) {}
static inline void glPopAttrib(void) {}
static inline void glLineWidth(GLfloat) {}
static inline void glPointSize(GLfloat) {}
static inline void glRasterPos2i(GLint, GLint) {}
static inline void glRasterPos2f(GLfloat, GLfloat) {}
static inline void glutBitmapCharacter(void *, int) {}
static inline int glutBitmapWidth(void *, int) { return 8; }
static inline void glutPostRedisplay(void) {}
#ifdef __cplusplus
}
#endif
#include "vv_gl_stub_renderer.h"
#endif
