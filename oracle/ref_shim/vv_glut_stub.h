/* vv_glut_stub.h -- TEST INFRASTRUCTURE ONLY.
 * The GLUT / GLEW surface VV/3DLIC.cpp needs to compile UNMODIFIED (its GLUT callbacks -- keyboard, keyboardSpecial, resize,
 * idle -- are then called directly by oracle/ref_host_driver.cpp; no window, no main loop). */
#ifndef VV_GLUT_STUB_H_
#define VV_GLUT_STUB_H_
enum {
    GLUT_KEY_F1 = 1, GLUT_KEY_F2, GLUT_KEY_F3, GLUT_KEY_F4, GLUT_KEY_F5, GLUT_KEY_F6, GLUT_KEY_F7, GLUT_KEY_F8, GLUT_KEY_F9, GLUT_KEY_F10,
    GLUT_KEY_F11, GLUT_KEY_F12, GLUT_KEY_LEFT = 100, GLUT_KEY_UP, GLUT_KEY_RIGHT, GLUT_KEY_DOWN,
    GLUT_ACTIVE_SHIFT = 1, GLUT_ACTIVE_CTRL = 2, GLUT_ACTIVE_ALT = 4, GLUT_DOWN = 0, GLUT_UP = 1,
    GLUT_LEFT_BUTTON = 0, GLUT_MIDDLE_BUTTON = 1, GLUT_RIGHT_BUTTON = 2,
    GLUT_RGBA = 0, GLUT_DOUBLE = 2, GLUT_ALPHA = 8, GLUT_DEPTH = 16, GLEW_OK = 0
};
#ifdef __cplusplus
extern "C" {
#endif
static inline void glutSwapBuffers(void) {}
static inline void glutIdleFunc(void (*)(void)) {}
int glutGetModifiers(void);                              /* set per event by oracle/ref_app_driver.cpp */
static inline int glutCreateMenu(void (*)(int)) { return 1; }
static inline void glutAddMenuEntry(const char *, int) {}
static inline void glutAttachMenu(int) {}
static inline void glutInit(int *, char **) {}
static inline void glutInitWindowPosition(int, int) {}
static inline void glutInitWindowSize(int, int) {}
static inline void glutInitDisplayMode(unsigned int) {}
static inline int glutCreateWindow(const char *) { return 1; }
static inline void glutDisplayFunc(void (*)(void)) {}
static inline void glutReshapeFunc(void (*)(int, int)) {}
static inline void glutKeyboardFunc(void (*)(unsigned char, int, int)) {}
static inline void glutSpecialFunc(void (*)(int, int, int)) {}
static inline void glutMotionFunc(void (*)(int, int)) {}
static inline void glutMouseFunc(void (*)(int, int, int, int)) {}
static inline void glutMainLoop(void) {}
static inline GLenum glewInit(void) { return GLEW_OK; }
static inline const GLubyte *glewGetErrorString(GLenum) { return (const GLubyte *)"stub"; }
static inline void glVertex3i(GLint, GLint, GLint) {}
static inline void glLightf(GLenum, GLenum, GLfloat) {}
#ifdef __cplusplus
}
#endif
#endif
